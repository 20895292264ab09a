#!/bin/bash
# Quick look at the tensor pass: its parity tests, then bench A/B (CTA pair vs single CTA).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -k "tensor_pass" -x -q > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_k2.log
for pair in 1 0; do
  timeout 300 python bench.py --no-cpu-baseline --no-small-probe --opt tensor_pair=$pair > gpurun_out/bench_pair$pair.json 2> gpurun_out/bench_pair$pair.err; echo "bench pair=$pair rc=$?"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_pair$pair.json"))
    print("pair=$pair value", j["value"], "roofline", j["roofline"]["achieved"], j["roofline"]["frac"], "us", j["roofline"]["us_per_launch"], j["paths"], j["clocks"])
except Exception as e:
    print("no json", e)
PY
done
