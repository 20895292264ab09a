#!/bin/bash
# Full GPU test suite, then one ncu capture of the select/rescore kernel inside the default bench.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-small-probe"
$B > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:select_rescore_kernel -s 6 -c 2 -f -o gpurun_out/prof_select_b1024 $B > gpurun_out/ncu_sel.log 2>&1
echo "select capture rc=$?"
