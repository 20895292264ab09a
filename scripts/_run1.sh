set -u
mkdir -p gpurun_out
export CORTEX_GPU_LIB=$PWD/cortex_b200/libcortex_gpu.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_edges.py tests/test_gpu_threshold.py tests/test_gpu_autolink.py -x -q 2>&1 | tail -4
timeout 600 python scripts/k2_probe.py --batch 1024 --debug-modes 0 --pairs 0 --growth -1 --leftover 1,0,1,0,1,0 --reps 30 2>&1 | tail -8
timeout 600 python scripts/k2_probe.py --batch 1024 --k 100 --debug-modes 0 --pairs 0 --growth -1 --leftover 1,0,1,0 --reps 10 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_lo.json 2> gpurun_out/bench_lo.err; echo "bench rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/bench_lo.json')); print(j['value'], j['ms_per_step'], j['roofline'], j['parity'], j['clocks'])"
