set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py -x -q 2>&1 | tail -4
bash scripts/gpu_check.sh
