set -u
mkdir -p gpurun_out
cp profiles/results_r2.json gpurun_out/results_r2.json
timeout 600 python bench_configs.py --only 2 --out gpurun_out/results_r2.json 2>&1 | cut -c1-330 | tail -16
bash scripts/gpu_check.sh
