set -u
mkdir -p gpurun_out
export CORTEX_GPU_LIB=$PWD/cortex_b200/libcortex_gpu.so
for b in 8 128; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tensor_scan|tau_|select_rescore|query_bf16|prepare_q" -c 20 --csv --log-file gpurun_out/launches_b$b.csv python scripts/k2_probe.py --batch $b --debug-modes 0 --pairs 0 --growth -1 --reps 2 > gpurun_out/ncu_b$b.log 2>&1; echo rc=$?
done
CX_BATCHES=3,8,64,128 timeout 300 python scripts/e2e_probe.py 2>&1 | grep pageable
