set -u
mkdir -p gpurun_out
timeout 300 python scripts/k1_probe.py 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:stream_scan_kernel -s 4 -c 1 -f -o gpurun_out/prof_k1h_b1 python scripts/k1_probe.py --batches 1 --bf16 1 --reps 3 > gpurun_out/ncu_k1h.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:stream_scan_kernel -s 4 -c 1 -f -o gpurun_out/prof_k1h_b4 python scripts/k1_probe.py --batches 4 --bf16 1 --reps 3 > gpurun_out/ncu_k1h4.log 2>&1; echo rc=$?
