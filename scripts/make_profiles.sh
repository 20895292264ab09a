#!/bin/bash
# After `gpurun -- bash scripts/gpu_check.sh`: turn what came back in gpurun_out/ into the tracked
# summaries under profiles/.  usage: scripts/make_profiles.sh r2
set -eu
R=${1:-r2}
python scripts/ncu_summary.py gpurun_out/prof_k2_b1024.ncu-rep > profiles/${R}_k2_tensor_scan_b1024.txt
python scripts/ncu_summary.py gpurun_out/prof_k2_k100.ncu-rep > profiles/${R}_k2_tensor_scan_b1024_k100.txt
python scripts/ncu_summary.py gpurun_out/prof_k1_b1.ncu-rep > profiles/${R}_k1_stream_scan_b1.txt
python scripts/ncu_summary.py gpurun_out/prof_k1_b1_fp32.ncu-rep > profiles/${R}_k1_stream_scan_b1_fp32.txt
python scripts/ncu_summary.py gpurun_out/prof_select_b1024.ncu-rep > profiles/${R}_select_rescore_b1024.txt
python scripts/launch_summary.py gpurun_out/launches.csv | cut -c1-110 > profiles/${R}_launches_bench_default.txt
python scripts/launch_summary.py gpurun_out/launches_autolink.csv | cut -c1-110 > profiles/${R}_launches_autolink_16k_x_1m.txt
python scripts/ncu_summary.py --traffic tensor_scan_kernel:1024:gpurun_out/prof_k2_b1024.ncu-rep:sum \
    stream_scan_kernel:1:gpurun_out/prof_k1_b1.ncu-rep:mean \
    stream_scan_kernel:1:gpurun_out/prof_k1_b1_fp32.ncu-rep:mean:stream_scan_kernel_fp32 > profiles/ncu_traffic.json
cp gpurun_out/bench_default.json profiles/bench_${R}_n1.json
ls -la profiles/
