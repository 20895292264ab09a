#!/bin/bash
# One gpurun call: the default bench line, the ncu launch list of the same bench command and one full capture of
# each hot kernel (each only after its own command ran clean without ncu).  Outputs land in gpurun_out/;
# scripts/make_profiles.sh turns them into the tracked summaries under profiles/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
# the profiled command is the bench step without the CPU legs and the extra measurements
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --autolink-new 0 --strong-rows 0 --opt graphs=0"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tensor_scan_kernel -s 12 -c 4 -f -o gpurun_out/prof_k2_b1024 $B > gpurun_out/ncu_k2.log 2>&1
echo "k2 capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stream_scan_kernel -s 3 -c 2 -f -o gpurun_out/prof_k1_b1 $B > gpurun_out/ncu_k1.log 2>&1
echo "k1 (bf16 shadow) capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stream_scan_kernel -s 3 -c 2 -f -o gpurun_out/prof_k1_b1_fp32 $B --opt stream_bf16=0 > gpurun_out/ncu_k1f.log 2>&1
echo "k1 (fp32 rows) capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:select_rescore_kernel -s 6 -c 2 -f -o gpurun_out/prof_select_b1024 $B > gpurun_out/ncu_sel.log 2>&1
echo "select capture rc=$?"
# the auto-link shape (k=100): one call = bootstrap + phases of tensor_scan_kernel; capture the second call
P="env CORTEX_GPU_LIB=cortex_b200/libcortex_gpu.so python scripts/k2_probe.py --batch 1024 --k 100 --debug-modes 0 --growth 6 --pairs 0 --epi 8 --reps 2"
$P > gpurun_out/plain_k100.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tensor_scan_kernel -s 4 -c 4 -f -o gpurun_out/prof_k2_k100 $P > gpurun_out/ncu_k2_k100.log 2>&1
echo "k2 k=100 capture rc=$?"
A="python scripts/al_probe.py --graphs 0 --reps 1"
$A > gpurun_out/plain_al.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_autolink.csv $A > gpurun_out/ncu_al.log 2>&1
echo "autolink launch list rc=$?"
