#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, the ncu launch list of the same
# bench command and one full capture of each scan kernel.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cat gpurun_out/bench_default.json
# the profiled command is the bench step without the CPU leg and without the extra auto-link measurement
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --autolink-new 0"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tensor_scan_kernel -s 9 -c 3 -f -o gpurun_out/prof_k2_b1024 $B > gpurun_out/ncu_k2.log 2>&1
echo "k2 capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stream_scan_kernel -s 3 -c 2 -f -o gpurun_out/prof_k1_b1 $B > gpurun_out/ncu_k1.log 2>&1
echo "k1 capture rc=$?"
# the auto-link shape (k=100): one call = bootstrap + 3 phases of tensor_scan_kernel; capture the second call
P="python scripts/k2_probe.py --batch 1024 --k 100 --debug-modes 0 --growth 4 --pairs 0 --epi 8 --reps 2"
$P > gpurun_out/plain_k100.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tensor_scan_kernel -s 4 -c 4 -f -o gpurun_out/prof_k2_k100 $P > gpurun_out/ncu_k2_k100.log 2>&1
echo "k2 k=100 capture rc=$?"
