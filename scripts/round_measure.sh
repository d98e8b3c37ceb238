#!/bin/bash
# One-GPU ncu pass for profiles/: launch list + ncu --set full captures.  usage: scripts/round_measure.sh TAG
TAG=$1
# launch list of 2 short sweeps (plain launches) + the isolated resample kernel + one fused sweep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_target.py both > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/${TAG}_launches_fused.csv \
    python scripts/profile_target.py fused > gpurun_out/${TAG}_ncu1f.log 2>&1
# full captures: the graded kernel in isolation (N = 2^25, cold L2), the three kernels of a sweep at N = 1e6
# (cache control off: inside a sweep their inputs are L2-resident), the fused persistent kernel (T = 4)
ncu --set full --clock-control none --import-source on -k regex:k_resample -s 3 -c 1 -o gpurun_out/${TAG}_k3 \
    python scripts/profile_target.py resample > gpurun_out/${TAG}_ncu2.log 2>&1
APS_PROF_T=4 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_propagate|k_normalise|k_resample' -s 6 -c 3 -o gpurun_out/${TAG}_sweep \
    python scripts/profile_target.py sweep > gpurun_out/${TAG}_ncu3.log 2>&1
APS_PROF_T=4 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_sweep_fused -s 1 -c 1 -o gpurun_out/${TAG}_fused \
    python scripts/profile_target.py fused > gpurun_out/${TAG}_ncu4.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/${TAG}_smi.csv
tail -2 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu1f.log gpurun_out/${TAG}_ncu2.log gpurun_out/${TAG}_ncu3.log gpurun_out/${TAG}_ncu4.log
