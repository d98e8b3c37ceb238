#!/bin/bash
# One-GPU measurement pass for profiles/: tests, bench (both arms), launch list, ncu --set full captures.
# usage: scripts/round_measure.sh TAG
TAG=$1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${TAG}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_benchref.json 2> gpurun_out/${TAG}_benchref.err
python scripts/bench_configs.py > gpurun_out/${TAG}_configs.log 2>&1
# launch list of 2 short sweeps (plain launches) + the isolated resample kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_target.py both > gpurun_out/${TAG}_ncu1.log 2>&1
# full capture: the graded kernel in isolation (N = 2^25), then the three kernels of a sweep at N = 1e6
ncu --set full --clock-control none --import-source on -k regex:k_resample -s 3 -c 1 -o gpurun_out/${TAG}_k3 \
    python scripts/profile_target.py resample > gpurun_out/${TAG}_ncu2.log 2>&1
APS_PROF_T=4 ncu --set full --clock-control none --import-source on -k regex:'k_propagate|k_normalise|k_resample' -s 6 -c 3 -o gpurun_out/${TAG}_sweep \
    python scripts/profile_target.py sweep > gpurun_out/${TAG}_ncu3.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/${TAG}_smi.csv
