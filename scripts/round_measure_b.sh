#!/bin/bash
# Round-2 final one-GPU pass: the driver's bench line, the reference arm, the launch list and one ncu --set full
# capture of the in-sweep kernels (draw / propagate / normalise / resample).  usage: scripts/round_measure_b.sh TAG
TAG=$1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_benchref.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/profile_target.py sweep > gpurun_out/${TAG}_ncu1.log 2>&1
APS_PROF_T=4 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'k_draw_normals|k_propagate|k_normalise|k_resample' -s 8 -c 4 -o gpurun_out/${TAG}_sweep \
    python scripts/profile_target.py sweep > gpurun_out/${TAG}_ncu3.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/${TAG}_smi.csv
tail -2 gpurun_out/${TAG}_ncu1.log gpurun_out/${TAG}_ncu3.log; tail -c 600 gpurun_out/${TAG}_bench.json
