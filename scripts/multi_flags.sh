#!/bin/bash
# In-graph K1/K2/K3 timeline of the sharded headline sweep under the diagnostic flags (needs a -DAPS_TIMELINE=1
# build in $APS_LIB_PATH). usage: scripts/multi_flags.sh NGPU FLAG...   (flags: 1 skip waits, 4 local gathers,
# 8 local scatter; 16 is added for the timeline). Results are WRONG for flags other than 0.
N=$1; shift
for f in "$@"; do
  echo "== world $N APS_DEBUG_MULTI=$((f | 16))"
  APS_DEBUG_SPIN=1 APS_DEBUG_MULTI=$((f | 16)) timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
     --master-addr 127.0.0.1 --master-port $((29700 + f)) scripts/time_sharded.py 2>&1 | grep -E "aps rank 0|^world" | tail -3
done
