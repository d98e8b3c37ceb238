#!/bin/bash
# timing diagnostics of the sharded sweep on 2 GPUs (results are WRONG for flags != 0 and != 16)
# flags (APS_DEBUG_MULTI): 1 skip waits, 4 local parent gathers, 8 local ancestor scatter, 16 kernel spans
for f in "$@"; do
  echo "== APS_DEBUG_MULTI=$f"
  APS_DEBUG_SPIN=1 APS_DEBUG_MULTI=$f timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((f % 10)) bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep -E "^\{|aps rank 0" | tail -3 | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print(d['ms_per_step'], {k: round(v['avg_us'],1) for k,v in d['roofline']['kernels'].items()})
    else: print(ln.strip()[:200])"
done
