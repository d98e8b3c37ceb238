#!/bin/bash
# Multi-GPU check on one box: the GPU test suite (includes the multi-process parity test when the box
# has >= 2 GPUs), then bench.py at each rank count given.
# usage: scripts/scale_run.sh TAG N [N ...]     (results in gpurun_out/TAG_*)
TAG=$1; shift
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest.log
for n in "$@"; do
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench$n.json 2> gpurun_out/${TAG}_bench$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${TAG}_bench$n.json 2> gpurun_out/${TAG}_bench$n.err
  fi
done
