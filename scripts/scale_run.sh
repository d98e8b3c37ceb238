#!/bin/bash
# Weak-scaling pass on one multi-GPU box: bench.py at 1/2/4/8 ranks back to back (the driver's SCALE contract),
# the one-process-per-GPU parity tests and the exchange stress test. usage: scripts/scale_run.sh TAG [MAXGPUS]
TAG=$1
MAX=${2:-8}
timeout 900 python -m pytest tests/test_gpu_multiprocess.py tests/test_gpu_stress.py -q -x 2>&1 | tail -6 > gpurun_out/${TAG}_mp_pytest.log
tail -3 gpurun_out/${TAG}_mp_pytest.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_scale_1gpu.json 2> gpurun_out/${TAG}_scale_1gpu.err
for n in 2 4 8; do
  [ $n -gt $MAX ] && break
  extra="--no-cpu-baseline"; [ $n -eq $MAX ] && extra=""
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 $extra > gpurun_out/${TAG}_scale_${n}gpu.json 2> gpurun_out/${TAG}_scale_${n}gpu.err
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open("gpurun_out/${TAG}_scale_%dgpu.json" % n))
    except Exception as e:
        print(n, "missing", e); continue
    if n == 1: base = d["value"]
    print("gpus %d value %.4g ms %.3f eff %.3f e2e %.4g" % (n, d["value"], d["ms_per_step"], d["value"] / (n * base) if base else 0, d["e2e"]["value"]))
    if "configs" in d:
        for k, v in d["configs"].items(): print("   ", k, v)
    oc = d.get("logZ", {}).get("oracle_check")
    if oc: print("    oracle_check abs_err", oc["abs_err"], "n", oc["n_particles"])
PY
