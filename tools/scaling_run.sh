#!/bin/bash
# usage: tools/scaling_run.sh <tag> [N ...]  -- bench.py under torchrun at every N (default 2 4 8), one JSON line each
cd "$(dirname "$0")/.."
tag=$1; shift
ns=${@:-2 4 8}
for n in $ns; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  echo "N=$n rc=$? $(tail -c 300 gpurun_out/scale_${tag}_n$n.err | tr '\n' ' ' | tail -c 200)"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${tag}_n$n.json").read().strip().splitlines()[-1])
    c=d.get("config3") or {}
    print("  value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "parity", d["parity_check"]["ok"], "exchange", d.get("exchange_ms"), "| config3 ms", c.get("ms_per_step"), "spans", c.get("spans"), "ok", (c.get("parity_check") or {}).get("ok"))
    print("  kernels", {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()}, "c3", c.get("kernels_ms_per_step_rank0"))
except Exception as e:
    print("  no line:", e)
PY
done
