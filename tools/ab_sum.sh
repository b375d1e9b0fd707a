#!/bin/bash
# usage: tools/ab_sum.sh N -- weak-scaling bench at N GPUs with the table sum by the peer-memory kernel and by NCCL
cd "$(dirname "$0")/.."
n=$1
for mode in 1 0 1 0; do
  KS_PEER_SUM=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + mode)) \
    bench.py --gpus $n --steps 20 --warmup 3 --no-config3 --no-extra > gpurun_out/ab_$mode.json 2> gpurun_out/ab_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab_$mode.json').read().strip().splitlines()[-1]); print('KS_PEER_SUM=$mode', round(d['value'],1), round(d['ms_per_step'],3), d['config'].get('sharding','')[:80], d['exchange_ms'], d['parity_check']['ok'])"
done
