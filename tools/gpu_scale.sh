#!/bin/bash
# bench.py at N GPUs (frames = weak, tiles = strong) + the peer-exchange check.  usage: tools/gpu_scale.sh N tag
N=$1; TAG=$2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 300 bash -c "$(declare -f run); N=$N; run 29511 tools/gpu_peer_check.py" 2>&1 | grep "^{"
for sh in frames tiles; do
  timeout 400 bash -c "$(declare -f run); N=$N; run 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --shard $sh" > gpurun_out/bench_n${N}_${sh}_$TAG.json 2> gpurun_out/bench_n${N}_${sh}_$TAG.err
  grep "^{" gpurun_out/bench_n${N}_${sh}_$TAG.json | python -c "
import sys, json
for l in sys.stdin:
    b = json.loads(l); print('$sh', b['n_gpus'], 'value', round(b['value'] / 1e6, 2), 'Mrays/s', 'ms', round(b['ms_per_step'], 4), b['scaling'], 'e2e', b['e2e'] and round(b['e2e']['value'] / 1e6, 2), b['config'].get('peer_check'))"
done
