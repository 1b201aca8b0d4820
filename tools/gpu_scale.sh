#!/bin/bash
# one N-GPU visit: render bench (frames + strong sub-record) and the training step; usage: tools/gpu_scale.sh N tag
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
tail -2 gpurun_out/bench_n${N}_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}_$TAG.json')); print('N', d['n_gpus'], 'frames ms', round(d['ms_per_step'],4), 'Mrays/s', round(d['value']/1e6,1), d['scaling']); print('strong', {k:(round(v['ms_per_frame'],4), v['full_image_on_every_rank']) for k,v in d['strong'].items()}); print('e2e Mrays/s', round(d['e2e']['value']/1e6,1), 'ms', round(d['e2e']['ms_per_step'],3), 'dense', round(d['e2e']['dense_levels']['ms_per_step'],3)); print('clocks', d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 10 > gpurun_out/bench_train_n${N}_$TAG.json 2> gpurun_out/bench_train_n${N}_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_train_n${N}_$TAG.json')); print('train N', d['n_gpus'], 'ms', round(d['ms_per_step'],3), 'rays/s', round(d['value']), 'full', d['full_pipeline'].get('ms_per_step'), d['config']['grads_identical_on_all_ranks'])"
