#!/bin/bash
# what bounds the fused gather → density kernel: run the frame with the producers' gathers off (1), with the head off (2)
for d in 0 1 2; do
  GPNERF_FUSED_DEBUG=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-strong > gpurun_out/whatif_$d.json 2> gpurun_out/whatif_$d.err
  python -c "
import json; d=json.load(open('gpurun_out/whatif_$d.json')); print('debug=$d', {k:v['ms'] for k,v in d['stages_ms'].items() if 'k23' in k or 'color' in k})"
done
