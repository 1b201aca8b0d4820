#!/bin/bash
# quick GPU visit: tests (optional -k filter) + device-only bench + stage times
TAG=${1:-q}; KF=${2:-}
mkdir -p gpurun_out
if [ -n "$KF" ]; then timeout 600 python -m pytest tests -m gpu -q -x -k "$KF" > gpurun_out/pytest_$TAG.log 2>&1; else timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$TAG.log 2>&1; fi
tail -6 gpurun_out/pytest_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print('ms/frame', round(d['ms_per_step'],4), 'Mrays/s', round(d['value']/1e6,2)); print({k:v['ms'] for k,v in d['stages_ms'].items()})"
