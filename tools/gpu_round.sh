#!/bin/bash
# One GPU-box visit: tests, bench, launch list, ncu --set full of the two head kernels.
# usage: tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-rXX}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
  tail -5 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.json
timeout 600 python tools/bench_configs.py dense512 zju1024 thu512 > gpurun_out/configs_$TAG.jsonl 2>> gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/b_launch_$TAG.log 2>&1
for k in gather_density_tc color_mlp_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/prof_${TAG}_$k \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > gpurun_out/b_$k.log 2>&1
done
ls -la gpurun_out | tail -8
