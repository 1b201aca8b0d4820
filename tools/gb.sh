#!/bin/bash
# build here (fails loudly), then run the given command on the GPU box
set -e
python gp-nerf_b200/build.py > /tmp/build.log 2>&1 || { grep -E "error" /tmp/build.log | head; echo BUILD FAILED; exit 1; }
T=${GB_TIMEOUT:-900}
gpurun --timeout $T -- "$@" 2>&1 | tail -${GB_TAIL:-14}
