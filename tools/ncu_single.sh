#!/bin/bash
# ncu --set full capture of one step of a single-grid workload (run under gpurun, one GPU)
# usage: tools/ncu_single.sh <workload> <launch-skip> <launch-count>
set -e
W=$1; SKIP=${2:-30}; CNT=${3:-9}
mkdir -p gpurun_out
python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 > gpurun_out/plain_$W.json 2> gpurun_out/plain_$W.err
ncu --set full --clock-control none --import-source on -k regex:tf_k_ --launch-skip $SKIP --launch-count $CNT \
  -f -o gpurun_out/full_$W python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 \
  > gpurun_out/ncu_$W.log 2>&1
ls -la gpurun_out
