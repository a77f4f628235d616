#!/bin/bash
# ncu --set full capture of one step of a workload (run under gpurun, one GPU); the report
# stays on the box (it can exceed the 64 MiB return limit), its raw page comes back as CSV
# usage: tools/ncu_single.sh <workload> <launch-skip> <launch-count> [extra bench args]
set -e
W=$1; SKIP=${2:-30}; CNT=${3:-9}; shift 3 || true
mkdir -p gpurun_out
python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 "$@" > gpurun_out/plain_$W.json 2> gpurun_out/plain_$W.err
ncu --set full --clock-control none --import-source on -k regex:tf_k_ --launch-skip $SKIP --launch-count $CNT \
  -f -o /tmp/full_$W python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1 "$@" \
  > gpurun_out/ncu_$W.log 2>&1
ncu -i /tmp/full_$W.ncu-rep --page raw --csv > gpurun_out/full_${W}_raw.csv
