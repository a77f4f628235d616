#!/bin/bash
# ncu --set full capture of ONE kernel launch + per-source-line stall samples (CSV), run under gpurun
# usage: tools/ncu_kernel.sh <tag> <kernel-regex> <launch-skip> -- <bench args...>
set -e
TAG=$1; KRE=$2; SKIP=$3; shift 4
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err
ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip $SKIP --launch-count 1 \
  -f -o /tmp/k_$TAG python bench.py "$@" > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/k_$TAG.ncu-rep --page raw --csv > gpurun_out/k_${TAG}_raw.csv
ncu -i /tmp/k_$TAG.ncu-rep --page source --print-source cuda --csv > gpurun_out/k_${TAG}_src.csv 2>/dev/null || true
ncu -i /tmp/k_$TAG.ncu-rep --page source --print-source sass --csv > /tmp/k_${TAG}_sass.csv 2>/dev/null || true
python - <<PY
import csv
rows = list(csv.reader(open("/tmp/k_${TAG}_sass.csv")))
print(rows[0])
PY
gzip -f -k /tmp/k_${TAG}_sass.csv && cp /tmp/k_${TAG}_sass.csv.gz gpurun_out/ || true
ls -la gpurun_out
