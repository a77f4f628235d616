#!/bin/bash
# Final bench lines of the round (plain runs, no profiler), one B200: gpurun_out/prof_r2b/final_*.json
OUT=gpurun_out/prof_r2b
mkdir -p $OUT
python bench.py > $OUT/final_bench_default.json 2> $OUT/final_bench_default.err
for w in ks burgers film; do
  python bench.py --workload $w --no-others > $OUT/final_bench_$w.json 2> $OUT/final_bench_$w.err
done
python bench.py --members 4096 --no-cpu --no-others > $OUT/final_bench_members4096.json 2> $OUT/final_bench_members4096.err
python tools/gs_check.py time > $OUT/final_gs_time.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/final_smoke.log 2>&1
tail -2 $OUT/final_smoke.log
