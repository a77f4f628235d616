#!/bin/bash
# Round-2 (final build) profile collection (run under gpurun on one B200).  Plain runs first (each must exit 0
# before the same command line is profiled), then ncu launch lists, then one ncu --set full
# capture per dominant kernel, reduced to CSV.  Everything lands in gpurun_out/prof_r2b/.
OUT=gpurun_out/prof_r2b
mkdir -p $OUT
python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err || exit 1
for w in ks burgers film; do
  python bench.py --workload $w --no-others > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
B="--steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1"
KEYS='Kernel Name|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|gpu__time_duration.sum|launch__block_size|launch__grid_size|launch__registers_per_thread$|launch__shared_mem_per_block_dynamic|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__inst_executed.sum$|smsp__issue_active.avg.pct_of_peak_sustained_active|lts__t_bytes.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on.sum$|smsp__sass_inst_executed_op_local_(ld|st).sum$|sm__cycles_elapsed.max$'
reduce() {  # raw-page CSV -> the columns above
python - "$1" "$2" "$KEYS" <<'PY'
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
rows = [r for r in rows if len(r) > 20]
hdr = rows[0]
pat = re.compile(sys.argv[3])
keep = [i for i, h in enumerate(hdr) if pat.fullmatch(h)]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
PY
}
python bench.py $B > $OUT/plain_ens.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_ensemble.csv \
  python bench.py $B > $OUT/ncu_launches_ensemble.log 2>&1
python bench.py $B --workload ks > $OUT/plain_ks.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_ks.csv \
  python bench.py $B --workload ks > $OUT/ncu_launches_ks.log 2>&1
python bench.py $B --members 8192 > $OUT/plain_ens8192.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_sysstep --launch-skip 4 --launch-count 1 \
  -f -o /tmp/full_sysstep python bench.py $B --members 8192 > $OUT/ncu_full_sysstep.log 2>&1
ncu -i /tmp/full_sysstep.ncu-rep --page raw --csv > /tmp/full_sysstep_raw.csv && reduce /tmp/full_sysstep_raw.csv $OUT/ncu_full_sysstep_8192.csv
python tools/one_case.py film 262144 6 > $OUT/plain_film.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_ --launch-skip 12 --launch-count 4 \
  -f -o /tmp/full_film python tools/one_case.py film 262144 6 > $OUT/ncu_full_film.log 2>&1
ncu -i /tmp/full_film.ncu-rep --page raw --csv > /tmp/full_film_raw.csv && reduce /tmp/full_film_raw.csv $OUT/ncu_full_film.csv
python tools/one_case.py ks 1048576 6 > $OUT/plain_ks_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_gridstep --launch-skip 3 --launch-count 1 \
  -f -o /tmp/full_gs python tools/one_case.py ks 1048576 6 > $OUT/ncu_full_gridstep.log 2>&1
ncu -i /tmp/full_gs.ncu-rep --page raw --csv > /tmp/full_gs_raw.csv && reduce /tmp/full_gs_raw.csv $OUT/ncu_full_gridstep_ks.csv
python tools/slab_one.py 1048576 6 > $OUT/plain_slab_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_gridstep_mr --launch-skip 4 --launch-count 1 \
  -f -o /tmp/full_gsmr python tools/slab_one.py 1048576 6 > $OUT/ncu_full_gridstep_mr.log 2>&1
ncu -i /tmp/full_gsmr.ncu-rep --page raw --csv > /tmp/full_gsmr_raw.csv && reduce /tmp/full_gsmr_raw.csv $OUT/ncu_full_gridstep_mr_ks.csv
python bench.py $B --workload burgers > $OUT/plain_burgers.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_burgers.csv \
  python bench.py $B --workload burgers > $OUT/ncu_launches_burgers.log 2>&1
TF_CFLAGS=-DTF_GS_TRACE python tools/gs_trace.py ks 1048576 > $OUT/trace_gridstep_ks.txt 2>&1
TF_CFLAGS=-DTF_GS_TRACE python tools/gs_trace.py burgers 131072 > $OUT/trace_gridstep_burgers.txt 2>&1
ls -la $OUT
