#!/bin/bash
# Refresh of the two grid-resident captures of tools/collect_profiles_r2b.sh (final kernels)
OUT=gpurun_out/prof_r2b
mkdir -p $OUT
KEYS='Kernel Name|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|gpu__time_duration.sum|launch__block_size|launch__grid_size|launch__registers_per_thread$|launch__shared_mem_per_block_dynamic|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__inst_executed.sum$|smsp__issue_active.avg.pct_of_peak_sustained_active|lts__t_bytes.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on.sum$|smsp__sass_inst_executed_op_local_(ld|st).sum$|sm__cycles_elapsed.max$'
reduce() {
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
python tools/one_case.py ks 1048576 6 > $OUT/plain_ks_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_gridstep --launch-skip 3 --launch-count 1 \
  -f -o /tmp/full_gs python tools/one_case.py ks 1048576 6 > $OUT/ncu_full_gridstep.log 2>&1
ncu -i /tmp/full_gs.ncu-rep --page raw --csv > /tmp/full_gs_raw.csv && reduce /tmp/full_gs_raw.csv $OUT/ncu_full_gridstep_ks.csv
python tools/slab_one.py 1048576 6 > $OUT/plain_slab_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_gridstep_mr --launch-skip 4 --launch-count 1 \
  -f -o /tmp/full_gsmr python tools/slab_one.py 1048576 6 > $OUT/ncu_full_gridstep_mr.log 2>&1
ncu -i /tmp/full_gsmr.ncu-rep --page raw --csv > /tmp/full_gsmr_raw.csv && reduce /tmp/full_gsmr_raw.csv $OUT/ncu_full_gridstep_mr_ks.csv
TF_CFLAGS=-DTF_GS_TRACE python tools/gs_trace.py ks 1048576 > $OUT/trace_gridstep_ks.txt 2>&1
TF_CFLAGS=-DTF_GS_TRACE python tools/gs_trace.py burgers 131072 > $OUT/trace_gridstep_burgers.txt 2>&1
tail -3 $OUT/trace_gridstep_ks.txt
