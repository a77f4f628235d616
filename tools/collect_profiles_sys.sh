#!/bin/bash
# Refresh of the system-resident capture of tools/collect_profiles_r2b.sh (final kernel)
OUT=gpurun_out/prof_r2b
mkdir -p $OUT
B="--steps 3 --warmup 3 --no-cpu --no-others --e2e-steps 1"
python bench.py $B --members 8192 > $OUT/plain_ens8192.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tf_k_sysstep --launch-skip 4 --launch-count 1 \
  -f -o /tmp/full_sysstep python bench.py $B --members 8192 > $OUT/ncu_full_sysstep.log 2>&1
ncu -i /tmp/full_sysstep.ncu-rep --page raw --csv > /tmp/full_sysstep_raw.csv
python - /tmp/full_sysstep_raw.csv $OUT/ncu_full_sysstep_8192.csv <<'PY'
import csv, re, sys
KEYS = r'Kernel Name|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|gpu__time_duration.sum|launch__block_size|launch__grid_size|launch__registers_per_thread$|launch__shared_mem_per_block_dynamic|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__inst_executed.sum$|smsp__issue_active.avg.pct_of_peak_sustained_active|lts__t_bytes.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__sass_thread_inst_executed_op_d(fma|add|mul)_pred_on.sum$|smsp__sass_inst_executed_op_local_(ld|st).sum$|sm__cycles_elapsed.max$'
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 20]
pat = re.compile(KEYS)
keep = [i for i, h in enumerate(rows[0]) if pat.fullmatch(h)]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep])
PY
tail -2 $OUT/ncu_full_sysstep.log
