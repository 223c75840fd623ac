#!/bin/bash
# Round-2 GPU call 3: k_pd_tv3d_f2t v2 (static stages, uniform operands): agreement, timing, DRAM traffic per variant
set -u
mkdir -p gpurun_out
timeout 400 python -u tools/check_f2t.py 256 1024 512 2048 > gpurun_out/r2c4_check_f2t.log 2>&1
grep "PD_TV\|MISMATCH\|agreement\|Error\|error" gpurun_out/r2c4_check_f2t.log | tail -50
for cfg in "4 4" "4 2" "5 2" "2 4"; do
  set -- $cfg
  TMB_TV_HOOK=11 TMB_F2T_CFG="$1 $2" timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:k_pd_tv3d_f2t -c 1 --csv --log-file gpurun_out/r2c4_ncu_$1_$2.csv python tools/prof_tv.py 2048 512 2 > /dev/null 2>&1
  python - "$1" "$2" <<'PY'
import csv, sys
rows = list(csv.reader(l for l in open(f"gpurun_out/r2c4_ncu_{sys.argv[1]}_{sys.argv[2]}.csv") if l.startswith('"')))
h = rows[0]
print("cfg", sys.argv[1:], {r[h.index("Metric Name")]: r[h.index("Metric Value")] for r in rows[1:]})
PY
done
