#!/bin/bash
# Instruction count / issue utilisation of the tuned kernels for one library variant (cheap ncu pass):
#   profiles/quick_ncu.sh TAG [G] [n_groups]      (SPECKV_LIB selects the variant)
TAG=$1; G=${2:-131072}; N=${3:-8192}
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:fast_kernel --launch-skip 8 -c 2 --csv python profiles/quick_time.py $G $N 1 2>/dev/null \
  | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r))
    print('$TAG', d['Kernel Name'][:40].split('::')[-1], d['Metric Name'], d['Metric Value'])
"
