#!/bin/bash
# usage: ncu_cycles.sh <kernel regex> <python script> [args]  -> duration / elapsed SM cycles / tensor-pipe % of the 3rd launch
K=$1; shift
M=gpu__time_duration.sum,sm__cycles_elapsed.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.avg.per_cycle_active
ncu --metrics $M --clock-control none -k regex:$K -c 3 --csv python "$@" 2>/dev/null | python -c "
import csv,sys
rows=[r for r in csv.reader(sys.stdin) if len(r)>10]
h=rows[0]
out={}
for r in rows[1:]:
    if r[h.index('ID')]=='2': out[r[h.index('Metric Name')].split('.')[0].replace('__','_')]=r[h.index('Metric Value')]
print(out)
"
