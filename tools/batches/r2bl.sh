#!/usr/bin/env bash
# round-2 batch bl: what do the autoreset payload draws cost next to the persistent TinyImageNetLoc window?
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 2 --workload tinyloc --steps 20"
APG_BENCH_DIAG=1 $B > $O/r2bl_diag.json 2> $O/r2bl_diag.err
APG_BENCH_DIAG=1 APG_IMG_TOP_UP=0 $B > $O/r2bl_diag_notopup.json 2> $O/r2bl_diag_notopup.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2bl_launches_tiny.csv $B --prime-ms 0 > /dev/null 2>&1
grep diag $O/r2bl_diag.err | head -12
echo ---
grep diag $O/r2bl_diag_notopup.err | head -12
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$O/r2bl_launches_tiny.csv")) if len(r)>10]
c=collections.defaultdict(list)
for r in rows[1:]:
    c[r[4][:60]].append(float(r[-1]))
for k,v in c.items(): print(k, len(v), 'mean ns %.0f'%(sum(v)/len(v)), 'max %.0f'%max(v))
PY
