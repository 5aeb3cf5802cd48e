#!/usr/bin/env bash
# round-2 batch bd: blocked-bank transport as the default (per-lane constant issue) vs the tensor-map variant; full GPU suite
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 4"
for v in default t_tensor; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  for rep in 1 2; do
    $B --workload tinyloc --steps 200 > $O/r2bd_tiny_${v}_k200_$rep.json 2> $O/r2bd_tiny_${v}_k200_$rep.err
    $B --workload tinyloc --steps 20 > $O/r2bd_tiny_${v}_k20_$rep.json 2> $O/r2bd_tiny_${v}_k20_$rep.err
  done
done
unset APG_B200_LIB
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum \
    --clock-control none -k regex:tma_step -s 4 -c 1 --csv --log-file $O/r2bd_ncu_default.csv \
    $B --workload tinyloc --steps 20 --prime-ms 0 > /dev/null 2>&1
python -m pytest tests -m gpu -q > $O/r2bd_pytest.log 2>&1
tail -n 3 $O/r2bd_pytest.log
python - <<PY
import json,glob,csv
for f in sorted(glob.glob("$O/r2bd_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("$O/r2bd_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f.split('/')[-1], [(r[-3],r[-1]) for r in rows[1:]])
PY
