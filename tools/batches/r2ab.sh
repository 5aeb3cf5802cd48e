#!/usr/bin/env bash
# round-2 batch ab: TMA kernel A/B (early issue / input prefetch / bulk out), procedural rooms across mass resets
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 8"
for rep in 1 2; do
for v in default t_e1p0 t_e0p1 t_e1p1 t_nobulk; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  $B --workload tinyloc --steps 200 > $O/r2ab_tiny_${v}_$rep.json 2> $O/r2ab_tiny_${v}_$rep.err
done
done
unset APG_B200_LIB
$B --workload lidar_rooms --lidar-maps procedural --steps 202 > $O/r2ab_rooms_proc_k202.json 2> $O/r2ab_rooms_proc_k202.err
$B --workload lidar_rooms --steps 202 > $O/r2ab_rooms_bank_k202.json 2> $O/r2ab_rooms_bank_k202.err
python -m pytest tests/test_lidar_gpu.py -m gpu -q -k procedural > $O/r2ab_pytest.log 2>&1
tail -n 2 $O/r2ab_pytest.log
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
PY
