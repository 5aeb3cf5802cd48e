#!/usr/bin/env bash
# round-2 batch aa: procedural-dataset mode of the LIDAR engine; TMA kernel A/B (early issue / input prefetch / bulk out)
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_lidar_gpu.py tests/test_maps.py -m gpu -q > $O/r2aa_pytest.log 2>&1
tail -n 4 $O/r2aa_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 16"
$B --workload lidar_rooms > $O/r2aa_rooms_bank.json 2> $O/r2aa_rooms_bank.err
$B --workload lidar_rooms --lidar-maps procedural > $O/r2aa_rooms_proc.json 2> $O/r2aa_rooms_proc.err
$B --workload lidar_maze --lidar-maps procedural > $O/r2aa_maze_proc.json 2> $O/r2aa_maze_proc.err
$B --workload lidar_rooms --lidar-maps procedural --steps 202 > $O/r2aa_rooms_proc_k202.json 2> $O/r2aa_rooms_proc_k202.err
for v in default t_e1p0 t_e0p1 t_e1p1 t_nobulk; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  $B --workload tinyloc --steps 200 > $O/r2aa_tiny_$v.json 2> $O/r2aa_tiny_$v.err
done
unset APG_B200_LIB
python - <<PY
import json
for n in ["rooms_bank","rooms_proc","maze_proc","rooms_proc_k202","tiny_default","tiny_t_e1p0","tiny_t_e0p1","tiny_t_e1p1","tiny_t_nobulk"]:
    try:
        d=json.loads(open("$O/r2aa_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "flags", d.get("error_flags"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(n, "FAILED", e)
PY
