#!/usr/bin/env bash
# round-2 batch ac: LIDAR step kernel, resident blocks per SM (register budget) A/B
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --steps 20 --no-cpu-baseline --e2e-steps 4"
for rep in 1 2; do
for v in default l_mb5 l_mb6 l_mb7; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  $B --workload lidar_rooms > $O/r2ac_rooms_${v}_$rep.json 2> $O/r2ac_rooms_${v}_$rep.err
done
done
for v in default l_mb6; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  $B --workload lidar_maze > $O/r2ac_maze_${v}.json 2> $O/r2ac_maze_${v}.err
done
unset APG_B200_LIB
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2ac_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
PY
