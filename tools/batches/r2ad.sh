#!/usr/bin/env bash
# round-2 batch ad: LIDAR step kernel, more resident blocks per SM
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --steps 20 --no-cpu-baseline --e2e-steps 4"
for v in default l_mb9 l_mb10 l_mb12; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  $B --workload lidar_rooms > $O/r2ad_rooms_${v}.json 2> $O/r2ad_rooms_${v}.err
  $B --workload lidar_maze > $O/r2ad_maze_${v}.json 2> $O/r2ad_maze_${v}.err
done
unset APG_B200_LIB
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2ad_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
PY
