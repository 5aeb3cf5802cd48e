#!/usr/bin/env bash
# round-2 batch bh: procedural-dataset mode on the bulk map staging (generator writes padded slots)
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 8"
python -m pytest tests/test_lidar_gpu.py tests/test_maps.py -m gpu -q -x > $O/r2bh_pytest.log 2>&1
tail -n 3 $O/r2bh_pytest.log
$B --workload lidar_rooms --steps 20 > $O/r2bh_rooms.json 2> $O/r2bh_rooms.err
$B --workload lidar_rooms --lidar-maps procedural --steps 20 > $O/r2bh_rooms_proc.json 2> $O/r2bh_rooms_proc.err
$B --workload lidar_rooms --lidar-maps procedural --steps 202 > $O/r2bh_rooms_proc_k202.json 2> $O/r2bh_rooms_proc_k202.err
$B --workload lidar_maze --lidar-maps procedural --steps 20 > $O/r2bh_maze_proc.json 2> $O/r2bh_maze_proc.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2bh_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
PY
