#!/usr/bin/env bash
# round-2 batch bo: LIDAR walk with the crossing parameters recomputed from the cell (no loop-carried float64 state)
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_lidar_gpu.py -m gpu -q -x > $O/r2bo_pytest.log 2>&1
tail -n 2 $O/r2bo_pytest.log
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 4 --steps 20"
$B --workload lidar_rooms > $O/r2bo_rooms.json 2> $O/r2bo_rooms.err
$B --workload lidar_maze > $O/r2bo_maze.json 2> $O/r2bo_maze.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2bo_*.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
PY
