#!/usr/bin/env bash
# round-2 batch bk: e2e with a pool of distinct pinned host inputs; LIDAR host-step chunks under that workload
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --steps 20 --e2e-steps 40"
for c in 1 2 4; do
  APG_HOST_CHUNKS=$c $B --workload lidar_rooms > $O/r2bk_rooms_c${c}.json 2> $O/r2bk_rooms_c${c}.err
done
for c in 2 4; do
  APG_HOST_CHUNKS=$c $B --workload lidar_maze > $O/r2bk_maze_c${c}.json 2> $O/r2bk_maze_c${c}.err
done
$B --workload mnist > $O/r2bk_mnist.json 2> $O/r2bk_mnist.err
$B --workload tinyloc > $O/r2bk_tiny.json 2> $O/r2bk_tiny.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2bk_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        n=d["config"]["num_envs_per_gpu"]
        print(f.split('/')[-1], "e2e %.3e (%.0f us/step)"%(d["e2e"]["value"], n/d["e2e"]["value"]*1e6), "obs %.3e (%.0f us)"%(d["e2e"]["with_obs"]["value"], n/d["e2e"]["with_obs"]["value"]*1e6), "dev", round(d["ms_per_step"]*1e3,1))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 $O/r2bk_*.err | tail -n 12
