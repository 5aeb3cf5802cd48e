#!/usr/bin/env bash
# round-2 batch bj: LIDAR host step -- pipeline chunks x mapped result stores vs device-to-host copies
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --workload lidar_rooms --steps 20 --e2e-steps 40"
for nomap in 0 1; do
  for c in 1 2 4 8; do
    APG_HOST_CHUNKS=$c APG_HOST_NOMAP=$nomap $B > $O/r2bj_rooms_c${c}_nomap${nomap}.json 2> $O/r2bj_rooms_c${c}_nomap${nomap}.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2bj_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        n=d["config"]["num_envs_per_gpu"]
        print(f.split('/')[-1], "e2e %.3e (%.0f us/step)"%(d["e2e"]["value"], n/d["e2e"]["value"]*1e6), "obs %.3e (%.0f us)"%(d["e2e"]["with_obs"]["value"], n/d["e2e"]["with_obs"]["value"]*1e6), "dev", round(d["ms_per_step"]*1e3,1))
    except Exception as e:
        print(f, "FAILED", e)
PY
