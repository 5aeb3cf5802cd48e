#!/usr/bin/env bash
# round-2 batch x: full GPU suite; MNIST (action prefetch, unrolled logits, rng stream); LIDAR staging / reset-mask changes
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r2x_pytest.log 2>&1
tail -n 5 $O/r2x_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 16"
$B --workload mnist > $O/r2x_mnist.json 2> $O/r2x_mnist.err
$B --workload mnist --steps 200 > $O/r2x_mnist_k200.json 2> $O/r2x_mnist_k200.err
APG_HOST_CHUNKS=2 $B --workload mnist > $O/r2x_mnist_c2.json 2> $O/r2x_mnist_c2.err
APG_HOST_CHUNKS=4 $B --workload mnist > $O/r2x_mnist_c4.json 2> $O/r2x_mnist_c4.err
$B --workload lidar_rooms > $O/r2x_lidar_rooms.json 2> $O/r2x_lidar_rooms.err
$B --workload lidar_maze > $O/r2x_lidar_maze.json 2> $O/r2x_lidar_maze.err
python - <<PY
import json
for n in ["mnist","mnist_k200","mnist_c2","mnist_c4","lidar_rooms","lidar_maze"]:
    try:
        d=json.loads(open("$O/r2x_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "obs %.3e"%d["e2e"]["with_obs"]["value"], "flags", d.get("error_flags"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(n, "FAILED", e)
PY
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:image_rollout -s 4 -c 1 -f -o $O/prof_r2x_mnist $B --workload mnist --prime-ms 0 > $O/r2x_ncu_mnist.log 2>&1
