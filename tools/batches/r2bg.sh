#!/usr/bin/env bash
# round-2 batch bg: LIDAR map staging as one bulk copy per environment from a pre-padded bank; MNIST host step with the
# copies on the step stream; LIDAR + image GPU tests
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline"
python -m pytest tests/test_lidar_gpu.py tests/test_stats_gpu.py -m gpu -q -x > $O/r2bg_pytest.log 2>&1
tail -n 3 $O/r2bg_pytest.log
for rep in 1 2; do
  $B --workload lidar_rooms --steps 20 --e2e-steps 8 > $O/r2bg_rooms_$rep.json 2> $O/r2bg_rooms_$rep.err
done
$B --workload lidar_maze --steps 20 --e2e-steps 8 > $O/r2bg_maze.json 2> $O/r2bg_maze.err
$B --workload lidar_rooms --lidar-maps procedural --steps 20 --e2e-steps 8 > $O/r2bg_rooms_proc.json 2> $O/r2bg_rooms_proc.err
$B --workload mnist --steps 20 --e2e-steps 200 > $O/r2bg_mnist.json 2> $O/r2bg_mnist.err
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed \
    --clock-control none -k regex:lidar_step -s 232 -c 1 --csv --log-file $O/r2bg_ncu_rooms.csv \
    $B --workload lidar_rooms --steps 20 --e2e-steps 2 --prime-ms 0 > /dev/null 2>&1
python - <<PY
import json,glob,csv
for f in sorted(glob.glob("$O/r2bg_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "obs %.3e"%d["e2e"]["with_obs"]["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("$O/r2bg_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f.split('/')[-1], [(r[-3],r[-1]) for r in rows[1:]])
PY
