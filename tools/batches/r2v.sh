#!/usr/bin/env bash
# round-2 batch v: table-driven beam filter, tinyloc issue/copy-out/visited changes
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_lidar_gpu.py tests/test_image_gpu.py -m gpu -x -q > $O/r2v_pytest.log 2>&1
tail -n 3 $O/r2v_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 4"
$B --workload tinyloc > $O/r2v_tiny.json 2> $O/r2v_tiny.err
for f in 0 1; do
  for wl in lidar_rooms lidar_maze; do
    APG_LIDAR_FILTER=$f $B --workload $wl > $O/r2v_${wl}_f$f.json 2> $O/r2v_${wl}_f$f.err
  done
done
python - <<PY
import json
for n in ["tiny","lidar_rooms_f0","lidar_rooms_f1","lidar_maze_f0","lidar_maze_f1"]:
    try:
        d=json.loads(open("$O/r2v_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(n, "FAILED", e)
PY
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:lidar_step -s 232 -c 1 -f -o $O/prof_r2v_lidar_rooms $B --workload lidar_rooms --prime-ms 0 > $O/r2v_ncu_lidar.log 2>&1
$NCU -k regex:tma_step -s 4 -c 1 -f -o $O/prof_r2v_tinyloc $B --workload tinyloc --prime-ms 0 > $O/r2v_ncu_tiny.log 2>&1
