#!/usr/bin/env bash
# round-2 final measurements on ONE GPU: GPU test suite, the driver's bench command (both arms), procedural-maps lines,
# ncu captures (launch list + one --set full capture per workload) -- each ncu run only after the plain command exited 0
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1
tail -n 3 $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
echo "bench rc $?"; tail -c 400 $O/r02_bench_1gpu.json; echo
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 8"
: > $O/r02_lidar_procedural.jsonl
$B --workload lidar_rooms --lidar-maps procedural --steps 20 >> $O/r02_lidar_procedural.jsonl 2>> $O/r02_lidar_procedural.err
$B --workload lidar_rooms --lidar-maps procedural --steps 202 >> $O/r02_lidar_procedural.jsonl 2>> $O/r02_lidar_procedural.err
$B --workload lidar_rooms --lidar-maps bank --steps 202 >> $O/r02_lidar_procedural.jsonl 2>> $O/r02_lidar_procedural.err
$B --workload lidar_maze --lidar-maps procedural --steps 20 >> $O/r02_lidar_procedural.jsonl 2>> $O/r02_lidar_procedural.err
: > $O/r02_long_windows.jsonl
$B --workload mnist --steps 200 >> $O/r02_long_windows.jsonl 2>> $O/r02_long_windows.err
$B --workload tinyloc --steps 200 >> $O/r02_long_windows.jsonl 2>> $O/r02_long_windows.err
B2="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 --prime-ms 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_mnist.csv $B2 --workload mnist > $O/r02_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:image_rollout -s 4 -c 1 -f -o $O/prof_r02_mnist $B2 --workload mnist > $O/r02_ncu_mnist.log 2>&1
$NCU -k regex:tma_step -s 4 -c 1 -f -o $O/prof_r02_tinyloc $B2 --workload tinyloc > $O/r02_ncu_tinyloc.log 2>&1
$NCU -k regex:lidar_step -s 232 -c 1 -f -o $O/prof_r02_lidar_rooms $B2 --workload lidar_rooms > $O/r02_ncu_lidar_rooms.log 2>&1
$NCU -k regex:lidar_step -s 232 -c 1 -f -o $O/prof_r02_lidar_maze $B2 --workload lidar_maze > $O/r02_ncu_lidar_maze.log 2>&1
ls -la $O | grep r02
