#!/usr/bin/env bash
# round-2 batch bp: ncu --set full captures of the final LIDAR kernel (rooms, maze)
set -u
O=gpurun_out
B2="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 --prime-ms 0"
NCU="ncu --set full --clock-control none --import-source on"
timeout 100 $NCU -k regex:lidar_step -s 232 -c 1 -f -o $O/prof_r02_lidar_rooms $B2 --workload lidar_rooms > $O/r02_ncu_lidar_rooms.log 2>&1
timeout 100 $NCU -k regex:lidar_step -s 232 -c 1 -f -o $O/prof_r02_lidar_maze $B2 --workload lidar_maze > $O/r02_ncu_lidar_maze.log 2>&1
ls -la $O/prof_r02_lidar_*.ncu-rep
