#!/usr/bin/env bash
# The measurement commands behind profiles/r01_bench_*.json (run on a B200 box from the repo root).
#   tools/run_benches.sh 1        one GPU: the four workloads (+ CPU baseline) and the reference arm
#   tools/run_benches.sh N        N in 2,4,8: weak scaling of mnist / tinyloc / lidar_rooms under torchrun
# ncu captures (one per process, after the plain command has exited 0):
#   ncu --set full --clock-control none --import-source on -k regex:image_ -s 6 -c 1 -o prof python bench.py ...
#   python tools/ncu_summary.py prof.ncu-rep profiles/<name>.txt
set -u
N=${1:-1}
OUT=${2:-gpurun_out}
mkdir -p "$OUT"
if [ "$N" = "1" ]; then
  for wl in mnist tinyloc lidar_rooms lidar_maze; do
    python bench.py --workload $wl > "$OUT/bench_${wl}_1gpu.json" 2> "$OUT/bench_${wl}_1gpu.err"
    tail -n 1 "$OUT/bench_${wl}_1gpu.json" | head -c 240; echo
  done
  python bench.py --impl reference > "$OUT/bench_reference_arm_mnist.json" 2> "$OUT/bench_reference_arm.err"
else
  for wl in mnist tinyloc lidar_rooms; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus "$N" --workload $wl > "$OUT/bench_${wl}_${N}gpu.json" 2> "$OUT/bench_${wl}_${N}gpu.err"
    tail -n 1 "$OUT/bench_${wl}_${N}gpu.json" | head -c 240; echo
  done
fi
