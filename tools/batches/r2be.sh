#!/usr/bin/env bash
# round-2 batch be: MNIST rollout kernel with full-warp specialisation (unrolled cooperative copies, running pointers,
# host-side time_obs); image GPU tests; instruction count
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 4"
python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -q -x > $O/r2be_pytest.log 2>&1
tail -n 3 $O/r2be_pytest.log
for rep in 1 2; do
  $B --workload mnist --steps 200 > $O/r2be_mnist_k200_$rep.json 2> $O/r2be_mnist_k200_$rep.err
  $B --workload mnist --steps 20 > $O/r2be_mnist_k20_$rep.json 2> $O/r2be_mnist_k20_$rep.err
done
$B --workload tinyloc --steps 20 > $O/r2be_tiny_k20.json 2> $O/r2be_tiny_k20.err
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:image_rollout -s 4 -c 1 --csv --log-file $O/r2be_ncu_mnist.csv \
    $B --workload mnist --steps 20 --prime-ms 0 > /dev/null 2>&1
python - <<PY
import json,glob,csv
for f in sorted(glob.glob("$O/r2be_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("$O/r2be_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f.split('/')[-1], [(r[-3],r[-1]) for r in rows[1:]])
PY
