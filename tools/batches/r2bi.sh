#!/usr/bin/env bash
# round-2 batch bi: conflict-free lane mapping of the footprint cp.async pieces (TinyImageNetLoc kernel)
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 4"
python -m pytest tests/test_image_gpu.py -m gpu -q -x -k "tiny or tma or loc" > $O/r2bi_pytest.log 2>&1
tail -n 2 $O/r2bi_pytest.log
for rep in 1 2; do
  $B --workload tinyloc --steps 200 > $O/r2bi_tiny_k200_$rep.json 2> $O/r2bi_tiny_k200_$rep.err
  $B --workload tinyloc --steps 20 > $O/r2bi_tiny_k20_$rep.json 2> $O/r2bi_tiny_k20_$rep.err
done
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed.avg.per_cycle_elapsed \
    --clock-control none -k regex:tma_step -s 4 -c 1 --csv --log-file $O/r2bi_ncu_tiny.csv \
    $B --workload tinyloc --steps 20 --prime-ms 0 > /dev/null 2>&1
python - <<PY
import json,glob,csv
for f in sorted(glob.glob("$O/r2bi_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("$O/r2bi_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f.split('/')[-1], [(r[-3],r[-1]) for r in rows[1:]])
PY
