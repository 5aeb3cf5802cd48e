#!/usr/bin/env bash
# round-2 batch ba: footprint transport of the TinyImageNetLoc kernel (tensor-map box / bulk copy per row / cp.async pieces)
# x L2 fetch granularity hint (default / 32 bytes): time per step and DRAM traffic of one 20-step window
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 4"
for v in default t_rowcopy t_ldgsts; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  for l2 in 0 32; do
    if [ "$l2" = "0" ]; then unset APG_L2_FETCH; else export APG_L2_FETCH=$l2; fi
    for rep in 1 2; do
      $B --workload tinyloc --steps 200 > $O/r2ba_tiny_${v}_l2f${l2}_$rep.json 2> $O/r2ba_tiny_${v}_l2f${l2}_$rep.err
    done
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum \
        --clock-control none -k regex:tma_step -s 4 -c 1 --csv --log-file $O/r2ba_ncu_${v}_l2f${l2}.csv \
        $B --workload tinyloc --steps 20 --prime-ms 0 > /dev/null 2>&1
  done
  unset APG_L2_FETCH
  python -m pytest tests/test_image_gpu.py -m gpu -q -x -k "tiny or tma or loc" > $O/r2ba_pytest_$v.log 2>&1
  tail -n 2 $O/r2ba_pytest_$v.log
done
unset APG_B200_LIB
export APG_L2_FETCH=32
$B --workload mnist --steps 200 > $O/r2ba_mnist_l2f32.json 2> $O/r2ba_mnist_l2f32.err
unset APG_L2_FETCH
$B --workload mnist --steps 200 > $O/r2ba_mnist_l2f0.json 2> $O/r2ba_mnist_l2f0.err
grep -h "FetchGranularity" $O/r2ba_*.err | sort | uniq -c
python - <<PY
import json,glob,csv
for f in sorted(glob.glob("$O/r2ba_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("$O/r2ba_ncu_*.csv")):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    print(f.split('/')[-1], [(r[-3],r[-2],r[-1]) for r in rows[1:]])
PY
