#!/usr/bin/env bash
# round-2 batch y: TMA kernel with bulk-store copy-out (A/B against the LDS/STG copy-out)
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -q -x > $O/r2y_pytest.log 2>&1
tail -n 3 $O/r2y_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 16"
$B --workload tinyloc > $O/r2y_tiny.json 2> $O/r2y_tiny.err
APG_B200_LIB=$PWD/build/variants/libapg_t_nobulk.so $B --workload tinyloc > $O/r2y_tiny_nobulk.json 2> $O/r2y_tiny_nobulk.err
$B --workload tinyloc --steps 200 > $O/r2y_tiny_k200.json 2> $O/r2y_tiny_k200.err
python - <<PY
import json
for n in ["tiny","tiny_nobulk","tiny_k200"]:
    try:
        d=json.loads(open("$O/r2y_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "frac", round(d["roofline"]["frac"],3), "e2e %.3e"%d["e2e"]["value"], "obs %.3e"%d["e2e"]["with_obs"]["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(n, "FAILED", e)
PY
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:tma_step -s 4 -c 1 -f -o $O/prof_r2y_tinyloc $B --workload tinyloc --prime-ms 0 > $O/r2y_ncu_tiny.log 2>&1
