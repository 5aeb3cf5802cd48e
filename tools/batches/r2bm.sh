#!/usr/bin/env bash
# round-2 batch bm: payload draw kernels in 64-thread blocks (resident next to the persistent window kernels)
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --only --warmup 5 --no-cpu-baseline --e2e-steps 4"
python -m pytest tests/test_rng_gpu.py -m gpu -q -x > $O/r2bm_pytest.log 2>&1
tail -n 2 $O/r2bm_pytest.log
for rep in 1 2; do
  $B --workload tinyloc --steps 20 > $O/r2bm_tiny_k20_$rep.json 2> $O/r2bm_tiny_k20_$rep.err
  $B --workload tinyloc --steps 200 > $O/r2bm_tiny_k200_$rep.json 2> $O/r2bm_tiny_k200_$rep.err
  $B --workload mnist --steps 20 > $O/r2bm_mnist_k20_$rep.json 2> $O/r2bm_mnist_k20_$rep.err
  $B --workload mnist --steps 200 > $O/r2bm_mnist_k200_$rep.json 2> $O/r2bm_mnist_k200_$rep.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/r2bm_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(f, "FAILED", e)
PY
