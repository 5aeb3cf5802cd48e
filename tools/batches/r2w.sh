#!/usr/bin/env bash
# round-2 batch w: payload draws on the rng stream, MNIST action prefetch + unrolled logits, L2 persistence experiment
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2w_pytest.log 2>&1
tail -n 3 $O/r2w_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 8"
$B --workload mnist > $O/r2w_mnist.json 2> $O/r2w_mnist.err
$B --workload tinyloc > $O/r2w_tiny.json 2> $O/r2w_tiny.err
APG_IMG_L2_PERSIST=64 $B --workload tinyloc > $O/r2w_tiny_p64.json 2> $O/r2w_tiny_p64.err
APG_IMG_L2_PERSIST=32 $B --workload tinyloc > $O/r2w_tiny_p32.json 2> $O/r2w_tiny_p32.err
$B --workload mnist --steps 200 > $O/r2w_mnist_k200.json 2> $O/r2w_mnist_k200.err
python - <<PY
import json
for n in ["mnist","mnist_k200","tiny","tiny_p64","tiny_p32"]:
    try:
        d=json.loads(open("$O/r2w_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "e2e %.3e"%d["e2e"]["value"], "flags", d.get("error_flags"), "launches", d.get("gpu_launches"))
    except Exception as e:
        print(n, "FAILED", e)
PY
grep -h "apg: persisting" $O/r2w_tiny_p64.err | head -2
