#!/usr/bin/env bash
# round-2 batch z: rooms generator on the device; TMA kernel with early first-pass copies + input prefetch
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r2z_pytest.log 2>&1
tail -n 4 $O/r2z_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 16"
for i in 1 2; do
$B --workload tinyloc > $O/r2z_tiny$i.json 2> $O/r2z_tiny$i.err
done
$B --workload tinyloc --steps 200 > $O/r2z_tiny_k200.json 2> $O/r2z_tiny_k200.err
python - <<PY
import json
for n in ["tiny1","tiny2","tiny_k200"]:
    try:
        d=json.loads(open("$O/r2z_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["ms_per_step"]*1e3,2), "us/step value %.3e"%d["value"], "frac", round(d["roofline"]["frac"],3), "e2e %.3e"%d["e2e"]["value"], "flags", d.get("error_flags"))
    except Exception as e:
        print(n, "FAILED", e)
PY
