#!/usr/bin/env bash
# round-2 final multi-GPU line: the driver's command on N GPUs of one box (torchrun, one rank per GPU)
#   gpurun --gpus N -- bash tools/batches/r2_final_multi.sh N
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_${N}gpu.json 2> $O/r02_bench_${N}gpu.err
echo "rc $?"
python - <<PY
import json
d=json.loads(open("$O/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
print(d["n_gpus"], "headline %.3e"%d["value"], round(d["ms_per_step"]*1e3,2), "us/step", {k:("%.3e"%w["value"], round(w["ms_per_step"]*1e3,2)) for k,w in d["workloads"].items()}, "flags", d.get("error_flags"))
PY
tail -n 5 $O/r02_bench_${N}gpu.err
