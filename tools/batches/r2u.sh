#!/usr/bin/env bash
# round-2 batch u: parity of the new default build, tinyloc kernel variants, LIDAR beam filter on/off, ncu captures
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r2u_pytest.log 2>&1
tail -n 3 $O/r2u_pytest.log
B="python bench.py --only --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 4"
for v in default t_head t_old t_f64x16 t_f32x32 t_f64_rs t_f32_rs_x1; do
  if [ "$v" = "default" ]; then unset APG_B200_LIB; else export APG_B200_LIB=$PWD/build/variants/libapg_$v.so; fi
  $B --workload tinyloc > $O/r2u_tiny_$v.json 2> $O/r2u_tiny_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2u_tiny_$v.json").read().strip().splitlines()[-1])
    print("tinyloc $v", round(d["ms_per_step"]*1e3,2), "us/step frac", round(d["roofline"]["frac"],3), "flags", d.get("error_flags"))
except Exception as e:
    print("tinyloc $v FAILED", e)
PY
done
export APG_B200_LIB=$PWD/build/variants/libapg_t_f32_rs_x1.so
python -m pytest tests/test_image_gpu.py -m gpu -x -q -k "tma or tiny or loc" > $O/r2u_pytest_x1.log 2>&1
tail -n 2 $O/r2u_pytest_x1.log
unset APG_B200_LIB
for f in 0 1; do
  for wl in lidar_rooms lidar_maze; do
    APG_LIDAR_FILTER=$f $B --workload $wl > $O/r2u_${wl}_f$f.json 2> $O/r2u_${wl}_f$f.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/r2u_${wl}_f$f.json").read().strip().splitlines()[-1])
    print("$wl filter $f", round(d["ms_per_step"]*1e3,1), "us/step value %.3e" % d["value"], "flags", d.get("error_flags"))
except Exception as e:
    print("$wl filter $f FAILED", e)
PY
  done
done
$B --workload mnist > $O/r2u_mnist.json 2> $O/r2u_mnist.err
tail -c 300 $O/r2u_mnist.json; echo
# ncu captures (after the plain commands above have exited): mnist rollout window, lidar rooms step (filter auto)
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:image_rollout -s 2 -c 3 -f -o $O/prof_r2u_mnist $B --workload mnist --prime-ms 0 > $O/r2u_ncu_mnist.log 2>&1
$NCU -k regex:lidar_step -s 232 -c 1 -f -o $O/prof_r2u_lidar_rooms $B --workload lidar_rooms --prime-ms 0 > $O/r2u_ncu_lidar.log 2>&1
$NCU -k regex:tma_step -s 2 -c 3 -f -o $O/prof_r2u_tinyloc $B --workload tinyloc --prime-ms 0 > $O/r2u_ncu_tiny.log 2>&1
ls -la $O | tail -n 8
