python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -x -q 2>&1 | tail -3
for wl in tinyloc mnist; do for b in float64 float32; do
python bench.py --workload $wl --blend $b --no-cpu-baseline --e2e-steps 4 > gpurun_out/c_${wl}_${b}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_${wl}_${b}.json').read().strip().splitlines()[-1]); print('$wl $b', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'])"
done; done
for cfg in "64 3" "64 4" "32 4" "96 2"; do set -- $cfg
APG_IMG_ENVS_PER_BLOCK=$1 APG_IMG_TB=$2 python bench.py --workload mnist --no-cpu-baseline --e2e-steps 4 > gpurun_out/sp_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/sp_t.json').read().strip().splitlines()[-1]); print('mnist E $1 tb $2', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3))"
done
APG_IMG_THREADS=128 APG_IMG_ENVS_PER_BLOCK=16 python bench.py --workload tinyloc --no-cpu-baseline --e2e-steps 4 > gpurun_out/sp_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/sp_t.json').read().strip().splitlines()[-1]); print('tiny E16 t128 f64', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3))"
