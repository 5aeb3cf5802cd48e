python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -x -q 2>&1 | tail -3
for wl in tinyloc mnist; do for b in float64 float32; do
python bench.py --workload $wl --blend $b --no-cpu-baseline --e2e-steps 4 > gpurun_out/c_${wl}_${b}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_${wl}_${b}.json').read().strip().splitlines()[-1]); print('$wl $b', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'])"
done; done
for cfg in "16 128" "16 256" "24 256" "32 512" "48 256"; do set -- $cfg
APG_IMG_THREADS=$2 APG_IMG_ENVS_PER_BLOCK=$1 python bench.py --workload tinyloc --no-cpu-baseline --e2e-steps 4 > gpurun_out/sp_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/sp_t.json').read().strip().splitlines()[-1]); print('tiny f64 E $1 thr $2', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3))"
done
