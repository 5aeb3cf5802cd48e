timeout 600 python -m pytest tests/test_rng_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/e2e_breakdown.py 2>&1 | head -3
for wl in mnist tinyloc; do
python bench.py --workload $wl --no-cpu-baseline --no-alt-blend > gpurun_out/c_${wl}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_${wl}.json').read().strip().splitlines()[-1]); print('$wl', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), 'e2e %.3e'%d['e2e']['value'], d['gpu_launches'])"
done
