timeout 600 python -m pytest tests/test_rng_gpu.py tests/test_image_gpu.py -m gpu -x -q 2>&1 | tail -3
for wl in mnist tinyloc; do
python bench.py --workload $wl --no-cpu-baseline > gpurun_out/final_${wl}_1gpu.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/final_${wl}_1gpu.json').read().strip().splitlines()[-1]); print('$wl', '%.3e'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], d['gpu_launches'], d.get('alt_blend',{}).get('value'))"
done
