python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for wl in mnist tinyloc; do
python bench.py --workload $wl > gpurun_out/final_${wl}_1gpu.json 2>gpurun_out/final_${wl}.err
python -c "
import json; d=json.loads(open('gpurun_out/final_${wl}_1gpu.json').read().strip().splitlines()[-1]); print('$wl', '%.3e'%d['value'], '%.2f us'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], d['gpu_launches'], d.get('alt_blend',{}).get('value'))"
done
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-alt-blend --e2e-steps 3 > gpurun_out/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-alt-blend --e2e-steps 3 > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log | head -c 300
