python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -x -q 2>&1 | tail -3
for sp in 1 0; do for wl in tinyloc mnist; do for b in float64 float32; do
APG_IMG_SPEC=$sp python bench.py --workload $wl --blend $b --no-cpu-baseline --e2e-steps 4 > gpurun_out/c_${wl}_${b}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_${wl}_${b}.json').read().strip().splitlines()[-1]); print('spec $sp $wl $b', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'])"
done; done; done
