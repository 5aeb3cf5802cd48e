python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in 1 0; do for b in float64 float32; do
APG_IMG_TB_CONSEC=$c python bench.py --workload mnist --blend $b --no-cpu-baseline --e2e-steps 4 > gpurun_out/c_mnist_${b}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_mnist_${b}.json').read().strip().splitlines()[-1]); print('mnist consec $c $b', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'])"
done; done
python bench.py --workload tinyloc --no-cpu-baseline --e2e-steps 4 > gpurun_out/c_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_t.json').read().strip().splitlines()[-1]); print('tinyloc f64', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'])"
