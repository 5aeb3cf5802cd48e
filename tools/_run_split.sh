python -m pytest tests/test_image_gpu.py -m gpu -x -q 2>&1 | tail -3
for sp in 0 1; do for b in float64 float32; do
APG_IMG_SPLIT=$sp python bench.py --workload tinyloc --blend $b --no-cpu-baseline --e2e-steps 4 > gpurun_out/sp_${sp}_${b}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/sp_${sp}_${b}.json').read().strip().splitlines()[-1]); print('split $sp $b', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'])"
done; done
for t in 128 512; do for E in 16 32; do
APG_IMG_THREADS=$t APG_IMG_ENVS_PER_BLOCK=$E python bench.py --workload tinyloc --blend float32 --no-cpu-baseline --e2e-steps 4 > gpurun_out/sp_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/sp_t.json').read().strip().splitlines()[-1]); print('thr $t E $E f32', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3))"
done; done
