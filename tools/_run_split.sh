python -m pytest tests/test_image_gpu.py tests/test_stats_gpu.py -m gpu -x -q 2>&1 | tail -3
for sp in 2 1; do for b in float64 float32; do
APG_IMG_SPLIT=$sp python bench.py --workload tinyloc --blend $b --no-cpu-baseline --no-alt-blend --e2e-steps 4 > gpurun_out/c_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/c_t.json').read().strip().splitlines()[-1]); print('tinyloc split $sp $b', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.3f'%d['roofline']['frac'], d['gpu_launches'])"
done; done
for cfg in "16 128" "32 128" "64 256" "48 256"; do set -- $cfg
APG_IMG_THREADS=$2 APG_IMG_ENVS_PER_BLOCK=$1 python bench.py --workload tinyloc --blend float32 --no-cpu-baseline --e2e-steps 4 > gpurun_out/sp_t.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/sp_t.json').read().strip().splitlines()[-1]); print('tiny f32 split2 E $1 thr $2', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3))"
done
