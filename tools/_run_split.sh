python -m pytest tests/test_lidar_gpu.py -m gpu -x -q 2>&1 | tail -2
for wl in lidar_rooms lidar_maze; do
python bench.py --workload $wl --no-cpu-baseline --e2e-steps 4 > gpurun_out/l_${wl}.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/l_${wl}.json').read().strip().splitlines()[-1]); print('$wl', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.4f'%d['roofline']['frac'])"
APG_B200_LIB=$PWD/active-perception-gym_b200/ap_gym_b200/_lib/libapg_b200_noexit.so python bench.py --workload $wl --no-cpu-baseline --e2e-steps 4 > gpurun_out/l_${wl}_ne.json 2>gpurun_out/sp.err
python -c "
import json; d=json.loads(open('gpurun_out/l_${wl}_ne.json').read().strip().splitlines()[-1]); print('$wl noexit', '%.3e'%d['value'], '%.1f us'%(d['ms_per_step']*1e3), '%.4f'%d['roofline']['frac'])"
done
