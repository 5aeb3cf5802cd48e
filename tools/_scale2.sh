for wl in mnist tinyloc lidar_rooms; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl > gpurun_out/final_${wl}_2gpu.json 2>gpurun_out/final_${wl}_2gpu.err; tail -n 1 gpurun_out/final_${wl}_2gpu.json | head -c 200; echo
done
