for wl in mnist tinyloc lidar_rooms; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $wl --e2e-steps 17 > gpurun_out/final_${wl}_8gpu.json 2>gpurun_out/final_${wl}_8gpu.err; tail -n 1 gpurun_out/final_${wl}_8gpu.json | head -c 200; echo
done
