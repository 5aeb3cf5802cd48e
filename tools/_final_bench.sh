set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in mnist tinyloc lidar_rooms lidar_maze; do
python bench.py --workload $wl > gpurun_out/final_${wl}_1gpu.json 2>gpurun_out/final_${wl}.err; tail -c 600 gpurun_out/final_${wl}_1gpu.json; echo
done
python bench.py --impl reference > gpurun_out/final_reference_mnist.json 2>gpurun_out/final_ref.err; tail -c 400 gpurun_out/final_reference_mnist.json
python bench.py > gpurun_out/final_default.json 2>gpurun_out/final_default.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
