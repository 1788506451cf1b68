set -x; mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_cluster.py tests/test_magnet.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_cluster_tests.log
tail -5 gpurun_out/r2_cluster_tests.log
python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_cluster.py 2>&1 | tail -25 > gpurun_out/r2_gpu_tests.log
tail -5 gpurun_out/r2_gpu_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 2500 gpurun_out/r2_bench_1gpu.json
tail -3 gpurun_out/r2_bench_1gpu.err
