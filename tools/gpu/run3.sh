set -x; mkdir -p gpurun_out
python tools/gpu/debug_cluster.py 512 > gpurun_out/r2_debug_cluster.log 2>&1
tail -20 gpurun_out/r2_debug_cluster.log
python -m pytest tests/test_gpu_cluster.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_cluster_tests.log
tail -8 gpurun_out/r2_cluster_tests.log
python -m pytest tests/test_gpu_gap.py tests/test_magnet.py tests/test_gpu_cli.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_gap_tests.log
tail -30 gpurun_out/r2_gap_tests.log
