set -x; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_multi_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2_slab_tests.log; tail -25 gpurun_out/r2_slab_tests.log
timeout 300 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_gap.py -q -m gpu 2>&1 | tail -3
