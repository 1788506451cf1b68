set -x; mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_slab_multi_gpu.py -q -m gpu 2>&1 | tail -8 > gpurun_out/s2_slab_tests.log; tail -8 gpurun_out/s2_slab_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s2_bench_2gpu.json 2> gpurun_out/s2_bench_2gpu.err
tail -c 2500 gpurun_out/s2_bench_2gpu.json
tail -3 gpurun_out/s2_bench_2gpu.err
