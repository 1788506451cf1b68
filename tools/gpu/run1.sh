set -x; mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_slab_multi_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2_slab_tests.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "halo or headline or slab or pair_table or graph" 2>&1 | tail -15 > gpurun_out/r2_parity_subset.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -5 gpurun_out/r2_slab_tests.log gpurun_out/r2_parity_subset.log
tail -c 3000 gpurun_out/r2_bench_2gpu.json
tail -5 gpurun_out/r2_bench_2gpu.err
