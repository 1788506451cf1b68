set -x; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/s2_bench_4gpu.json 2> gpurun_out/s2_bench_4gpu.err
tail -c 600 gpurun_out/s2_bench_4gpu.json
tail -3 gpurun_out/s2_bench_4gpu.err
