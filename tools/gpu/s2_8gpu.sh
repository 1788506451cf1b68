set -x; mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s2_bench_8gpu.json 2> gpurun_out/s2_bench_8gpu.err
tail -c 1500 gpurun_out/s2_bench_8gpu.json
tail -3 gpurun_out/s2_bench_8gpu.err
