set -x; mkdir -p gpurun_out
nvidia-smi -L | head -8
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -c 600 gpurun_out/r2_bench_8gpu.json
grep -c "Init COMPLETE" gpurun_out/r2_bench_8gpu.err; grep "nranks 8" gpurun_out/r2_bench_8gpu.err | head -3; tail -3 gpurun_out/r2_bench_8gpu.err
