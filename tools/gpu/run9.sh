set -x; mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_gpu_tests_2gpu.log; tail -6 gpurun_out/r2_gpu_tests_2gpu.log
mkdir -p /tmp/dump && cd /tmp/dump && $GRAFT_REPO_ROOT/blume_capel_b200/host/bc_run 0 d 2 3 0.608 1.966 1.0 --L 4096 --burn-sweeps 30 --sweeps-per-sample 10 --mkdir --seed 5 --time-dumps 2> $GRAFT_REPO_ROOT/gpurun_out/r2_dump_timing_L4096.jsonl | tail -1; cd $GRAFT_REPO_ROOT
cat gpurun_out/r2_dump_timing_L4096.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -c 400 gpurun_out/r2_bench_2gpu.json; tail -2 gpurun_out/r2_bench_2gpu.err
