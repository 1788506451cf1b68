set -x; mkdir -p gpurun_out
python -m pytest tests/test_gpu_cluster.py tests/test_gpu_gap.py -q -m gpu 2>&1 | tail -6
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "packed or halo" 2>&1 | tail -8
python tools/time_cluster.py > gpurun_out/r2_time_cluster.jsonl 2>&1
cat gpurun_out/r2_time_cluster.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_cluster.csv python tools/prof_cluster.py > gpurun_out/ncu_cl.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
tail -c 1500 gpurun_out/r2_bench_1gpu.json; tail -3 gpurun_out/r2_bench_1gpu.err
