set -x; mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/s2_gpu_tests.log; tail -5 gpurun_out/s2_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench_1gpu.json 2> gpurun_out/s2_bench_1gpu.err
tail -c 300 gpurun_out/s2_bench_1gpu.json; tail -2 gpurun_out/s2_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_fast -s 8 -c 4 -o gpurun_out/s2_prof_sweep -f python tools/prof_cmd.py 64 6 > gpurun_out/ncu_s.log 2>&1
tail -2 gpurun_out/ncu_s.log
