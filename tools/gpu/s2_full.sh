set -x; mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/s2_gpu_tests.log; tail -15 gpurun_out/s2_gpu_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench_1gpu.json 2> gpurun_out/s2_bench_1gpu.err
tail -c 600 gpurun_out/s2_bench_1gpu.json; tail -2 gpurun_out/s2_bench_1gpu.err
rm -f gpurun_out/s2_tune_tie.jsonl
for l in base ""; do
  f=blume_capel_b200/libbc_engine${l:+_$l}.so
  BC_ENGINE_LIB=$PWD/$f python tools/tune_tie.py >> gpurun_out/s2_tune_tie.jsonl
done
