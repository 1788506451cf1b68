set -x; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -q -m gpu -x 2>&1 | tail -3
for l in v3 default v3 default; do
  if [ "$l" = default ]; then f=blume_capel_b200/libbc_engine.so; else f=blume_capel_b200/libbc_engine_$l.so; fi
  echo "== $l"; BC_ENGINE_LIB=$PWD/$f python tools/prof_single.py 4096 2048 8192 1024
done > gpurun_out/s2_bar_single.log 2>&1
cat gpurun_out/s2_bar_single.log
LIBS="v3 default" bash tools/gpu/s2_tune.sh > /dev/null 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s2_bar_bench.json 2>/dev/null
