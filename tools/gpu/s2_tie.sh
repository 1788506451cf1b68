set -x; mkdir -p gpurun_out
python tools/gpu/debug_tie.py > gpurun_out/s2_debug_tie.log 2>&1; cat gpurun_out/s2_debug_tie.log | cut -c1-400
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -m gpu 2>&1 | tail -12 > gpurun_out/s2_tie_parity.log; tail -12 gpurun_out/s2_tie_parity.log
rm -f gpurun_out/s2_tune_tie.jsonl
for l in base "" swp7; do
  f=blume_capel_b200/libbc_engine${l:+_$l}.so
  BC_ENGINE_LIB=$PWD/$f python tools/tune_tie.py >> gpurun_out/s2_tune_tie.jsonl
done
