set -x; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_group.py -q -m gpu 2>&1 | tail -12 > gpurun_out/s2_tie_parity.log; tail -12 gpurun_out/s2_tie_parity.log
rm -f gpurun_out/s2_tune_tie.jsonl
for l in base "" pairdefer; do
  f=blume_capel_b200/libbc_engine${l:+_$l}.so
  BC_ENGINE_LIB=$PWD/$f python tools/tune_tie.py >> gpurun_out/s2_tune_tie.jsonl
done
