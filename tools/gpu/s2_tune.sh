set -x; mkdir -p gpurun_out
rm -f gpurun_out/s2_tune_tie.jsonl
for l in $LIBS; do
  if [ "$l" = default ]; then f=blume_capel_b200/libbc_engine.so; else f=blume_capel_b200/libbc_engine_$l.so; fi
  BC_ENGINE_LIB=$PWD/$f python tools/tune_tie.py >> gpurun_out/s2_tune_tie.jsonl
done
