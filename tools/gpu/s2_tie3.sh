set -x; mkdir -p gpurun_out
rm -f gpurun_out/s2_tune_tie.jsonl
for l in base nofix nofix7; do
  f=blume_capel_b200/libbc_engine${l:+_$l}.so
  BC_ENGINE_LIB=$PWD/$f python tools/tune_tie.py >> gpurun_out/s2_tune_tie.jsonl
done
