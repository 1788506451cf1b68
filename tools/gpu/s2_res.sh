mkdir -p gpurun_out; rm -f gpurun_out/s2_res.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_statistics.py tests/test_gpu_cli.py -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
for l in v5 default v5 default; do
  if [ "$l" = default ]; then f=blume_capel_b200/libbc_engine.so; else f=blume_capel_b200/libbc_engine_$l.so; fi
  BC_ENGINE_LIB=$PWD/$f python tools/tune_resident.py >> gpurun_out/s2_res.jsonl
done
