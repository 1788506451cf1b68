set -x; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/s2_gpu_tests.log; tail -6 gpurun_out/s2_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
python tools/time_small.py > gpurun_out/s2_time_small.jsonl 2>&1; tail -4 gpurun_out/s2_time_small.jsonl | cut -c1-300
