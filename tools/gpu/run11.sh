set -x; mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_cmd.py > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_memcheck.log
tail -5 gpurun_out/r2_memcheck.log
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "65536" 2>&1 | tail -5
