set -x; mkdir -p gpurun_out
python -m pytest tests/test_gpu_cluster.py tests/test_gpu_gap.py tests/test_magnet.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_cluster_tests.log
tail -6 gpurun_out/r2_cluster_tests.log
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "packed or halo" 2>&1 | tail -8
python tools/time_cluster.py > gpurun_out/r2_time_cluster.jsonl 2>&1
cat gpurun_out/r2_time_cluster.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_cluster.csv python tools/prof_cluster.py > gpurun_out/ncu_cl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cl_ -s 6 -c 6 -o gpurun_out/r2_prof_cluster -f python tools/prof_cluster.py > gpurun_out/ncu_cl2.log 2>&1
tail -2 gpurun_out/ncu_cl2.log
