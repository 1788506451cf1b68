set -x; mkdir -p gpurun_out
python -m pytest tests/test_gpu_cluster.py tests/test_gpu_gap.py -q -m gpu 2>&1 | tail -4
python tools/tune_waves.py > gpurun_out/r2_tune_waves.jsonl 2>&1; cat gpurun_out/r2_tune_waves.jsonl
python tools/time_cluster.py > gpurun_out/r2_time_cluster.jsonl 2>&1; cat gpurun_out/r2_time_cluster.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_cluster.csv python tools/prof_cluster.py > gpurun_out/ncu_cl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cl_ -s 6 -c 3 -o gpurun_out/r2_prof_cluster -f python tools/prof_cluster.py > gpurun_out/ncu_cl2.log 2>&1
