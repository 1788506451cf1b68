set -x; mkdir -p gpurun_out
python tools/prof_single.py > gpurun_out/s2_single.jsonl 2>&1; cat gpurun_out/s2_single.jsonl
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,launch__waves_per_multiprocessor --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/s2_ncu_single_4096.csv python tools/prof_single.py 4096 > /dev/null 2>&1
tail -5 gpurun_out/s2_ncu_single_4096.csv | cut -c1-300
