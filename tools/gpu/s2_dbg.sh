python tools/gpu/debug_tie.py > gpurun_out/s2_debug_tie.log 2>&1; cat gpurun_out/s2_debug_tie.log | cut -c1-900
