"""Short, fixed command for ncu: the bench workload (64 x L=4096, int8, periodic) for a few sweeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
e = bc.Engine(4096, bc.PERIODIC, n_rep, 0, 0x5EED0002)
e.set_params(0.608, 1.966, 1.0)
e.init_random()
e.sweep(sweeps, 3, fetch=False)
e.sync()
print("ms", e.last_sweep_ms(), "launches", e.launch_count)
e.close()
