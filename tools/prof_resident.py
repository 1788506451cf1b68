"""Workload for an ncu capture of the shared-memory-resident kernel: one L = 64 lattice, 2000 sweeps (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_rep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
e = bc.Engine(L, bc.PERIODIC, n_rep, 0, 7)
e.set_params(0.608, 1.966, 1.0)
e.init_random()
e.sweep(2000, 0)
e.sync()
e.close()
