"""Workload for an ncu launch list of the cluster move and the cluster-size measurement (development aid):
ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_cluster.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blume_capel_b200 as bc

e = bc.Engine(4096, bc.PERIODIC, 4, 0, 7)
e.set_params(0.608, 1.966, 1.0)
e.init_random()
e.sweep(50)
e.cluster_update(3)
e.cluster_moments(0)
e.cluster_moments(1, 1)
e.sync()
e.close()
