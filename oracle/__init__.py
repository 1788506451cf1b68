"""CPU oracle package -- TEST INFRASTRUCTURE ONLY (see oracle/bc_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  The product (blume_capel_b200) never does.
"""
