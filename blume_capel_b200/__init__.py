"""blume_capel_b200 -- host-side mirror of the B200 Blume-Capel Metropolis engine.

The product is the C-ABI shared library ``libbc_engine.so`` (include/bc_engine.h,
built from csrc/ for sm_100a).  This package is the thin ctypes binding used by the
tests, bench.py and the Python examples; the C++ command-line driver with the
reference's argv (``run root burn nsteps T D J``, /root/reference/main.cpp:316-324)
lives in host/.
"""
from .engine import Engine, BCError, Obs, load_library, library_path, slab_unique_id, sweep_group, pack_spins, unpack_spins, PERIODIC, OPEN  # noqa: F401
