"""ctypes binding of oracle/libbc_oracle.so (the plain-C restatement in bc_oracle.c).

TEST INFRASTRUCTURE ONLY -- the checker, never the thing measured or shipped.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libbc_oracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
OBS_DTYPE = np.dtype([("sweep", "<i8"), ("M", "<i8"), ("Q", "<i8"), ("B", "<i8")])
PERIODIC, OPEN = 0, 1

_lib = None


def build(quiet: bool = True):
    """(Re)build libbc_oracle.so, and oracle/_ref when /root/reference exists."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        l = C.CDLL(LIB)
        l.bc_oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.bc_oracle_build_table.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p]
        l.bc_oracle_reference_de.restype = C.c_double
        l.bc_oracle_reference_de.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
        l.bc_oracle_init_random.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_void_p]
        l.bc_oracle_observables.argtypes = [C.c_int, C.c_int, C.c_void_p] + [C.POINTER(C.c_int64)] * 3
        l.bc_oracle_sweep.restype = C.c_int64
        l.bc_oracle_sweep.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
        l.bc_oracle_half_sweep_rows.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                                C.c_int64, C.c_int, C.c_int, C.c_int]
        l.bc_oracle_tie_count.restype = C.c_int64
        l.bc_oracle_tie_count.argtypes = [C.c_int]
        l.bc_oracle_enumerate.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        l.bc_oracle_reference_metropolis.restype = C.c_int64
        l.bc_oracle_reference_metropolis.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                                     C.c_int64, C.c_uint64]
        l.bc_oracle_bond_threshold.restype = C.c_uint32
        l.bc_oracle_bond_threshold.argtypes = [C.c_double, C.c_double]
        l.bc_oracle_cluster_update.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int64]
        l.bc_oracle_cluster_moments.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32,
                                                C.c_int64, C.c_void_p]
        l.bc_oracle_cluster_labels.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32,
                                               C.c_int64, C.c_void_p]
        l.bc_oracle_gap_histogram.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32,
                                              C.c_int64, C.c_int, C.c_void_p]
        _lib = l
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, np.uint32)
    k = np.asarray(key, np.uint32)
    out = np.zeros(4, np.uint32)
    lib().bc_oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def build_table(T, D, J=1.0):
    t = np.zeros(54, np.uint32)
    lib().bc_oracle_build_table(T, D, J, t.ctypes.data)
    return t


def reference_de(cur, flip, nb, D, J):
    p = C.c_int()
    de = lib().bc_oracle_reference_de(cur, flip, nb, D, J, C.byref(p))
    return de, p.value


def init_random(L, seed, replica=0):
    s = np.zeros((L, L), np.int8)
    lib().bc_oracle_init_random(L, seed, replica, s.ctypes.data)
    return s


def observables(spins, boundary=PERIODIC):
    s = np.ascontiguousarray(spins, np.int8)
    L = s.shape[0]
    M, Q, B = C.c_int64(), C.c_int64(), C.c_int64()
    lib().bc_oracle_observables(L, boundary, s.ctypes.data, C.byref(M), C.byref(Q), C.byref(B))
    return M.value, Q.value, B.value


def sweep(spins, table, seed, replica=0, t0=0, n_sweeps=1, measure_every=0, boundary=PERIODIC):
    """In-place sweeps on `spins` (L x L int8).  Returns structured obs array."""
    assert spins.dtype == np.int8 and spins.flags.c_contiguous
    L = spins.shape[0]
    n_meas = ((t0 + n_sweeps) // measure_every - t0 // measure_every) if measure_every > 0 else 0
    obs = np.zeros(n_meas, OBS_DTYPE)
    tab = np.ascontiguousarray(table, np.uint32)
    n = lib().bc_oracle_sweep(L, boundary, spins.ctypes.data, tab.ctypes.data, seed, replica, t0, n_sweeps,
                              measure_every, obs.ctypes.data if n_meas else None)
    if n < 0:
        raise ValueError(f"bc_oracle_sweep failed: {n}")
    assert n == n_meas
    return obs


def half_sweep_rows(spins, table, seed, replica, t, colour, y0, y1, boundary=PERIODIC):
    L = spins.shape[0]
    tab = np.ascontiguousarray(table, np.uint32)
    lib().bc_oracle_half_sweep_rows(L, boundary, spins.ctypes.data, tab.ctypes.data, seed, replica, t, colour, y0, y1)


def tie_count(reset=False):
    return lib().bc_oracle_tie_count(1 if reset else 0)


def enumerate_exact(L, T, D, J=1.0):
    out = np.zeros(5, np.float64)
    rc = lib().bc_oracle_enumerate(L, T, D, J, out.ctypes.data)
    if rc != 0:
        raise ValueError("enumerate: L must be <= 4")
    return dict(zip(("E", "absm", "q", "U4", "chi"), out))


def reference_metropolis(lattice_i32, T, D, J, n_attempts, seed):
    assert lattice_i32.dtype == np.int32 and lattice_i32.flags.c_contiguous
    L = lattice_i32.shape[0]
    return lib().bc_oracle_reference_metropolis(L, lattice_i32.ctypes.data, T, D, J, n_attempts, seed)


def bond_threshold(T, J=1.0):
    return lib().bc_oracle_bond_threshold(T, J)


def cluster_update(spins, T, J, seed, replica=0, cstep=0, boundary=PERIODIC):
    """One Swendsen-Wang move on the +-1 spins (in place)."""
    assert spins.dtype == np.int8 and spins.flags.c_contiguous
    lib().bc_oracle_cluster_update(spins.shape[0], boundary, spins.ctypes.data, bond_threshold(T, J), seed, replica, cstep)


def cluster_moments(spins, kind, T, J, seed, replica=0, mstep=0, boundary=PERIODIC):
    """(sum size^2, max size, number of clusters) of the spin (kind 0) or FK (kind 1) clusters."""
    s = np.ascontiguousarray(spins, np.int8)
    out = np.zeros(3, np.int64)
    lib().bc_oracle_cluster_moments(s.shape[0], boundary, s.ctypes.data, kind, bond_threshold(T, J), seed, replica, mstep,
                                    out.ctypes.data)
    return tuple(int(x) for x in out)


def cluster_labels(spins, kind, T, J, seed, replica=0, mstep=0, boundary=PERIODIC):
    """Smallest site index of every site's cluster (-1 for zero spins); kind 0 = spin clusters, 1 = FK clusters."""
    s = np.ascontiguousarray(spins, np.int8)
    out = np.zeros(s.shape, np.int32)
    lib().bc_oracle_cluster_labels(s.shape[0], boundary, s.ctypes.data, kind, bond_threshold(T, J), seed, replica, mstep,
                                   out.ctypes.data)
    return out


def gap_histogram(spins, kind, T, J, seed, replica=0, mstep=0, boundary=PERIODIC, variant=0):
    """Gap-size histogram of one configuration as gap_size_statistics.cpp derives it from a cluster file
    (variant 0: L/2 bins; variant 1 = open variant: L-1 bins, no folding, no wrap-around gap)."""
    s = np.ascontiguousarray(spins, np.int8)
    L = s.shape[0]
    hist = np.zeros(L // 2 if variant == 0 else L - 1, np.int64)
    lib().bc_oracle_gap_histogram(L, boundary, s.ctypes.data, kind, bond_threshold(T, J), seed, replica, mstep, variant,
                                  hist.ctypes.data)
    return hist
