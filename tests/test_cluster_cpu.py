"""The cluster kernels' tile algorithm on the CPU.

Every phase of cl_label_kernel / cl_border_kernel / cl_flip_kernel / cl_labels_kernel is a host + device function
(blume_capel_b200/csrc/bc_cluster_tile.cuh); tests/cpu/cluster_emulation.cpp runs them lane by lane in the kernels'
order.  The development container has no GPU, so this is where the bit-parallel labelling (row masks, bit-plane
bond draws, run-start union-find, perimeter flags, border merge, flip masks) meets the oracle first; the GPU tests
(tests/test_gpu_cluster.py) then check the kernels themselves.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpu", "cluster_emulation.cpp")
LIB = os.path.join(ROOT, "tests", "cpu", "libcluster_emulation.so")
HDR = os.path.join(ROOT, "blume_capel_b200", "csrc", "bc_cluster_tile.cuh")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        cxx = "/usr/bin/g++" if os.access("/usr/bin/g++", os.X_OK) else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    lib = C.CDLL(LIB)
    lib.emu_cluster_update.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int64]
    lib.emu_cluster_labels.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32,
                                       C.c_int64, C.c_void_p]
    return lib


CASES = [(2, 0, 0.608), (4, 0, 0.608), (8, 0, 2.0), (12, 0, 0.608), (33, 1, 0.608), (64, 0, 0.608), (64, 1, 0.3), (96, 0, 1.2),
         (100, 0, 0.608), (128, 0, 0.608), (130, 0, 2.27), (161, 1, 0.608), (192, 0, 0.3)]


@pytest.mark.parametrize("L,boundary,T", CASES)
@pytest.mark.parametrize("zeros", [0.33, 0.05, 0.0])
def test_tile_algorithm_matches_oracle(emu, oracle, L, boundary, T, zeros):
    rng = np.random.default_rng(1000 * L + int(100 * zeros))
    s = rng.choice(np.array([-1, 0, 1], np.int8), size=(L, L), p=[(1 - zeros) / 2, zeros, (1 - zeros) / 2]).astype(np.int8)
    if zeros == 0.0:  # ordered: one cluster spans every tile
        s[:] = 1
        s[rng.integers(0, L), rng.integers(0, L)] = -1
    seed, rep, cstep = 0xABC0 + L, 3, 5
    thr = oracle.bond_threshold(T, 1.0)
    for ghost in (0, 1):  # plain handle / array geometry of a one-rank slab
        want = s.copy()
        oracle.cluster_update(want, T, 1.0, seed, rep, cstep, boundary)
        got = s.copy()
        emu.emu_cluster_update(L, boundary, ghost, 1, got.ctypes.data, thr, seed, rep, cstep)
        assert np.array_equal(got, want), (L, boundary, ghost)
        for kind in (0, 1):
            lab = np.zeros((L, L), np.uint32)
            emu.emu_cluster_labels(L, boundary, ghost, s.ctypes.data, 1 if kind == 0 else 0, thr, seed, rep, (1 << 62) + 7,
                                   lab.ctypes.data)
            ref = oracle.cluster_labels(s, kind, T, 1.0, seed, rep, 7, boundary).astype(np.int64)
            ref[ref < 0] = 0xFFFFFFFF
            assert np.array_equal(lab.astype(np.int64), ref), (L, boundary, ghost, kind)


@pytest.mark.parametrize("L,boundary,n_ranks,T", [(128, 0, 2, 0.608), (256, 0, 4, 0.608), (256, 1, 2, 0.608), (512, 0, 8, 0.3),
                                                  (192, 0, 3, 1.2), (128, 1, 2, 0.3)])
@pytest.mark.parametrize("zeros", [0.3, 0.0])
def test_multi_rank_slab_decomposition_matches_oracle(emu, oracle, L, boundary, n_ranks, T, zeros):
    """The cluster move on a lattice split into row slabs (BASELINE config 4): every rank labels the tiles of its rows
    (the row below its last one is the bottom ghost row), the global union-find is sharded by rows and the border merge
    crosses ranks through the rank below's perimeter roots; the result must be the single-lattice oracle's."""
    rng = np.random.default_rng(77 * L + n_ranks)
    s = rng.choice(np.array([-1, 0, 1], np.int8), size=(L, L), p=[(1 - zeros) / 2, zeros, (1 - zeros) / 2]).astype(np.int8)
    if zeros == 0.0:
        s[:] = -1
        s[L // 3, :] = 1  # one stripe, so that two clusters span every slab border ... of a torus: one, wrapped
    seed, rep, cstep = 0x51AB + L, 1, 2
    want = s.copy()
    oracle.cluster_update(want, T, 1.0, seed, rep, cstep, boundary)
    got = s.copy()
    assert emu.emu_cluster_update(L, boundary, 1, n_ranks, got.ctypes.data, oracle.bond_threshold(T, 1.0), seed, rep, cstep) == 0
    assert np.array_equal(got, want), (L, boundary, n_ranks)
