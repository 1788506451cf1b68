"""ctypes binding of include/bc_engine.h.

Mirrors the reference's run interface for the hot path: the object owns what main()
owns in /root/reference/main.cpp (lattice main.cpp:335, parameters B, D, J main.cpp:24,
321-323) and ``sweep`` replaces the loop body ``refill_random(); step(lattice);``
(main.cpp:343-347, 363-367).  There is no CPU fallback: if the CUDA library is missing or
no GPU is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

PERIODIC, OPEN = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))


class BCError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"bc_engine error {code}: {msg}")
        self.code = code


class Obs(C.Structure):
    _fields_ = [("sweep", C.c_int64), ("M", C.c_int64), ("Q", C.c_int64), ("B", C.c_int64)]


OBS_DTYPE = np.dtype([("sweep", "<i8"), ("M", "<i8"), ("Q", "<i8"), ("B", "<i8")])

_lib = None

# every symbol include/bc_engine.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "bc_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64]),
    "bc_slab_unique_id": (C.c_int, [_P]),
    "bc_create_slab": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, _P]),
    "bc_slab_ipc_export": (C.c_int, [_P, _P]),
    "bc_slab_ipc_connect": (C.c_int, [_P, _P, _P]),
    "bc_slab_cluster_export": (C.c_int, [_P, _P]),
    "bc_slab_cluster_connect": (C.c_int, [_P, _P]),
    "bc_local_rows": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bc_destroy": (None, [_P]),
    "bc_last_error": (C.c_char_p, [_P]),
    "bc_abi_version": (C.c_int, []),
    "bc_set_params": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double]),
    "bc_get_table": (C.c_int, [_P, C.c_int, _P]),
    "bc_init_random": (C.c_int, [_P, C.c_int]),
    "bc_upload": (C.c_int, [_P, C.c_int, _P]),
    "bc_download": (C.c_int, [_P, C.c_int, _P]),
    "bc_upload_all": (C.c_int, [_P, _P]),
    "bc_download_all": (C.c_int, [_P, _P]),
    "bc_upload_all_async": (C.c_int, [_P, _P]),
    "bc_download_all_async": (C.c_int, [_P, _P]),
    "bc_packed_size": (C.c_int64, [_P, C.c_int]),
    "bc_upload_packed": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "bc_download_packed": (C.c_int, [_P, C.c_int, _P, C.c_int]),
    "bc_pack_spins": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "bc_unpack_spins": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "bc_get_sweep_counter": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "bc_set_sweep_counter": (C.c_int, [_P, C.c_int64]),
    "bc_sweep": (C.c_int64, [_P, C.c_int64, C.c_int64, _P]),
    "bc_sweep_group": (C.c_int64, [C.POINTER(_P), C.c_int, C.c_int64, C.c_int64]),
    "bc_cluster_update": (C.c_int, [_P, C.c_int64]),
    "bc_get_cluster_counter": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "bc_set_cluster_counter": (C.c_int, [_P, C.c_int64]),
    "bc_cluster_moments": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "bc_gap_histogram": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, _P]),
    "bc_cluster_lists": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bc_read_obs": (C.c_int64, [_P, _P, C.c_int64]),
    "bc_measure": (C.c_int, [_P, _P]),
    "bc_sync": (C.c_int, [_P]),
    "bc_launch_count": (C.c_int64, [_P]),
    "bc_last_sweep_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "bc_timer_start": (C.c_int, [_P]),
    "bc_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "bc_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "bc_describe_plan": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bc_tie_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
}


def library_path() -> str:
    return os.environ.get("BC_ENGINE_LIB", os.path.join(_HERE, "libbc_engine.so"))


def load_library():
    """Load libbc_engine.so (in-tree build) and type every exported entry point."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise BCError(-2, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def slab_unique_id() -> bytes:
    """128-byte NCCL unique id (call on rank 0, broadcast to the other ranks)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.bc_slab_unique_id(buf)
    if rc != 0:
        raise BCError(rc, (lib.bc_last_error(None) or b"").decode())
    return buf.raw


def sweep_group(engines, n_sweeps: int, measure_every: int = 0, fetch: bool = True):
    """n_sweeps of several engines (different L, same device and boundary) with ONE kernel launch per half-sweep
    (bc_sweep_group).  Returns one obs array (n_meas, n_replicas) per engine, or None."""
    lib = load_library()
    arr = (_P * len(engines))(*[e._h for e in engines])
    t0 = engines[0].sweep_counter
    rc = lib.bc_sweep_group(arr, len(engines), n_sweeps, measure_every)
    if rc < 0:
        raise BCError(rc, (lib.bc_last_error(engines[0]._h) or b"").decode())
    if measure_every > 0 and fetch:
        n_meas = (t0 + n_sweeps) // measure_every - t0 // measure_every
        return [e.read_obs(n_meas) for e in engines]
    return None


class Engine:
    """n_replicas independent L x L spin-1 lattices on one GPU."""

    def __init__(self, L: int, boundary: int = PERIODIC, n_replicas: int = 1, device: int = 0, seed: int = 0,
                 slab: Optional[tuple] = None):
        """slab = (n_ranks, rank, nccl_id_bytes_or_None): this handle owns one row slab of the lattice."""
        self._lib = load_library()
        self._h = _P()
        if slab is None:
            rc = self._lib.bc_create(C.byref(self._h), L, boundary, n_replicas, 0, device, seed & (2**64 - 1))
        else:
            n_ranks, rank, nid = slab
            buf = C.create_string_buffer(bytes(nid), 128) if nid is not None else None
            rc = self._lib.bc_create_slab(C.byref(self._h), L, boundary, n_replicas, 0, device, seed & (2**64 - 1),
                                          n_ranks, rank, buf)
        if rc != 0:
            raise BCError(rc, (self._lib.bc_last_error(None) or b"").decode())
        self.L, self.boundary, self.n_replicas, self.device, self.seed = L, boundary, n_replicas, device, seed
        rows, y0 = C.c_int(), C.c_int()
        self._lib.bc_local_rows(self._h, C.byref(rows), C.byref(y0))
        self.rows, self.first_row = rows.value, y0.value

    # -- helpers
    def _check(self, rc: int) -> int:
        if rc < 0:
            raise BCError(rc, (self._lib.bc_last_error(self._h) or b"").decode())
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- parameters
    def set_params(self, T: float, D: float, J: float = 1.0, replica: int = -1):
        self._check(self._lib.bc_set_params(self._h, replica, T, D, J))

    def table(self, replica: int = 0) -> np.ndarray:
        t = np.zeros(54, np.uint32)
        self._check(self._lib.bc_get_table(self._h, replica, t.ctypes.data))
        return t

    # -- state
    def init_random(self, replica: int = -1):
        self._check(self._lib.bc_init_random(self._h, replica))

    def upload(self, spins: np.ndarray, replica: Optional[int] = None):
        a = np.ascontiguousarray(spins, dtype=np.int8)
        if replica is None:
            assert a.size == self.n_replicas * self.rows * self.L
            self._check(self._lib.bc_upload_all(self._h, a.ctypes.data))
        else:
            assert a.size == self.rows * self.L
            self._check(self._lib.bc_upload(self._h, replica, a.ctypes.data))

    def download(self, replica: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        if replica is None:
            a = out if out is not None else np.empty((self.n_replicas, self.rows, self.L), np.int8)
            self._check(self._lib.bc_download_all(self._h, a.ctypes.data))
        else:
            a = out if out is not None else np.empty((self.rows, self.L), np.int8)
            self._check(self._lib.bc_download(self._h, replica, a.ctypes.data))
        return a

    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(192)
        self._check(self._lib.bc_slab_ipc_export(self._h, buf))
        return buf.raw

    def ipc_connect(self, prev_blob: Optional[bytes], next_blob: Optional[bytes]):
        """Switch the slab halo exchange to direct peer stores (blobs of the ranks above / below, None at an open edge)."""
        p = C.create_string_buffer(prev_blob, 192) if prev_blob is not None else None
        n = C.create_string_buffer(next_blob, 192) if next_blob is not None else None
        self._check(self._lib.bc_slab_ipc_connect(self._h, p, n))

    def cluster_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self._lib.bc_slab_cluster_export(self._h, buf))
        return buf.raw

    def cluster_connect(self, blobs):
        """blobs: the 128-byte cluster_export() of every rank, in rank order (enables cluster_update on the decomposition)."""
        raw = b"".join(blobs)
        self._check(self._lib.bc_slab_cluster_connect(self._h, C.create_string_buffer(raw, len(raw))))

    def upload_async(self, spins: np.ndarray):
        """All replicas, no wait: `spins` (pinned, C-contiguous int8) must stay untouched until sync()."""
        assert spins.dtype == np.int8 and spins.flags.c_contiguous and spins.size == self.n_replicas * self.rows * self.L
        self._check(self._lib.bc_upload_all_async(self._h, spins.ctypes.data))

    def download_async(self, out: np.ndarray):
        assert out.dtype == np.int8 and out.flags.c_contiguous and out.size == self.n_replicas * self.rows * self.L
        self._check(self._lib.bc_download_all_async(self._h, out.ctypes.data))

    # -- 2-bit wire format (four sites per byte; conversion on the device)
    def packed_size(self, n_lattices: Optional[int] = None) -> int:
        return self._check(self._lib.bc_packed_size(self._h, self.n_replicas if n_lattices is None else n_lattices))

    def upload_packed(self, packed: np.ndarray, replica: int = -1, wait: bool = True):
        assert packed.dtype == np.uint8 and packed.flags.c_contiguous
        assert packed.size == self.packed_size(self.n_replicas if replica < 0 else 1)
        self._check(self._lib.bc_upload_packed(self._h, replica, packed.ctypes.data, 0 if wait else 1))

    def download_packed(self, replica: int = -1, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        n = self.packed_size(self.n_replicas if replica < 0 else 1)
        a = out if out is not None else np.empty(n, np.uint8)
        assert a.dtype == np.uint8 and a.flags.c_contiguous and a.size == n
        self._check(self._lib.bc_download_packed(self._h, replica, a.ctypes.data, 0 if wait else 1))
        return a

    @property
    def sweep_counter(self) -> int:
        t = C.c_int64()
        self._check(self._lib.bc_get_sweep_counter(self._h, C.byref(t)))
        return t.value

    @sweep_counter.setter
    def sweep_counter(self, t: int):
        self._check(self._lib.bc_set_sweep_counter(self._h, t))

    # -- hot path
    def sweep(self, n_sweeps: int, measure_every: int = 0, fetch: bool = True) -> Optional[np.ndarray]:
        """Run n_sweeps; returns obs records shaped (n_meas, n_replicas) (structured) or None."""
        if measure_every > 0 and fetch:
            t0 = self.sweep_counter
            n_meas = (t0 + n_sweeps) // measure_every - t0 // measure_every
            buf = np.zeros((max(n_meas, 0), self.n_replicas), OBS_DTYPE)
            n = self._check(self._lib.bc_sweep(self._h, n_sweeps, measure_every, buf.ctypes.data if n_meas else None))
            assert n == n_meas * self.n_replicas, (n, n_meas)
            return buf
        self._check(self._lib.bc_sweep(self._h, n_sweeps, measure_every, None))
        return None

    def cluster_update(self, n_updates: int = 1):
        """Swendsen-Wang move on the +-1 spins of every replica (zeros never join a cluster)."""
        self._check(self._lib.bc_cluster_update(self._h, n_updates))

    @property
    def cluster_counter(self) -> int:
        c = C.c_int64()
        self._check(self._lib.bc_get_cluster_counter(self._h, C.byref(c)))
        return c.value

    @cluster_counter.setter
    def cluster_counter(self, c: int):
        self._check(self._lib.bc_set_cluster_counter(self._h, c))

    def cluster_moments(self, kind: int = 0, mstep: int = 0) -> np.ndarray:
        """Per replica (sum size^2, max size, number of clusters) of the spin (kind 0) or FK (kind 1) clusters."""
        buf = np.zeros(self.n_replicas, np.dtype([("sum_size_sq", "<i8"), ("max_size", "<i8"), ("n_clusters", "<i8")]))
        self._check(self._lib.bc_cluster_moments(self._h, kind, mstep, buf.ctypes.data))
        return buf

    def gap_histogram(self, kind: int = 0, mstep: int = 0, variant: int = 0, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Gap-size histogram of the current configurations (one sample per replica added to `out`):
        (n_replicas, L/2) for variant 0 (the reference's), (n_replicas, L-1) for the open variant 1."""
        bins = self.L // 2 if variant == 0 else self.L - 1
        h = out if out is not None else np.zeros((self.n_replicas, bins), np.int64)
        assert h.dtype == np.int64 and h.flags.c_contiguous and h.shape == (self.n_replicas, bins)
        self._check(self._lib.bc_gap_histogram(self._h, kind, mstep, variant, h.ctypes.data))
        return h

    def cluster_lists(self, replica: int = 0, kind: int = 0, mstep: int = 0):
        """(sites, starts, signs) of one replica's clusters, ordered like the reference's cluster files."""
        N = self.L * self.L
        sites, starts, signs = np.empty(N, np.uint32), np.empty(N + 1, np.uint32), np.empty(N, np.int8)
        ns, nc = C.c_int64(), C.c_int64()
        self._check(self._lib.bc_cluster_lists(self._h, replica, kind, mstep, sites.ctypes.data, starts.ctypes.data,
                                               signs.ctypes.data, C.byref(ns), C.byref(nc)))
        return sites[:ns.value].copy(), starts[:nc.value + 1].copy(), signs[:nc.value].copy()

    def read_obs(self, n_meas: int) -> np.ndarray:
        buf = np.zeros((n_meas, self.n_replicas), OBS_DTYPE)
        n = self._check(self._lib.bc_read_obs(self._h, buf.ctypes.data, buf.size))
        return buf.reshape(-1)[:n].reshape(-1, self.n_replicas)

    def measure(self) -> np.ndarray:
        buf = np.zeros(self.n_replicas, OBS_DTYPE)
        self._check(self._lib.bc_measure(self._h, buf.ctypes.data))
        return buf

    def sync(self):
        self._check(self._lib.bc_sync(self._h))

    # -- introspection
    @property
    def launch_count(self) -> int:
        return self._lib.bc_launch_count(self._h)

    def last_sweep_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.bc_last_sweep_ms(self._h, C.byref(ms)))
        return ms.value

    def timer_start(self):
        self._check(self._lib.bc_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.bc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_option(self, name: str, value: int):
        self._check(self._lib.bc_set_option(self._h, name.encode(), value))

    def plan(self) -> dict:
        f, r, g, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.bc_describe_plan(self._h, C.byref(f), C.byref(r), C.byref(g), C.byref(b)))
        return {"fast": bool(f.value), "resident": f.value == 2, "pair": f.value == 3, "rows_per_thread": r.value, "grid": g.value,
                "block": b.value}

    @property
    def tie_count(self) -> int:
        t = C.c_int64()
        self._check(self._lib.bc_tie_count(self._h, C.byref(t)))
        return t.value


def pack_spins(spins: np.ndarray) -> np.ndarray:
    """int8 spins (..., L) -> 2-bit wire format (rows padded to whole bytes), on the host."""
    a = np.ascontiguousarray(spins, np.int8)
    L = a.shape[-1]
    rows = a.size // L
    out = np.empty(rows * ((L + 3) // 4), np.uint8)
    rc = load_library().bc_pack_spins(a.ctypes.data, L, rows, out.ctypes.data)
    if rc != 0:
        raise BCError(rc, "bc_pack_spins")
    return out


def unpack_spins(packed: np.ndarray, L: int) -> np.ndarray:
    p = np.ascontiguousarray(packed, np.uint8)
    rows = p.size // ((L + 3) // 4)
    out = np.empty((rows, L), np.int8)
    rc = load_library().bc_unpack_spins(p.ctypes.data, L, rows, out.ctypes.data)
    if rc != 0:
        raise BCError(rc, "bc_unpack_spins")
    return out
