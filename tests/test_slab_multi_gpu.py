"""Slab decomposition across ranks.

* CPU (gloo, world_size 2): the decomposition LOGIC -- row slabs, one-row halo of the freshly
  updated colour per half-sweep to the ring neighbours, global-coordinate RNG -- is exercised with
  the oracle's row-range half-sweep as the per-rank update and torch.distributed send/recv as the
  exchange, and must reproduce the single-process oracle bit for bit.
* GPU (nccl, >= 2 GPUs): the product path bc_create_slab + ncclSend/ncclRecv against the single-GPU
  engine; launched as a torchrun subprocess, skipped when fewer than 2 GPUs are visible.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "slab_worker.py")


def _torchrun(n, args, timeout=600):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("boundary", ["periodic", "open"])
def test_slab_logic_gloo_world2(boundary):
    r = _torchrun(2, ["cpu", boundary])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SLAB_OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("boundary", ["periodic", "open"])
def test_slab_nccl_two_gpus(boundary):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["gpu", boundary])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SLAB_OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("boundary,L,rows,fused", [("periodic", 2048, 0, 1), ("open", 2048, 0, 1), ("periodic", 2048, 64, 1),
                                                   ("periodic", 8192, 64, 1), ("open", 8192, 16, 1), ("periodic", 2048, 0, 0),
                                                   ("periodic", 256, 0, 1)])
def test_slab_peer_store_halo_two_gpus(boundary, L, rows, fused):
    """Halo rows stored directly into the neighbour's ghost rows over NVLink (CUDA IPC peer pointers +
    epoch flags) instead of NCCL send/recv.  fused = 1: the whole half-sweep of a slab is one launch of the HALO
    kernels (plain and pair-table variants depending on the strip height; L = 8192 has two CTAs per
    boundary strip); fused = 0: boundary strips, push kernel, signal kernel, interior as separate launches."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _torchrun(2, ["gpu_p2p", boundary, str(L), str(rows), str(fused)])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SLAB_OK" in r.stdout
