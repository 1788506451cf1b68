"""Config 4 of BASELINE.json: one L x L lattice (default L=65536, 4 GiB int8) split into row slabs
across the ranks of a torchrun job, halo exchange over NCCL per half-sweep.  Prints one JSON line.
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/bench_slab.py [L] [sweeps]
(G = 1 works without torchrun.)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import blume_capel_b200 as bc

L = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
SWEEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
nid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(bc.slab_unique_id()), dtype=torch.uint8))
    dist.broadcast(t, 0)
    nid = bytes(t.cpu().numpy().tobytes())
e = bc.Engine(L, bc.PERIODIC, 1, local, 0x5EED0004, slab=(world, rank, nid))
P2P = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if P2P and world > 1:  # halo rows by direct peer stores instead of NCCL send/recv
    mine = torch.frombuffer(bytearray(e.ipc_export()), dtype=torch.uint8).cuda()
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    blobs = [bytes(t.cpu().numpy().tobytes()) for t in allb]
    e.ipc_connect(blobs[(rank - 1) % world], blobs[(rank + 1) % world])
    dist.barrier()
FUSED = int(sys.argv[4]) if len(sys.argv) > 4 else 1
e.set_option("fused_halo", FUSED)
e.set_params(0.608, 1.966, 1.0)
e.init_random()
e.sweep(5)
e.sync()
if dist: dist.barrier()
torch.cuda.synchronize()
e.timer_start()
e.sweep(SWEEPS, SWEEPS, fetch=False)
ms = e.timer_stop()
obs = e.read_obs(1)
tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
if dist: dist.all_reduce(tms, op=dist.ReduceOp.MAX)
if rank == 0:
    ups = float(L) * L * SWEEPS / (tms.item() * 1e6)
    print(json.dumps({"config": f"L={L} periodic int8, {world} row slab(s), halo exchange per half-sweep by " + (("peer stores fused into the boundary-strip kernel" if FUSED else "peer stores (separate push kernel)") + " over NVLink (CUDA IPC) + epoch flags" if P2P and world > 1 else "NCCL send/recv"),
                      "n_gpus": world, "sweeps": SWEEPS, "ms": tms.item(), "updates_per_ns": ups,
                      "updates_per_ns_per_gpu": ups / world, "hbm_frac_per_gpu": ups / world * 3 / 6546.2,
                      "M": int(obs[0, 0]["M"]), "Q": int(obs[0, 0]["Q"]), "B": int(obs[0, 0]["B"])}), flush=True)
e.close()
if dist: dist.destroy_process_group()
