"""Worker for tests/test_slab_multi_gpu.py (run under torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TC = (0.608, 1.966, 1.0)


def cpu_slab(boundary):
    """Oracle per rank + gloo halo exchange == whole-lattice oracle."""
    from oracle import oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    L, n, seed = 16, 6, 77
    b = oracle.PERIODIC if boundary == "periodic" else oracle.OPEN
    tab = oracle.build_table(*TC)
    full = oracle.init_random(L, seed, 0)
    want = full.copy()
    oracle.sweep(want, tab, seed, 0, 0, n, 0, b)
    rows = L // world
    y0, y1 = rank * rows, (rank + 1) * rows
    # local view: the full-size array is used as storage, but only rows [y0, y1) and the two halo rows
    # (y0-1, y1 mod L) are ever kept up to date on this rank
    loc = np.zeros_like(full)
    loc[y0:y1] = full[y0:y1]
    up, dn = (rank - 1) % world, (rank + 1) % world
    periodic = b == oracle.PERIODIC

    def exchange():
        first, last = torch.from_numpy(loc[y0].copy()), torch.from_numpy(loc[y1 - 1].copy())
        got_dn, got_up = torch.empty_like(first), torch.empty_like(first)
        ops = []
        # tag 0: rows travelling upwards (my first row -> bottom halo of the rank above)
        # tag 1: rows travelling downwards (my last row -> top halo of the rank below)
        if periodic or rank > 0:
            ops += [dist.P2POp(dist.isend, first, up, tag=0), dist.P2POp(dist.irecv, got_up, up, tag=1)]
        if periodic or rank < world - 1:
            ops += [dist.P2POp(dist.isend, last, dn, tag=1), dist.P2POp(dist.irecv, got_dn, dn, tag=0)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if periodic or rank > 0:
            loc[(y0 - 1) % L] = got_up.numpy()
        if periodic or rank < world - 1:
            loc[y1 % L] = got_dn.numpy()

    exchange()
    for t in range(n):
        for col in (0, 1):
            oracle.half_sweep_rows(loc, tab, seed, 0, t, col, y0, y1, b)
            exchange()
    ok = np.array_equal(loc[y0:y1], want[y0:y1])
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SLAB_OK" if flag.item() == 1 else "SLAB_MISMATCH", flush=True)
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


def connect_p2p(bc, e, rank, world, periodic):
    """All-gather the 192-byte IPC blobs and hand each rank its ring neighbours'."""
    mine = torch.frombuffer(bytearray(e.ipc_export()), dtype=torch.uint8).cuda()
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    blobs = [bytes(t.cpu().numpy().tobytes()) for t in allb]
    prev = blobs[(rank - 1) % world] if (periodic or rank > 0) else None
    nxt = blobs[(rank + 1) % world] if (periodic or rank + 1 < world) else None
    e.ipc_connect(prev, nxt)
    dist.barrier()


def gpu_slab(boundary, p2p=False, L=None, rows=0, fused=1):
    """bc_create_slab over NCCL (or peer stores) == single-GPU engine, lattice and fused observables."""
    import blume_capel_b200 as bc
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, n, seed = (L or (2048 if p2p else 256)), 8, 4321
    b = bc.PERIODIC if boundary == "periodic" else bc.OPEN
    nid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        nid.copy_(torch.frombuffer(bytearray(bc.slab_unique_id()), dtype=torch.uint8))
    dist.broadcast(nid, 0)
    e = bc.Engine(L, b, 1, local, seed, slab=(world, rank, bytes(nid.cpu().numpy().tobytes())))
    if p2p:
        connect_p2p(bc, e, rank, world, boundary == "periodic")
        e.set_option("fused_halo", fused)
    e.set_option("rows_per_thread", rows)
    e.set_option("graph_block", 2)
    e.set_params(*TC)
    e.init_random()
    obs = e.sweep(n, 2)
    # the cluster move on the decomposition: union-find pieces and perimeter roots of the other ranks over peer memory
    mine_b = torch.frombuffer(bytearray(e.cluster_export()), dtype=torch.uint8).cuda()
    all_b = [torch.empty_like(mine_b) for _ in range(world)]
    dist.all_gather(all_b, mine_b)
    e.cluster_connect([bytes(t.cpu().numpy().tobytes()) for t in all_b])
    dist.barrier()
    for _ in range(2):
        e.cluster_update(1)
        e.sweep(2)
    if p2p:  # a second lattice through upload: exercises the out-of-cadence exchange
        e.upload(e.download(0), 0)
        obs2 = e.sweep(3, 0)
    mine = e.download(0)
    m = e.measure()
    e.close()
    ok = True
    if rank == 0:
        s = bc.Engine(L, b, 1, local, seed)
        s.set_params(*TC)
        s.init_random()
        so = s.sweep(n, 2)
        for _ in range(2):
            s.cluster_update(1)
            s.sweep(2)
        if p2p:
            s.sweep(3, 0)
        whole = s.download(0)
        sm = s.measure()
        s.close()
        ok = np.array_equal(mine, whole[: L // world])
        for f in ("M", "Q", "B"):
            ok = ok and np.array_equal(obs[:, 0][f], so[:, 0][f]) and m[0][f] == sm[0][f]
        ref = torch.from_numpy(whole).cuda()
    else:
        ref = torch.empty((L, L), dtype=torch.int8, device="cuda")
    dist.broadcast(ref, 0)
    if rank != 0:
        ok = np.array_equal(mine, ref.cpu().numpy()[rank * (L // world):(rank + 1) * (L // world)])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SLAB_OK" if flag.item() == 1 else "SLAB_MISMATCH", flush=True)
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    mode, boundary = sys.argv[1], sys.argv[2]
    if mode == "cpu":
        sys.exit(cpu_slab(boundary))
    L = int(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    fused = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    sys.exit(gpu_slab(boundary, p2p=(mode == "gpu_p2p"), L=L, rows=rows, fused=fused))
