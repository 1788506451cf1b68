#!/usr/bin/env python
"""bench.py -- spin-updates/ns of the B200 Blume-Capel checkerboard Metropolis engine.

Contract (one JSON line on stdout from rank 0):
    python bench.py --gpus N --steps K --warmup W            # our CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code

Workload (BASELINE.json configs[1]): 2D square L=4096 periodic, single (T, D) = (0.608, 1.966)
near the tricritical point, int8 spins.  To keep the working set larger than the 126 MB L2
(timing rule) the launch is batched over 64 independent replicas of that lattice at that
(T, D) (1 GiB of spins): one "step" = SWEEPS_PER_STEP checkerboard sweeps of the whole
batch with the fused observables measured every MEASURE_EVERY sweeps.
  value : whole-job spin-updates/ns with the lattices resident in HBM (CUDA events on the
          engine's stream, max over ranks).
  e2e   : the same step through the C ABI from HOST buffers: bc_upload_all (pinned host ->
          device) + bc_sweep + observables device -> host + bc_download_all, all inside the
          timed region.
Multi-GPU (N > 1): one process per GPU under torchrun, the replicas are sharded (64 per GPU,
weak scaling), no data-path collective; torch.distributed only for barrier + max-over-ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L = 4096
N_REPLICAS = 64
T, D, J = 0.608, 1.966, 1.0
SWEEPS_PER_STEP = 100
MEASURE_EVERY = 10
SEED = 0x5EED0002
ALGO_BYTES_PER_UPDATE = 3.0  # int8: per sweep every spin is read twice and written once


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation
# ---------------------------------------------------------------------------------------------
def reference_sample(n_procs, L_ref=32, nsteps=1):
    """Runs n_procs independent instances of the UNMODIFIED reference binary (oracle/_ref/tri-L,
    built from /root/reference/main.cpp; OpenMP threads do not speed a single run up)
    and returns (Metropolis attempts, wall seconds).  Falls back to the oracle's literal port of
    main.cpp:89-112 when the binary is not present."""
    from oracle import ref_runner
    if ref_runner.available(L_ref):
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=n_procs) as ex:
            list(ex.map(lambda r: ref_runner.run_reference(L_ref, T, D, J, nsteps, run=r), range(n_procs)))
        wall = time.perf_counter() - t0
        return n_procs * ref_runner.metropolis_attempts(L_ref, nsteps), wall, "reference", \
            (f"{n_procs} x oracle/_ref/tri-{L_ref} (unmodified main.cpp, burner.sh:11 flags) burn=0 nsteps={nsteps}: "
             f"{9 * L_ref * L_ref} step() calls each = Metropolis attempts + Wolff clusters + mt19937 refills; "
             "rate counts Metropolis attempts only")
    import numpy as np
    from oracle import oracle
    lat = np.ones((64, 64), np.int32)
    n = 100_000_000
    t0 = time.perf_counter()
    oracle.reference_metropolis(lat, T, D, J, n, 1)
    wall = time.perf_counter() - t0
    return n, wall, "port", "oracle literal port of main.cpp:89-112, L=64, 1e8 attempts, 1 thread"


def cpu_model():
    """Model name of the host CPU (state what the reference was timed on)."""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.lower().startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def loop_only_sample(n_threads, attempts=30_000_000):
    """the Metropolis attempt loop alone (main.cpp:89-112 / the inner loop of step(),
    main.cpp:159-161), without Wolff clusters and without the mt19937 refills -- the oracle's literal port of that
    function (bc_oracle_reference_metropolis; the reference has no way to time it without editing main.cpp), one
    independent L = 64 lattice per thread.  Returns (attempts, wall seconds)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    lats = [np.ones((64, 64), np.int32) for _ in range(n_threads)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=n_threads) as ex:  # ctypes releases the GIL
        list(ex.map(lambda i: oracle.reference_metropolis(lats[i], T, D, J, attempts, 1 + i), range(n_threads)))
    return n_threads * attempts, time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        reference_sample(cores)
    attempts, wall = 0, 0.0
    kind = sample = None
    for _ in range(args.steps):
        a, w, kind, sample = reference_sample(cores)
        attempts += a
        wall += w
    value = attempts / wall / 1e9
    # the nominal L = 64 run (config 1: 36 864 step() calls per sample, about two minutes per
    # process) once, one process per core, and the attempt loop in isolation
    extra = {}
    if not args.no_l64:
        try:
            a64, w64, k64, s64 = reference_sample(cores, 64, 1)
            extra["l64_nominal"] = {"value": a64 / w64 / 1e9, "unit": "spin-updates/ns", "cores": cores, "kind": k64,
                                    "wall_s": w64, "sample": s64}
        except Exception as ex:
            extra["l64_nominal"] = {"error": str(ex)}
    try:
        al, wl = loop_only_sample(cores)
        extra["loop_only"] = {"value": al / wl / 1e9, "per_core": al / wl / 1e9 / cores, "unit": "spin-updates/ns",
                              "cores": cores, "kind": "port", "wall_s": wl,
                              "sample": f"{cores} threads x 3e7 attempts of oracle bc_oracle_reference_metropolis (literal "
                                        "main.cpp:89-112, L=64): Metropolis attempts only, no Wolff, no RNG refill"}
    except Exception as ex:
        extra["loop_only"] = {"error": str(ex)}
    line = {
        "impl": "reference", "metric": "spin-updates/ns", "value": value, "unit": "spin-updates/ns",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "spin-updates/ns", "cores": cores, "kind": kind, "sample": sample,
                         "cpu_model": cpu_model(), **extra},
        "e2e": {"value": value, "unit": "spin-updates/ns", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference cannot be built for L >= 512 (main.cpp:57-59); its per-attempt rate is size independent; "
                "value = whole-program rate at L=32 (i), cpu_baseline.l64_nominal = the same at L=64, "
                "cpu_baseline.loop_only = attempt loop alone (ii)",
    }
    emit(line)


def workload_config(n_gpus):
    return {"workload": f"L={L} periodic int8, (T,D,J)=({T},{D},{J}), {N_REPLICAS} replicas/GPU batched "
                        f"(1 GiB/GPU > 126 MB L2, no L2 flush needed), {SWEEPS_PER_STEP} sweeps/step, "
                        f"fused observables every {MEASURE_EVERY} sweeps",
            "L": L, "replicas_per_gpu": N_REPLICAS, "sweeps_per_step": SWEEPS_PER_STEP,
            "measure_every": MEASURE_EVERY, "boundary": "periodic", "storage": "int8",
            "parallelism": f"replicas x{n_gpus}", "l2": "working set > L2"}


def other_configs(bc, device):
    """Device-resident rates of BASELINE.json configs 3 and 5 on one GPU and of the cluster move (not the headline)."""
    out = {}

    def series(sizes, boundary, n_rep, points, sweeps, seed):
        engines = []
        for Ls in sizes:
            e = bc.Engine(Ls, boundary, n_rep, device, seed + Ls)
            for r in range(n_rep):
                t, d = points[r % len(points)]
                e.set_params(t, d, J, replica=r)
            e.init_random()
            e.sweep(4)
            engines.append(e)
        small = [e for e in engines if e.L < 128]
        group = sorted((e for e in engines if e.L >= 128), key=lambda e: -e.L)
        bc.sweep_group(group, 16)
        for e in engines:
            e.sync()
        t0 = time.perf_counter()
        for e in small:
            e.sweep(sweeps, 0, fetch=False)
        bc.sweep_group(group, sweeps, 0)
        for e in engines:
            e.sync()
        wall = time.perf_counter() - t0
        sites = sum(e.L * e.L * n_rep for e in engines)
        for e in engines:
            e.close()
        return sites * sweeps / (wall * 1e9)

    pts = [(0.58 + 0.06 * i / 7, 1.94 + 0.05 * (i % 8) / 7) for i in range(8)]
    out["config3_fss_ladder"] = {
        "workload": "L=32..2048 x 8 (T,D) points (one GPU's share of the 64-point grid), periodic, 400 sweeps; "
                    "L>=128 in one launch per half-sweep (bc_sweep_group), L=32,64 shared-memory resident",
        "updates_per_ns": series([32, 64, 128, 256, 512, 1024, 2048], bc.PERIODIC, 8, pts, 400, 0x5EED0003)}
    out["config5_open_boundary"] = {
        "workload": "open boundary, L=64..1024 x 64 replicas, 200 sweeps; L>=128 in one launch per half-sweep",
        "updates_per_ns": series([64, 128, 256, 512, 1024], bc.OPEN, 64, [(T, D)], 200, 0x5EED0005)}
    # the cluster move (bc_cluster_update: Swendsen-Wang on the +-1 spins, three launches per update straight on the
    # compact colour arrays): one L2-resident lattice and the bench batch (64 lattices, 1 GiB >> L2).  Algorithmic
    # bytes per site: label kernel 1 B spins read + 2 B labels written, flip kernel 2 B + 1 B read, spins of flipped
    # chunks written back (counted as 1 B): 7 B / site.
    peak, _ = measured_peak()
    out["cluster_move"] = {"workload": f"bc_cluster_update on L={L} lattices at (T,D)=({T},{D}) after 50 sweeps from a hot start",
                           "algorithmic_bytes_per_site": 7.0}
    for n_rep, n_upd in ((1, 40), (N_REPLICAS, 6)):
        e = bc.Engine(L, bc.PERIODIC, n_rep, device, SEED)
        e.set_params(T, D, J)
        e.init_random()
        e.sweep(50)
        e.cluster_update(2)
        e.sync()
        e.timer_start()
        e.cluster_update(n_upd)
        dt = e.timer_stop() * 1e-3 / n_upd
        e.close()
        rate = n_rep * L * L / (dt * 1e9)
        out["cluster_move"][f"x{n_rep}"] = {"ms_per_update": dt * 1e3, "sites_per_ns": rate,
                                            "hbm_frac_on_algorithmic_bytes": rate * 7.0 / peak}
    out["cluster_move"]["ms_per_update"] = out["cluster_move"]["x1"]["ms_per_update"]
    out["cluster_move"]["sites_per_ns"] = out["cluster_move"]["x1"]["sites_per_ns"]
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE config 4: one L = 65536 lattice (4 GiB int8) split into row slabs over the ranks
# ---------------------------------------------------------------------------------------------
SLAB_L = 65536
SLAB_SEED = 0x5EED0004
SLAB_WARMUP = 5
SLAB_SWEEPS = 100


def slab_golden():
    p = os.path.join(ROOT, "tests", "golden", f"slab_L{SLAB_L}.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def slab_record(bc, torch, dist, rank, world, local_rank, L=SLAB_L, sweeps=SLAB_SWEEPS):
    """Config 4 through bc_create_slab: halo exchange per half-sweep by NCCL send/recv and by peer stores fused into
    the sweep kernel, next to the same lattice on ONE GPU (rank 0, same invocation).  (M, Q, B) after 2 sweeps is
    compared with the CPU oracle's golden triple (tests/golden/make_slab_golden.py), after all sweeps with the
    one-GPU run and with the committed one-GPU golden.  A failure fails this record, never the bench line."""
    peak, _ = measured_peak()
    gold = slab_golden()
    total_sweeps = 2 + SLAB_WARMUP + sweeps
    rec = {"workload": f"L={L} periodic int8 (one lattice, {L * L / 2**30:.0f} GiB), {world} row slabs of {L // world} rows, "
                       f"2 + {SLAB_WARMUP} warm-up + {sweeps} timed sweeps, halo exchange per half-sweep",
           "L": L, "sweeps": sweeps, "seed": SLAB_SEED, "golden_file": f"tests/golden/slab_L{L}.json"}

    def run(n_ranks, r, nid, transport):
        e = bc.Engine(L, bc.PERIODIC, 1, local_rank, SLAB_SEED, slab=(n_ranks, r, nid))
        try:
            if transport == "peer_store" and n_ranks > 1:
                mine = torch.frombuffer(bytearray(e.ipc_export()), dtype=torch.uint8).cuda()
                allb = [torch.empty_like(mine) for _ in range(n_ranks)]
                dist.all_gather(allb, mine)
                blobs = [bytes(t.cpu().numpy().tobytes()) for t in allb]
                e.ipc_connect(blobs[(r - 1) % n_ranks], blobs[(r + 1) % n_ranks])
                dist.barrier()
            elif transport == "peer_store":
                e.ipc_connect(None, None)  # one slab is its own ring neighbour: same kernels, no NVLink
            e.set_params(T, D, J)
            e.init_random()
            e.sweep(2)
            early = e.measure()[0]
            e.sweep(SLAB_WARMUP)
            e.sync()
            if n_ranks > 1:
                dist.barrier()
            torch.cuda.synchronize()
            l0 = e.launch_count
            e.timer_start()
            e.sweep(sweeps)
            ms = e.timer_stop()
            launches = e.launch_count - l0
            final = e.measure()[0]
            e.sync()
            plan = e.plan()
            # the cluster move on the same decomposition (after the observables that are compared with the goldens):
            # union-find pieces and perimeter roots of the other ranks over peer memory, NCCL barriers between phases
            cluster = None
            try:
                if n_ranks > 1:
                    mine = torch.frombuffer(bytearray(e.cluster_export()), dtype=torch.uint8).cuda()
                    allb = [torch.empty_like(mine) for _ in range(n_ranks)]
                    dist.all_gather(allb, mine)
                    e.cluster_connect([bytes(t.cpu().numpy().tobytes()) for t in allb])
                    dist.barrier()
                e.cluster_update(1)
                e.sync()
                e.timer_start()
                e.cluster_update(3)
                cms = e.timer_stop() / 3
                after = e.measure()[0]
                if n_ranks > 1:
                    t_c = torch.tensor([cms], dtype=torch.float64, device=f"cuda:{local_rank}")
                    dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
                    cms = t_c.item()
                cluster = {"ms_per_update": cms, "sites_per_ns": float(L) * L / (cms * 1e6),
                           "zeros_conserved": int(after["Q"]) == int(final["Q"]), "M_after": int(after["M"])}
            except Exception as ex:
                cluster = {"error": str(ex)}
        finally:
            e.close()
        if n_ranks > 1:
            t_ms = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            ms = t_ms.item()
        ups = float(L) * L * sweeps / (ms * 1e6)
        return {"ms": ms, "updates_per_ns": ups, "updates_per_ns_per_gpu": ups / n_ranks,
                "hbm_frac_per_gpu": ups / n_ranks * ALGO_BYTES_PER_UPDATE / peak, "gpu_launches_per_rank": int(launches),
                "plan": plan, "obs_after_2": [int(early[f]) for f in ("M", "Q", "B")],
                "obs_final": [int(final[f]) for f in ("M", "Q", "B")], "cluster_move": cluster}

    # the same lattice on one GPU (rank 0 only; the other ranks wait at the barrier)
    single = None
    if rank == 0:
        try:
            single = run(1, 0, None, "peer_store")
        except Exception as ex:
            single = {"error": str(ex)}
    dist.barrier()
    rec["single_gpu"] = single
    nid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
    for transport in ("nccl", "peer_store"):
        try:
            if rank == 0:
                nid.copy_(torch.frombuffer(bytearray(bc.slab_unique_id()), dtype=torch.uint8))
            dist.broadcast(nid, 0)
            out = run(world, rank, bytes(nid.cpu().numpy().tobytes()), transport)
        except Exception as ex:
            out = {"error": str(ex)}
        if rank == 0 and "error" not in out:
            if single and "error" not in single:
                out["efficiency_vs_1gpu"] = out["updates_per_ns"] / (world * single["updates_per_ns"])
                out["equals_single_gpu"] = out["obs_final"] == single["obs_final"] and out["obs_after_2"] == single["obs_after_2"]
            g2 = gold.get("after", {}).get("2")
            out["equals_oracle_after_2"] = (out["obs_after_2"] == g2) if g2 else None
            gf = gold.get("gpu_n1", {}).get(str(total_sweeps))
            out["equals_committed_golden"] = (out["obs_final"] == gf) if gf else None
            cm, cs = out.get("cluster_move") or {}, (single or {}).get("cluster_move") or {}
            if "M_after" in cm and "M_after" in cs:  # four cluster moves later the magnetisation is still the one-GPU run's
                cm["equals_single_gpu"] = cm["M_after"] == cs["M_after"]
            checks = [out.get("equals_single_gpu"), out["equals_oracle_after_2"], out["equals_committed_golden"],
                      cm.get("equals_single_gpu"), cm.get("zeros_conserved")]
            out["ok"] = all(c is not False for c in checks) and any(c for c in checks)
        rec[transport] = out
    if rank == 0 and single and "error" not in single:
        g2 = gold.get("after", {}).get("2")
        single["equals_oracle_after_2"] = (single["obs_after_2"] == g2) if g2 else None
        gf = gold.get("gpu_n1", {}).get(str(total_sweeps))
        single["equals_committed_golden"] = (single["obs_final"] == gf) if gf else None
    rec["golden"] = {"oracle_after_2": gold.get("after", {}).get("2"),
                     "gpu_n1_final": gold.get("gpu_n1", {}).get(str(total_sweeps)), "total_sweeps": total_sweeps}
    return rec


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    import blume_capel_b200 as bc

    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    eng = bc.Engine(L, bc.PERIODIC, N_REPLICAS, local_rank, SEED + (rank << 40))
    eng.set_params(T, D, J)
    eng.init_random()
    n_meas = SWEEPS_PER_STEP // MEASURE_EVERY
    updates_per_step = float(L) * L * N_REPLICAS * SWEEPS_PER_STEP

    # ---- device-resident throughput
    for _ in range(args.warmup):
        eng.sweep(SWEEPS_PER_STEP, MEASURE_EVERY, fetch=False)
    eng.sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        eng.sweep(SWEEPS_PER_STEP, MEASURE_EVERY, fetch=False)
    ms = eng.timer_stop()
    launches = eng.launch_count - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    last_obs = eng.read_obs(n_meas)
    ties_resident = eng.tie_count

    # ---- for the record: ONE L=4096 lattice (16 MiB, L2 resident, launch-bound), not the headline
    single = None
    if rank == 0:
        e1 = bc.Engine(L, bc.PERIODIC, 1, local_rank, SEED)
        e1.set_params(T, D, J)
        e1.init_random()
        e1.sweep(64)
        e1.sync()
        e1.timer_start()
        e1.sweep(1024)
        single = float(L) * L * 1024 / (e1.timer_stop() * 1e6)
        e1.close()

    # ---- for the record as well (N = 1 only, outside every timed region of the headline): BASELINE configs 3 and 5
    # ---- with bc_sweep_group (one launch per half-sweep for all sizes >= 128) and the cluster move
    others = None
    if rank == 0 and world == 1:
        try:
            others = other_configs(bc, local_rank)
        except Exception as ex:  # informational: never fail the bench line over it
            others = {"error": str(ex)}

    # ---- end to end through the C ABI from pinned host buffers.  Several handles (streams) are used the way a
    # ---- caller with many batches would: while one batch is swept, other batches' lattices are uploaded /
    # ---- downloaded.  A step's 64 replicas travel as --e2e-split sub-batches (own handle each), so the first
    # ---- sweep starts after the first sub-batch has arrived instead of after the whole GiB.  Every step still
    # ---- uploads its 1 GiB of input lattices and reads back its observables and its 1 GiB of final lattices
    # ---- inside the timed region.
    n_handles = max(2, args.e2e_handles)
    split = max(1, args.e2e_split)
    while N_REPLICAS % split:
        split -= 1
    sub = N_REPLICAS // split
    full_host = torch.empty((N_REPLICAS, L, L), dtype=torch.int8, pin_memory=True).numpy()
    eng.download(out=full_host)
    eng.close()  # the device-resident handle is done; the end-to-end loops own their handles
    engs = []
    for k in range(n_handles):
        engs.append([bc.Engine(L, bc.PERIODIC, sub, local_rank, SEED + (rank << 40) + 64 * k + p_) for p_ in range(split)])
        for e_k in engs[-1]:
            e_k.set_params(T, D, J)

    def e2e_run(packed):
        """The step through the C ABI from pinned host buffers, `packed`: in the 2-bit wire format (four sites per
        byte, converted on the device) or as int8 spins.  Returns (seconds, h2d bytes, d2h bytes per step)."""
        hosts = []
        for k in range(n_handles):
            row_h = []
            for p_ in range(split):
                part = full_host[p_ * sub:(p_ + 1) * sub]
                if packed:
                    h = torch.empty(engs[k][p_].packed_size(), dtype=torch.uint8, pin_memory=True).numpy()
                    h[:] = bc.pack_spins(part)
                else:
                    h = torch.empty((sub, L, L), dtype=torch.int8, pin_memory=True).numpy()
                    h[:] = part
                row_h.append(h)
            hosts.append(row_h)
        obs_buf = [None]

        def steps(n):
            pending = [False] * n_handles
            for i in range(n):
                k = i % n_handles
                for p_ in range(split):
                    e = engs[k][p_]
                    if pending[k]:
                        obs_buf[0] = e.read_obs(n_meas)          # waits for step i-n_handles on this handle (d2h observables)
                    if packed:
                        e.upload_packed(hosts[k][p_], wait=False)  # h2d: the sub-batch's input lattices
                    else:
                        e.upload_async(hosts[k][p_])
                    e.sweep(SWEEPS_PER_STEP, MEASURE_EVERY, fetch=False)
                    if packed:
                        e.download_packed(out=hosts[k][p_], wait=False)  # d2h: its final lattices (the sample dump)
                    else:
                        e.download_async(hosts[k][p_])
                pending[k] = True
            for k in range(n_handles):
                for p_ in range(split):
                    if pending[k]:
                        obs_buf[0] = engs[k][p_].read_obs(n_meas)
                        engs[k][p_].sync()

        steps(max(min(args.warmup, 2), n_handles))
        barrier()
        t0 = time.perf_counter()
        steps(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        h2d = sum(h.nbytes for h in hosts[0])
        return dt, h2d, h2d + obs_buf[0].nbytes * split

    # headline end-to-end number: host buffers in the 2-bit wire format (bc_upload_packed / bc_download_packed);
    # the same loop with int8 host buffers (one byte per site each way) is reported next to it
    e2e_s, h2d_bytes, d2h_bytes = e2e_run(True)
    e2e8_s, h2d8_bytes, d2h8_bytes = e2e_run(False)
    del full_host

    for row in engs:  # free the end-to-end handles before the 4 GiB slab lattice is created
        for e_k in row:
            e_k.close()
    engs = []
    slab = None
    if dist is not None and not args.no_slab:
        try:
            slab = slab_record(bc, torch, dist, rank, world, local_rank)
        except Exception as ex:  # the record fails, the line does not
            slab = {"error": str(ex)}

    ms_t = torch.tensor([ms, e2e_s * 1e3, e2e8_s * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
    if dist is not None:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e8_ms_max = ms_t.tolist()

    if rank == 0:
        total_updates = updates_per_step * args.steps * world
        value = total_updates / (ms_max * 1e6)
        e2e_value = total_updates / (e2e_ms_max * 1e6)
        peak, peak_src = measured_peak()
        # dominant kernel = sweep_fast_kernel (one launch per half-sweep): algorithmic bytes per launch
        # = 1.5 B x sites of the batch; duration = timed region / launches (back-to-back on one stream)
        sweep_launches = 2 * SWEEPS_PER_STEP * args.steps
        bytes_per_launch = 1.5 * L * L * N_REPLICAS
        avg_launch_ms = ms / max(launches, 1)
        achieved = bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
        # DRAM bytes per launch from the tracked ncu --set full capture of this workload: a STORED number (a profiler
        # cannot run inside the timed region), labelled as such in traffic_source
        traffic = traffic_src = None
        prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
        if os.path.exists(prof):
            try:
                pj = json.load(open(prof))
                traffic = pj.get("dram_bytes_per_launch")
                traffic_src = (f"stored: profiles/ncu_summary.json ({pj.get('source', 'ncu --set full')}, "
                               f"kernel {pj.get('kernel', '?')}), not measured in this run")
            except Exception:
                traffic = None
        # bounded CPU sample of the reference on this box's host cores
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            a, w, kind, sample = reference_sample(1, 32, 2)
            cpu = {"value": a / w / 1e9, "unit": "spin-updates/ns", "cores": 1, "kind": kind, "sample": sample,
                   "cpu_model": cpu_model(), "host_cores": os.cpu_count()}
            try:  # (ii) of BASELINE.md 4.3: the attempt loop alone, one thread
                al, wl = loop_only_sample(1)
                cpu["loop_only"] = {"value": al / wl / 1e9, "unit": "spin-updates/ns", "cores": 1, "kind": "port",
                                    "sample": "3e7 attempts of the oracle's literal main.cpp:89-112 loop, L=64"}
            except Exception as ex:
                cpu["loop_only"] = {"error": str(ex)}
        N = L * L
        line = {
            "metric": "spin-updates/ns", "value": value, "unit": "spin-updates/ns", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": "spin-updates/ns",
                    "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes),
                    "handles": n_handles, "sub_batches_per_step": split,
                    "host_format": "2-bit wire format, four sites per byte (bc_upload_packed / bc_download_packed; "
                                   "converted on the device); bytes are per GPU",
                    "int8_host_buffers": {"value": total_updates / (e2e8_ms_max * 1e6), "unit": "spin-updates/ns",
                                          "h2d_bytes_per_step": int(h2d8_bytes), "d2h_bytes_per_step": int(d2h8_bytes),
                                          "host_format": "int8 spins, the reference layout (bc_upload_all_async / "
                                                         "bc_download_all_async)"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "bc::sweep_fast_kernel",
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": avg_launch_ms, "sweep_launches": sweep_launches},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "single_lattice": {"workload": f"one L={L} lattice (16 MiB, L2-resident; CUDA-graph replay + PDL)",
                               "updates_per_ns": single},
            "other_configs": others,
            "slab": slab,
            "check": {"abs_m_last": float(np.abs(last_obs[-1]["M"]).mean() / N),
                      "q_last": float(last_obs[-1]["Q"].mean() / N), "ties": ties_resident},
        }
        emit(line)
    for row in engs:
        for e_k in row:
            e_k.close()
    if dist is not None:
        dist.destroy_process_group()


_RESULT_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private copy of fd 1 for it and point fd 1 at stderr, so
    whatever libraries print there (NCCL's version banner at NCCL_DEBUG=VERSION, build output) cannot mix in."""
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l64", action="store_true", help="reference arm: skip the two-minute nominal L=64 run")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the config-4 slab record (L=65536 over the ranks)")
    ap.add_argument("--e2e-handles", type=int, default=3, help="steps in flight in the end-to-end loop")
    ap.add_argument("--e2e-split", type=int, default=4, help="sub-batches (handles) a step's replicas travel as")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not os.path.exists(os.path.join(ROOT, "blume_capel_b200", "libbc_engine.so")) and rank == 0:
        import __graft_entry__  # built artefacts are git-ignored: build once if the checkout has none
        __graft_entry__.build()
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
