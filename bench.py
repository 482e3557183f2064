#!/usr/bin/env python
"""bench.py — particle-steps/s of the device mcParticleCloud::evolve on the
synthetic 3-D box of BASELINE.json configs[4] / SURVEY.md §8(d).

  python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun)
  python bench.py --impl reference ...                      CPU arm (the oracle on all host cores)
  python bench.py --config SandiaD|SydneyBluffBodyFlame|SydneyBluffBodyFlow|SandiaPropaneJet4x
                                                            the tutorial-shaped configs of BASELINE.json configs[0..3]
  python bench.py --scaling strong --gpus N                 fixed total of 2.5e8 particles split over N GPUs

Workload (weak scaling): 200x100x100 hexes = 2.0e6 cells (4.8e7 tets), 64
particles per cell PER GPU (1.28e8 particles per GPU, 1.02e9 at 8 GPUs),
particles sharded across the GPUs, mesh/fields replicated, one NCCL all-reduce
of the per-cell moment block per step. One "step" = one evolve() of all
particles. Inputs are much larger than L2 (>= 25 GB of particle state per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAFFIC_PROFILE = "r02_k_evolve_final_metrics.csv"   # ncu --set full capture of the shipped k_evolve_jit (profiles/README.md)
METRIC = "particle-steps/sec"
UNIT = "particle-steps/s"


def algorithmic_bytes_per_particle_step(n_phi: int) -> int:
    """every persistent column read once and written once:
    pos(3)+U(3)+m+Omega+rho+eta+Phi[nPhi] doubles + cell + tet int32."""
    return 2 * (8 * (10 + n_phi) + 8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # NVML directly (a sample every 20 ms); nvidia-smi as the fallback (0.3 s per sample)
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.device)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.rows.append([str(sm), str(mx), "0"] + ["Active" if r & bits[n] else "Not Active"
                                                             for n in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
                self._stop.wait(0.02)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ppc_per_gpu(args, n_ranks):
    """particles per cell on each GPU: fixed (weak scaling) or the fixed total split over the ranks (strong)"""
    if args.scaling == "strong":
        return max(1, int(round(args.ppc_total / n_ranks)))
    if args.config != "synthetic" and args.ppc_per_gpu <= 0:
        return 0                       # the tutorial's own particlesPerCell
    return args.ppc_per_gpu if args.ppc_per_gpu > 0 else 64


def build_case(args, n_ranks):
    """The case at full size: every rank holds the whole mesh and its share of each cell's particles."""
    from pdffoam_b200 import cases
    ppc = ppc_per_gpu(args, n_ranks)
    if args.config == "synthetic":
        nx, ny, nz = args.cells
        return cases.synthetic_box(nx, ny, nz, ppc=ppc * n_ranks, closed=False, number_control=True,
                                   reaction="cold", lts="cell", poscorr="limitedSimple")
    case = cases.tutorial_case(args.config)
    if ppc > 0 or n_ranks > 1:
        case.params.particlesPerCell = (ppc if ppc > 0 else case.params.particlesPerCell) * n_ranks
    return case


def workload_text(args, case, n_ranks, per_gpu):
    P = case.params
    names = {v: k for k, v in __import__("pdffoam_b200.cases", fromlist=["x"]).REACTION_MODELS.items()}
    mix = {0: "none", 1: "IEM", 2: "modifiedCurl"}[P.mixingModel]
    models = f"SLMFull+{mix}+RAS+{names[P.reactionModel]}, cell LTS, limitedSimple, number control"
    if args.config == "synthetic":
        nx, ny, nz = args.cells
        return (f"synthetic 3D hex-to-tet box {nx}x{ny}x{nz} = {case.mesh.n_cells} cells, {P.particlesPerCell // n_ranks} particles/cell/GPU "
                f"({per_gpu} particles per GPU), {models}, inletOutlet x-faces (BASELINE.json configs[4])")
    idx = {"SydneyBluffBodyFlow": 0, "SandiaD": 1, "SydneyBluffBodyFlame": 2, "SandiaPropaneJet4x": 3}[args.config]
    return (f"tutorial-shaped {args.config}: axisymmetric wedge, {case.mesh.n_cells} cells x {P.particlesPerCell // n_ranks} particles/cell/GPU "
            f"({per_gpu} particles per GPU), scalars {P.nPhi}, {models}, synthetic frozen FV fields (BASELINE.json configs[{idx}])")


def cpu_case(config, cells, ppc):
    """The bounded sample of the workload the CPU legs run: same fields, models and particles per cell, fewer
    cells (synthetic box) or the unrefined mesh (SandiaPropaneJet4x); the other tutorial shapes run at full size."""
    from pdffoam_b200 import cases
    if config == "synthetic":
        return cases.synthetic_box(*cells, ppc=ppc, closed=False, number_control=True)
    if config == "SandiaPropaneJet4x":
        return _propane_unrefined(ppc if ppc > 0 else 400)   # the tutorial's own resolution, same particles per cell
    return cases.tutorial_case(config, ppc=ppc if ppc > 0 else None)


def _propane_unrefined(ppc):
    from pdffoam_b200 import cases, meshgen
    IO, SLIP = meshgen.BC_INLET_OUTLET, meshgen.BC_SLIP
    return cases._streams_case(
        "SandiaPropaneJet", meshgen.graded(54, 0.0, 0.3, 6.0), [(7, 1.0e-4, 2.6e-3, 1.0), (5, 2.6e-3, 4.5e-3, 1.0), (24, 4.5e-3, 0.04, 6.0)], 2.5,
        [("jet", 2.6e-3, 69.0, 1.0, 1.8, IO), ("bluffBody", 4.5e-3, 0.0, 0.0, 1.2, SLIP), ("coflow", 1e30, 9.2, 0.0, 1.2, IO)],
        None, ("z", "T"), "BurkeSchumann",
        dict(CFL=0.1, averagingCoeff=5e3, particlesPerCell=ppc, ltsUpperBound=20.0, pcC=0.2, Cmix=2.0, mixingModel="IEM"))


def cpu_baseline(args, ppc, sample_cells=(40, 20, 20), budget_s=15.0):
    """The oracle ("port" of the reference's evolve) timed single-threaded on a
    bounded sample of the same workload (same ppc, same fields, fewer cells)."""
    from oracle.oracle import OracleLib
    lib = OracleLib()
    case = cpu_case(args.config, sample_cells, ppc)
    h = case.make_cloud(lib)
    h.seed_particles()
    h.evolve(4)
    n0 = h.num_particles()
    t0 = time.perf_counter()
    psteps, steps = 0, 0
    while steps < 3 or (time.perf_counter() - t0 < budget_s and steps < 200):
        psteps += h.evolve(1)[0].nParticles
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": psteps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle evolve, {case.name}: {case.mesh.n_cells} cells x {case.params.particlesPerCell} ppc = {n0} particles, {steps} steps, 1 thread"}


# ---------------------------------------------------------------------------
# reference arm: the oracle on all host cores (particle-sharded processes, the
# per-cell moment block summed between evolve_begin / evolve_end)
# ---------------------------------------------------------------------------
def _ref_worker(conn, config, cells, ppc, n_ranks, rank, shared_blocks, shared_total, nblk):
    import ctypes as C
    from pdffoam_b200.capi import StepReport
    from oracle.oracle import OracleLib, dp
    lib = OracleLib()
    case = cpu_case(config, cells, ppc)
    h = case.make_cloud(lib)
    lib.fn("set_rank")(h.h, n_ranks, rank)
    h.seed_particles()
    assert int(lib.fn("moment_block_size")(h.h)) == nblk
    # the per-cell moment blocks live in shared memory (inherited through fork): nothing big crosses the pipes
    mine = np.frombuffer(shared_blocks, dtype=np.float64)[rank * nblk:(rank + 1) * nblk]
    total = np.frombuffer(shared_total, dtype=np.float64)
    rep = StepReport()
    conn.send(("ready", h.num_particles()))
    while True:
        msg = conn.recv()
        if msg[0] == "begin":
            lib.check(lib.fn("evolve_begin")(h.h, C.byref(rep), dp(mine)), "evolve_begin")
            conn.send((rep.CoMax, rep.UMax))
        elif msg[0] == "end":
            rep.CoMax = msg[1]
            lib.check(lib.fn("evolve_end")(h.h, C.byref(rep), dp(total)), "evolve_end")
            conn.send(int(rep.nParticles))
        else:
            break


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle.oracle import OracleLib
    cores = os.cpu_count() or 1
    n_proc = max(1, min(cores, 64))
    cells = tuple(args.ref_cells)
    ppc = ppc_per_gpu(args, 1)
    case0 = cpu_case(args.config, cells, ppc)
    ppc_total = case0.params.particlesPerCell
    n_proc = min(n_proc, ppc_total)                     # every process holds at least one particle per cell
    h0 = case0.make_cloud(OracleLib())
    nblk = int(OracleLib().fn("moment_block_size")(h0.h))
    del h0
    ctx = mp.get_context("fork")
    shared_blocks = ctx.RawArray("d", n_proc * nblk)
    shared_total = ctx.RawArray("d", nblk)
    blocks = np.frombuffer(shared_blocks, dtype=np.float64).reshape(n_proc, nblk)
    total = np.frombuffer(shared_total, dtype=np.float64)
    conns, procs = [], []
    for r in range(n_proc):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, args.config, cells, ppc, n_proc, r, shared_blocks, shared_total, nblk), daemon=True)
        p.start()
        conns.append(a); procs.append(p)
    n0 = sum(c.recv()[1] for c in conns)

    def step():
        for c in conns:
            c.send(("begin",))
        outs = [c.recv() for c in conns]
        np.sum(blocks, axis=0, out=total)
        comax = max(o[0] for o in outs)
        for c in conns:
            c.send(("end", comax))
        return sum(c.recv() for c in conns)

    for _ in range(min(args.spinup, 4) + args.warmup):
        step()
    t0 = time.perf_counter()
    psteps = 0
    for _ in range(args.steps):
        psteps += step()
    dt = time.perf_counter() - t0
    for c in conns:
        c.send(("quit",))
    value = psteps / dt
    sample = (f"oracle evolve (CPU restatement of mcParticleCloud::evolve; pdfFoam itself needs OpenFOAM and cannot be built here), "
              f"{case0.name}: {case0.mesh.n_cells} cells x {ppc_total} ppc = {n0} particles per step, "
              f"{n_proc} particle-sharded processes, moment blocks summed in shared memory")
    full = build_case(args, 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args, full, 1, full.mesh.n_cells * full.params.particlesPerCell), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_proc, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def ncu_traffic_bytes(args, n_phi):
    """DRAM bytes of one k_evolve launch as captured by ncu for the default workload
    (profiles/TRAFFIC_PROFILE); None when the bench runs another workload."""
    if args.config != "synthetic" or args.scaling != "weak" or list(args.cells) != [200, 100, 100] or args.ppc_per_gpu != 64 or n_phi != 1:
        return None
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", TRAFFIC_PROFILE)
    try:
        tot = 0.0
        with open(path) as f:
            for row in f:
                name, unit, val = (row.strip().split(",") + ["", ""])[:3]
                if name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        return tot or None
    except (OSError, ValueError, KeyError):
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pdffoam_b200.capi import DeviceLibrary
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = DeviceLibrary()
    case = build_case(args, world)
    n_phi = case.params.nPhi
    ppc_gpu = case.params.particlesPerCell // world
    per_gpu = case.mesh.n_cells * ppc_gpu
    h = case.make_cloud(lib, device=local_rank, capacity=int(per_gpu * 1.2) + 65536)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = C.create_string_buffer(128)
            lib.check(lib.fn("comm_unique_id")(buf), "comm_unique_id")
            uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        h.comm_init(world, rank, bytes(uid.cpu().tolist()))
    h.seed_particles()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    # spin-up (part of the set-up, like seeding): the freshly seeded cloud starts with
    # deltaT = 1e-13 and needs a few steps before deltaT = CFL/CoMax, eta and the particle
    # distribution are statistically stationary; then the W untimed warm-up steps
    h.evolve(args.spinup)
    h.evolve(max(args.warmup, 3))
    h.profile(reset=True)
    launches0 = h.kernel_launches()

    # ---- device-resident timing: K steps, CUDA events on the library's stream ----
    # inputs larger than L2 (the default workload: 12 GB of particle state) need no flush; the small tutorial shapes
    # fit the 126 MB L2, so there every timed step is preceded by a write of a 512 MB buffer (not timed)
    state_bytes = per_gpu * (algorithmic_bytes_per_particle_step(n_phi) // 2)
    flush = state_bytes < 1e9
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if flush else None
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        if not flush:
            reps = h.evolve(args.steps)
            dev_ms = h.last_evolve_ms()
        else:
            reps, dev_ms = [], 0.0
            for _ in range(args.steps):
                flush_buf.fill_(1)
                torch.cuda.synchronize()
                reps += h.evolve(1)
                dev_ms += h.last_evolve_ms()
        barrier()
        wall = time.perf_counter() - t0
    prof = h.profile(reset=True)
    launches = h.kernel_launches() - launches0
    breakdown_note = None
    if prof["steps"] == 0:
        # the timed steps were replayed from a CUDA graph (small cloud): no per-region times there, so the breakdown
        # comes from a few plain-launch steps taken afterwards (not part of the timed region)
        h.set_graph_mode(0)
        h.evolve(min(args.steps, 10))
        prof = h.profile(reset=True)
        h.set_graph_mode(-1)
        breakdown_note = "timed steps replayed from a CUDA graph; breakdown and roofline from plain-launch steps taken after the timed region"
    psteps_local = float(sum(r.nParticles for r in reps))
    ar_ms = float(sum(r.momentsAllreduceMs for r in reps))

    # ---- end-to-end through the C ABI with HOST buffers: upload the FV fields,
    #      evolve one step, read rho back (what pdfSimpleFoam does per PDF cycle) ----
    #      Host buffers are pinned; the read-back is rho, the only field fed back to the FV
    #      equations (pdfSimpleFoam/createFields.H:18).
    def pinned_like(a):
        t = torch.empty(a.shape, dtype=torch.float64 if a.dtype == np.float64 else torch.uint8).pin_memory()
        v = t.numpy()
        v[...] = a
        return t, v
    keep, fv_pinned = [], {}
    for k_, a in case.fv.items():
        t, v = pinned_like(np.ascontiguousarray(a))
        keep.append(t); fv_pinned[k_] = v
    rho_t = torch.empty(case.mesh.n_cells, dtype=torch.float64).pin_memory()
    rho_host = rho_t.numpy()
    # Every step's FV state is copied from pinned host memory inside the timed region; the copy of step n+1 is staged
    # (pdfb200_stage_fv_fields: own stream, second set of device buffers) before step n is evolved, so that it runs
    # under the particle kernels, and committed (pointer swap) before step n+1. Two host buffer sets alternate.
    # Every step's rho is read back into pinned host memory inside the timed region as well
    # (pdfb200_download_cell_fields_async, queued behind the step on a stream of its own); the host waits for the
    # rho of step n and reads it after step n+1 has been issued, so the read-back runs under the next step.
    fv_sets = [fv_pinned]
    fv2 = {}
    for k_, a in fv_pinned.items():
        t, v = pinned_like(a)
        keep.append(t); fv2[k_] = v
    fv_sets.append(fv2)
    rho_t2 = torch.empty(case.mesh.n_cells, dtype=torch.float64).pin_memory()
    rho_bufs = [rho_host, rho_t2.numpy()]
    rho_check = 0.0
    barrier()
    t0 = time.perf_counter()
    e2e_psteps = 0.0
    h.stage_fv_fields(fv_sets[0])
    for n_ in range(args.steps):
        h.commit_fv_fields()
        if n_ + 1 < args.steps:
            h.stage_fv_fields(fv_sets[(n_ + 1) % 2])
        r = h.evolve(1)[0]
        if n_ > 0:
            h.sync_downloads()
            rho_check += float(rho_bufs[(n_ - 1) % 2][0])      # the host consumes the previous step's field
        h.download_rho_async(rho_bufs[n_ % 2])
        e2e_psteps += r.nParticles
    h.sync_downloads()
    rho_check += float(rho_bufs[(args.steps - 1) % 2][0])
    barrier()
    e2e_wall = time.perf_counter() - t0
    nc, nb = case.mesh.n_cells, case.mesh.n_boundary_faces
    h2d = 8 * (nc * (3 + 1 + 1 + 1 + 6) + nb * (3 + 1 + 1 + 1 + 6)) + 2 * nb
    d2h = 8 * nc

    # ---- the same measurement on the stationary cloud (default workload only; reported beside the headline) ----
    # The timed region above sees the cloud 13 steps after seeding (as in round 1: every cell still holds close to
    # its 64 particles). After a few hundred steps number control clones and eliminates ~4 % of the particles per
    # step, the populations per cell spread, and a step costs more; a long run lives there, so it is measured too.
    st_ms, st_psteps, st_prof, st_steps_before = 0.0, 0.0, None, 0
    if args.config == "synthetic" and args.stationary_spinup > 0:
        h.evolve(args.stationary_spinup)
        st_steps_before = h.step_counter()
        h.profile(reset=True)
        barrier()
        st_reps = h.evolve(args.steps)
        st_ms = h.last_evolve_ms()
        barrier()
        st_prof = h.profile(reset=True)
        st_psteps = float(sum(r.nParticles for r in st_reps))

    rank_times = None
    if world > 1 and st_prof is not None:
        # per-rank view of the stationary step: device time and the regions' times (the all-reduce region contains the
        # wait for the slowest rank)
        mine = torch.tensor([st_ms / max(args.steps, 1)] + [st_prof[k] / max(st_prof["steps"], 1) for k in
                            ("prep_ms", "evolve_kernel_ms", "sort_ms", "moments_ms", "allreduce_ms", "finalise_ms")], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_times = [[round(float(x), 2) for x in t.tolist()] for t in allr]
    vals = torch.tensor([dev_ms, wall, e2e_wall, st_ms], dtype=torch.float64, device="cuda")
    sums = torch.tensor([psteps_local, e2e_psteps, st_psteps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, wall, e2e_wall, st_ms = vals.tolist()
    psteps, e2e_psteps, st_psteps = sums.tolist()

    if rank == 0:
        value = psteps / (dev_ms * 1e-3)
        peak, peak_src = measured_hbm_peak()
        bpp = algorithmic_bytes_per_particle_step(n_phi)
        k_ms = prof["evolve_kernel_ms"] / max(prof["steps"], 1)
        k_particles = prof["particle_steps"] / max(prof["steps"], 1)
        achieved = bpp * k_particles / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_text(args, case, world, per_gpu),
                       "parallelism": f"particle-sharded x{world}, moment all-reduce", "n_phi": n_phi,
                       "l2_policy": ("inputs >> L2 (particle state %.1f GB per GPU)" % (state_bytes / 1e9)) if not flush
                       else ("L2 flushed before every timed step (512 MB written); particle state %.3f GB per GPU" % (state_bytes / 1e9))},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "allreduce_ms_per_step": ar_ms / args.steps,
            "hbm_GBps_whole_step": bpp * psteps / world / (dev_ms * 1e-3) / 1e9,
            "breakdown_ms_per_step": {k: v / max(prof["steps"], 1) for k, v in prof.items() if k.endswith("_ms")},
            "roofline": {"bound": "hbm", "kernel": "k_evolve", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(args, n_phi), "peak_source": peak_src,
                         "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one k_evolve "
                                           f"launch of this workload (profiles/{TRAFFIC_PROFILE}); null for other workloads",
                         "algorithmic_bytes_per_launch": bpp * k_particles,
                         "algorithmic_bytes_per_particle_step": bpp},
            "e2e": {"value": e2e_psteps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "pipeline": "FV fields of step n+1 staged under step n; rho of step n read back (pinned host buffer) under step n+1"},
            "gpu_launches": int(launches),
            "step_issue": breakdown_note or "plain launches, reports fetched once per batch of 64 steps",
            "clocks": clocks.summary(),
        }
        if st_prof is not None and st_ms > 0:
            st_k = st_prof["evolve_kernel_ms"] / max(st_prof["steps"], 1)
            st_kp = st_prof["particle_steps"] / max(st_prof["steps"], 1)
            line["stationary_state"] = {
                "value": st_psteps / (st_ms * 1e-3), "unit": UNIT, "ms_per_step": st_ms / args.steps, "steps_before": int(st_steps_before),
                "breakdown_ms_per_step": {k: v / max(st_prof["steps"], 1) for k, v in st_prof.items() if k.endswith("_ms")},
                "roofline_frac": (bpp * st_kp / (st_k * 1e-3) / 1e9 / peak) if st_k > 0 else None,
                "note": "same workload and code, timed after the cloud has become statistically stationary (number control active in a third of the cells)"}
            if rank_times is not None:
                line["stationary_state"]["per_rank_ms"] = {"columns": ["step", "prep", "evolve_kernel", "sort", "moments", "allreduce_incl_wait", "finalise"], "rows": rank_times}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, ppc_per_gpu(args, 1))
            except Exception as e:  # the baseline must never hide the GPU number
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, nargs=3, default=[200, 100, 100])
    ap.add_argument("--ppc-per-gpu", type=int, default=-1, help="particles per cell and GPU (default: 64 for the synthetic box, the tutorial's own value otherwise)")
    ap.add_argument("--config", default="synthetic", choices=["synthetic", "SydneyBluffBodyFlow", "SandiaD", "SydneyBluffBodyFlame", "SandiaPropaneJet4x"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--ppc-total", type=int, default=125, help="--scaling strong: particles per cell in total (125 x 2e6 cells = 2.5e8)")
    ap.add_argument("--ref-cells", type=int, nargs=3, default=[48, 24, 24])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spinup", type=int, default=10, help="set-up steps before the warm-up (not timed)")
    ap.add_argument("--stationary-spinup", type=int, default=360, help="further untimed steps before the second, stationary-state measurement of the default workload (0: skip it)")
    args = ap.parse_args()
    if args.config == "synthetic" and args.ppc_per_gpu <= 0:
        args.ppc_per_gpu = 64
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
