#!/usr/bin/env python
"""bench.py — particle-steps/s of the device mcParticleCloud::evolve on the
synthetic 3-D box of BASELINE.json configs[4] / SURVEY.md §8(d).

  python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun)
  python bench.py --impl reference ...                      CPU arm (the oracle on all host cores)

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


def build_case(args, n_ranks):
    from pdffoam_b200 import cases
    nx, ny, nz = args.cells
    case = cases.synthetic_box(nx, ny, nz, ppc=args.ppc_per_gpu * n_ranks, closed=False, number_control=True,
                               reaction="cold", lts="cell", poscorr="limitedSimple")
    return case


def cpu_baseline(args, sample_cells=(40, 20, 20), steps=10):
    """The oracle ("port" of the reference's evolve) timed single-threaded on a
    bounded sample of the same workload (same ppc, same fields, fewer cells)."""
    from pdffoam_b200 import cases
    from oracle.oracle import OracleLib
    lib = OracleLib()
    nx, ny, nz = sample_cells
    case = cases.synthetic_box(nx, ny, nz, ppc=args.ppc_per_gpu, closed=False, number_control=True)
    h = case.make_cloud(lib)
    h.seed_particles()
    h.evolve(4)
    n0 = h.num_particles()
    t0 = time.perf_counter()
    reps = h.evolve(steps)
    dt = time.perf_counter() - t0
    psteps = sum(r.nParticles for r in reps)
    return {"value": psteps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle evolve, {nx}x{ny}x{nz} cells x {args.ppc_per_gpu} ppc = {n0} particles, {steps} steps, 1 thread"}


# ---------------------------------------------------------------------------
# reference arm: the oracle on all host cores (particle-sharded processes, the
# per-cell moment block summed between evolve_begin / evolve_end)
# ---------------------------------------------------------------------------
def _ref_worker(conn, cells, ppc, n_ranks, rank):
    import ctypes as C
    from pdffoam_b200 import cases
    from pdffoam_b200.capi import StepReport
    from oracle.oracle import OracleLib, dp
    lib = OracleLib()
    case = cases.synthetic_box(*cells, ppc=ppc, closed=False, number_control=True)
    h = case.make_cloud(lib)
    lib.fn("set_rank")(h.h, n_ranks, rank)
    h.seed_particles()
    nblk = int(lib.fn("moment_block_size")(h.h))
    block = np.zeros(nblk)
    rep = StepReport()
    conn.send(("ready", h.num_particles()))
    while True:
        msg = conn.recv()
        if msg[0] == "begin":
            lib.check(lib.fn("evolve_begin")(h.h, C.byref(rep), dp(block)), "evolve_begin")
            conn.send((block.copy(), rep.CoMax, rep.UMax))
        elif msg[0] == "end":
            block[:] = msg[1]
            rep.CoMax = msg[2]
            lib.check(lib.fn("evolve_end")(h.h, C.byref(rep), dp(block)), "evolve_end")
            conn.send(int(rep.nParticles))
        else:
            break


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_proc = max(1, min(cores, 64))
    cells = tuple(args.ref_cells)
    ctx = mp.get_context("fork")
    conns, procs = [], []
    for r in range(n_proc):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, cells, args.ppc_per_gpu, n_proc, r), daemon=True)
        p.start()
        conns.append(a); procs.append(p)
    n0 = sum(c.recv()[1] for c in conns)

    def step():
        for c in conns:
            c.send(("begin",))
        outs = [c.recv() for c in conns]
        block = np.sum([o[0] for o in outs], axis=0)
        comax = max(o[1] for o in outs)
        for c in conns:
            c.send(("end", block, comax))
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
              f"{cells[0]}x{cells[1]}x{cells[2]} cells x {args.ppc_per_gpu} ppc = {n0} particles per step, "
              f"{n_proc} particle-sharded processes")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic 3D hex-to-tet box {args.cells[0]}x{args.cells[1]}x{args.cells[2]} cells, "
                                   f"{args.ppc_per_gpu} particles/cell/GPU (BASELINE.json configs[4])", "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_proc, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def ncu_traffic_bytes(args, n_phi):
    """DRAM bytes of one k_evolve launch as captured by ncu for the default workload
    (profiles/r01_k_evolve_jit_metrics.csv); None when the bench runs another workload."""
    if list(args.cells) != [200, 100, 100] or args.ppc_per_gpu != 64 or n_phi != 1:
        return None
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_k_evolve_jit_metrics.csv")
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
    per_gpu = case.mesh.n_cells * args.ppc_per_gpu
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
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        reps = h.evolve(args.steps)
        barrier()
        wall = time.perf_counter() - t0
    dev_ms = h.last_evolve_ms()
    prof = h.profile(reset=True)
    launches = h.kernel_launches() - launches0
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
    barrier()
    t0 = time.perf_counter()
    e2e_psteps = 0.0
    for _ in range(args.steps):
        h.upload_fv_fields(fv_pinned)
        r = h.evolve(1)[0]
        h.download_rho(rho_host)
        e2e_psteps += r.nParticles
    barrier()
    e2e_wall = time.perf_counter() - t0
    nc, nb = case.mesh.n_cells, case.mesh.n_boundary_faces
    h2d = 8 * (nc * (3 + 1 + 1 + 1 + 6) + nb * (3 + 1 + 1 + 1 + 6)) + 2 * nb
    d2h = 8 * nc

    vals = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device="cuda")
    sums = torch.tensor([psteps_local, e2e_psteps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms, wall, e2e_wall = vals.tolist()
    psteps, e2e_psteps = sums.tolist()

    if rank == 0:
        value = psteps / (dev_ms * 1e-3)
        peak, peak_src = measured_hbm_peak()
        bpp = algorithmic_bytes_per_particle_step(n_phi)
        k_ms = prof["evolve_kernel_ms"] / max(prof["steps"], 1)
        k_particles = prof["particle_steps"] / max(prof["steps"], 1)
        achieved = bpp * k_particles / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic 3D hex-to-tet box {args.cells[0]}x{args.cells[1]}x{args.cells[2]} = {nc} cells, "
                                   f"{args.ppc_per_gpu} particles/cell/GPU ({per_gpu} particles per GPU), SLMFull+IEM+RAS+cold, "
                                   f"cell LTS, limitedSimple, number control, inletOutlet x-faces (BASELINE.json configs[4])",
                       "parallelism": f"particle-sharded x{world}, moment all-reduce", "n_phi": n_phi,
                       "l2_policy": "inputs >> L2 (particle state %.1f GB per GPU)" % (per_gpu * (bpp / 2) / 1e9)},
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "allreduce_ms_per_step": ar_ms / args.steps,
            "hbm_GBps_whole_step": bpp * psteps / world / (dev_ms * 1e-3) / 1e9,
            "breakdown_ms_per_step": {k: v / max(prof["steps"], 1) for k, v in prof.items() if k.endswith("_ms")},
            "roofline": {"bound": "hbm", "kernel": "k_evolve", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(args, n_phi), "peak_source": peak_src,
                         "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one k_evolve "
                                           "launch of this workload (profiles/r01_k_evolve_jit_metrics.csv); null for other workloads",
                         "algorithmic_bytes_per_launch": bpp * k_particles,
                         "algorithmic_bytes_per_particle_step": bpp},
            "e2e": {"value": e2e_psteps / e2e_wall, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args)
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
    ap.add_argument("--ppc-per-gpu", type=int, default=64)
    ap.add_argument("--ref-cells", type=int, nargs=3, default=[48, 24, 24])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spinup", type=int, default=10, help="set-up steps before the warm-up (not timed)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
