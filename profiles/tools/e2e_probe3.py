"""per-phase host times of the e2e loop under torchrun (N ranks)"""
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ctypes as C
import torch.distributed as dist
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = DeviceLibrary()
case = cases.synthetic_box(200, 100, 100, ppc=64 * world, closed=False, number_control=True)
h = case.make_cloud(lib, device=lr, capacity=int(case.mesh.n_cells * 64 * 1.2) + 65536)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = C.create_string_buffer(128); lib.check(lib.fn("comm_unique_id")(buf), "uid"); uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0); h.comm_init(world, rank, bytes(uid.cpu().tolist()))
h.seed_particles(); h.evolve(13)
def pinned_like(a):
    t = torch.empty(a.shape, dtype=torch.float64 if a.dtype == np.float64 else torch.uint8).pin_memory(); v = t.numpy(); v[...] = a; return t, v
keep = []; sets = []
for _ in range(2):
    d = {}
    for k, a in case.fv.items():
        t, v = pinned_like(np.ascontiguousarray(a)); keep.append(t); d[k] = v
    sets.append(d)
rho = torch.empty(case.mesh.n_cells, dtype=torch.float64).pin_memory(); rho_h = rho.numpy()
for mode in ("plain", "e2e", "e2e_nodl", "e2e_nostage"):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    ts = []
    if mode.startswith("e2e") and mode != "e2e_nostage": h.stage_fv_fields(sets[0])
    T0 = time.perf_counter()
    for n in range(10):
        t0 = time.perf_counter()
        if mode.startswith("e2e") and mode != "e2e_nostage":
            h.commit_fv_fields(); t1 = time.perf_counter()
            if n < 9: h.stage_fv_fields(sets[(n + 1) % 2])
        else: t1 = t0
        t2 = time.perf_counter()
        h.evolve(1); t3 = time.perf_counter()
        if mode in ("e2e", "e2e_nostage"): h.download_rho(rho_h)
        t4 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, h.last_evolve_ms() * 1e-3))
    torch.cuda.synchronize(); T1 = time.perf_counter()
    a = np.array(ts[1:]) * 1e3
    print(f"rank {rank} {mode}: wall/step {1e2*(T1-T0):.2f} ms; commit {a[:,0].mean():.2f} stage {a[:,1].mean():.2f} evolve {a[:,2].mean():.2f} (dev {a[:,4].mean():.2f}) download {a[:,3].mean():.2f}", flush=True)
if world > 1: dist.destroy_process_group()
