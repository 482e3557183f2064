import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ctypes as C
import torch.distributed as dist
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
lib = DeviceLibrary()
case = cases.synthetic_box(200, 100, 100, ppc=64 * world, closed=False, number_control=True)
h = case.make_cloud(lib, device=lr, capacity=int(case.mesh.n_cells * 64 * 1.3) + 65536)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = C.create_string_buffer(128); lib.check(lib.fn("comm_unique_id")(buf), "uid"); uid = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0); h.comm_init(world, rank, bytes(uid.cpu().tolist()))
h.seed_particles()
for blk in range(8):
    reps = h.evolve(50)
    pr = h.profile(reset=True)
    t = torch.tensor([float(h.num_particles()), sum(r.nCloned for r in reps) / 50, sum(r.nEliminated for r in reps) / 50, sum(r.nInjected for r in reps) / 50,
                      sum(r.nDeleted for r in reps) / 50, pr["evolve_kernel_ms"] / 50, pr["finalise_ms"] / 50], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        print("step", 50 * (blk + 1), "global n", int(reps[-1].nParticles), flush=True)
        for r_, a in enumerate(allr):
            a = a.tolist()
            print("   rank %d local n %.0f cloned %.0f elim %.0f inj %.0f del %.0f evolve %.2f finalise %.2f" % (r_, *a), flush=True)
dist.destroy_process_group()
