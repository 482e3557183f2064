"""Fresh cloud per variant (same state for every variant): seed, 10 spin-up + 3 warm-up steps, STEPS timed.
usage: python scratch/variants2.py "name|flags|blocks" ...   env CELLS PPC STEPS; name 'drift' prints a per-step trace"""
import os, sys, time, json
sys.path.insert(0, os.getcwd())   # run from the repository root
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
cells = [int(x) for x in os.environ.get("CELLS", "200 100 100").split()]
ppc = int(os.environ.get("PPC", "64")); steps = int(os.environ.get("STEPS", "6"))
lib = DeviceLibrary()
case = cases.synthetic_box(*cells, ppc=ppc, closed=False, number_control=True)
for spec in sys.argv[1:]:
    name, flags, blocks = (spec.split("|") + ["", ""])[:3]
    if flags: os.environ["PDFB200_JIT_FLAGS"] = flags
    else: os.environ.pop("PDFB200_JIT_FLAGS", None)
    if blocks: os.environ["PDFB200_K1_BLOCKS"] = blocks
    else: os.environ.pop("PDFB200_K1_BLOCKS", None)
    h = case.make_cloud(lib, device=0, capacity=int(case.mesh.n_cells * ppc * 1.2) + 65536)
    h.seed_particles()
    if name.startswith("drift"):
        for blk in range(int(os.environ.get("DRIFT_BLOCKS", "16"))):
            h.profile(reset=True)
            reps = h.evolve(5); pr = h.profile(reset=True)
            r = reps[-1]
            print(json.dumps(dict(step=5 * (blk + 1), k_ms=round(pr["evolve_kernel_ms"] / 5, 3), n=int(r.nParticles), dT=r.deltaT, CoMax=r.CoMax,
                                  etaMax=r.etaMax, UMax=r.UMax, inj=int(r.nInjected), dele=int(r.nDeleted), cl=int(r.nCloned), el=int(r.nEliminated))), flush=True)
        del h
        continue
    h.evolve(int(os.environ.get("SPINUP", "10"))); h.evolve(3); h.profile(reset=True)
    reps = h.evolve(steps); ms = h.last_evolve_ms() / steps
    pr = h.profile(reset=True)
    row = dict(name=name, ms_per_step=round(ms, 3), evolve_kernel_ms=round(pr["evolve_kernel_ms"] / steps, 3),
               particles=int(reps[-1].nParticles),
               rest={k: round(v / steps, 2) for k, v in pr.items() if k.endswith("_ms") and k != "evolve_kernel_ms"})
    print(json.dumps(row), flush=True)
    del h
