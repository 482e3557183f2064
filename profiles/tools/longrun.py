"""Long run of the bench workload on one GPU: N steps in batches of 100, one line per batch (population, time step,
mass balance, kernel time) - evidence that the step is stable far beyond the few steps a bench times."""
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import numpy as np
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
lib = DeviceLibrary()
case = cases.synthetic_box(200, 100, 100, ppc=64, closed=False, number_control=True)
h = case.make_cloud(lib, device=0, capacity=int(case.mesh.n_cells * 64 * 1.2) + 65536)
h.seed_particles()
# ids start just below 2^32 so that the run crosses the 32-bit boundary
h.set_id_counter((1 << 32) - 50_000_000)
out = []
t0 = time.perf_counter()
for blk in range(n_steps // 100):
    h.profile(reset=True)
    reps = h.evolve(100)
    pr = h.profile(reset=True)
    r = reps[-1]
    row = dict(step=100 * (blk + 1), wall_s=round(time.perf_counter() - t0, 2), ms_per_step=round(h.last_evolve_ms() / 100, 3),
               k_evolve_ms=round(pr["evolve_kernel_ms"] / 100, 3), n=int(r.nParticles), dT=r.deltaT, CoMax=r.CoMax, UMax=r.UMax,
               etaMin=r.etaMin, etaMax=r.etaMax, rhoRes=r.rhoRes, totalMass=r.totalMass, totalParticleMass=r.totalParticleMass,
               injected=int(sum(x.nInjected for x in reps)), deleted=int(sum(x.nDeleted for x in reps)),
               cloned=int(sum(x.nCloned for x in reps)), eliminated=int(sum(x.nEliminated for x in reps)), lost=int(sum(x.nLost for x in reps)),
               id_counter=h.id_counter())
    out.append(row)
    print(json.dumps(row), flush=True)
P = h.download_particles()
ids = P["id"]
print(json.dumps(dict(final_particles=int(ids.size), unique_ids=int(np.unique(ids).size), max_id=int(ids.max()), ids_above_2_32=int((ids >= (1 << 32)).sum()),
                      finite=bool(np.isfinite(P["pos"]).all() and np.isfinite(P["U"]).all() and np.isfinite(P["m"]).all()),
                      m_min=float(P["m"].min()), m_max=float(P["m"].max()))), flush=True)
