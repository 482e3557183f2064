import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from pdffoam_b200 import cases
from oracle.oracle import OracleLib
N=48; STEPS=25
def run(lib, seed, mode):
    case = cases.synthetic_box(8, 5, 4, ppc=30, closed=False, number_control=True)
    case.params.seed = seed; case.params.rngMode = mode; case.params.deltaT0 = 2e-3; case.params.averagingCoeff = 5.0
    h = case.make_cloud(lib); h.seed_particles(); reps = h.evolve(STEPS)
    c = h.download_cell_fields()
    return dict(k=c["k"], N=c["PaNIC"], UMax=max(r.UMax for r in reps), inj=sum(r.nInjected for r in reps), dele=sum(r.nDeleted for r in reps), dt=np.mean([r.deltaT for r in reps[5:]]))
ok = [run(OracleLib(), 900+s, 0) for s in range(N)]
os_ = [run(OracleLib(), 500+s, 1) for s in range(N)]
for name, e in (("keyed", ok), ("seq", os_)):
    K = np.stack([d["k"] for d in e])
    print(name, "k mean %.3f" % K.mean(), "se %.3f" % (K.mean(1).std(ddof=1)/np.sqrt(N)), "median of domain avg %.3f" % np.median(K.mean(1)), "max cell k %.1f" % K.max(),
          "UMax mean %.1f" % np.mean([d["UMax"] for d in e]), "inj %.1f del %.1f dt %.3e" % (np.mean([d["inj"] for d in e]), np.mean([d["dele"] for d in e]), np.mean([d["dt"] for d in e])))
    print("   per-x-slab k:", np.round(K.mean(0).reshape(4,5,8).mean((0,1)),2))
