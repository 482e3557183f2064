import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
from oracle.oracle import OracleLib
N=24; STEPS=25
def run(lib, seed, mode):
    case = cases.synthetic_box(8, 5, 4, ppc=30, closed=True, number_control=True)
    case.params.seed = seed; case.params.rngMode = mode; case.params.deltaT0 = 2e-3; case.params.averagingCoeff = 5.0
    h = case.make_cloud(lib); h.seed_particles(); h.evolve(STEPS)
    c = h.download_cell_fields()
    return dict(Ux=c["U"][:, 0], k=c["k"], z=c["Phi"][0], zz=c["Cov"][0], rho_err=(c["pnd"] - c["rho"]) / c["rho"], N=c["PaNIC"])
dev = [run(DeviceLibrary(), 100+s, 0) for s in range(N)]
ok = [run(OracleLib(), 900+s, 0) for s in range(N)]
os_ = [run(OracleLib(), 500+s, 1) for s in range(N)]
for f in ("Ux","k","z","zz","rho_err","N"):
    A=np.stack([d[f] for d in dev]); B=np.stack([d[f] for d in ok]); Cc=np.stack([d[f] for d in os_])
    print(f, "mean dev/okey/oseq %.5g %.5g %.5g" % (A.mean(), B.mean(), Cc.mean()), " ens-var(mean over cells) %.4g %.4g %.4g" % (A.var(0,ddof=1).mean(), B.var(0,ddof=1).mean(), Cc.var(0,ddof=1).mean()),
          " var of domain-avg %.4g %.4g %.4g" % (A.mean(1).var(ddof=1), B.mean(1).var(ddof=1), Cc.mean(1).var(ddof=1)))
