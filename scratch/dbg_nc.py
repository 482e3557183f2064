import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from pdffoam_b200 import cases
from pdffoam_b200.capi import DeviceLibrary
from oracle.oracle import OracleLib
from parity_util import sort_by_id
case = cases.synthetic_box(6, 4, 4, ppc=20, closed=True, number_control=True); case.params.deltaT0 = 2e-3
P = case.random_particles(20, seed=9)
keep = np.ones(P["m"].shape[0], bool)
rng = np.random.default_rng(0)
for c in range(0, case.mesh.n_cells, 5):
    idx = np.nonzero(P["cell"] == c)[0]
    keep[idx[rng.random(idx.size) < 0.6]] = False
Q = {k: (v[keep] if k != "Phi" else [a[keep] for a in v]) for k, v in P.items()}
dup = np.nonzero(np.isin(Q["cell"], np.arange(1, case.mesh.n_cells, 7)))[0]
R = {k: (np.concatenate([v, v[dup]]) if k != "Phi" else [np.concatenate([a, a[dup]]) for a in v]) for k, v in Q.items()}
R["id"] = np.arange(R["m"].shape[0], dtype=np.uint32)
R["pos"] = R["pos"].copy()
R["pos"][-dup.size:] += 1e-4 * (rng.random((dup.size, 3)) - 0.5)
dev = case.make_cloud(DeviceLibrary()); ora = case.make_cloud(OracleLib())
dev.upload_particles(R); ora.upload_particles(R)
rd = dev.evolve(1)[0]; ro = ora.evolve(1)[0]
print(rd.nCloned, ro.nCloned, rd.nEliminated, ro.nEliminated, rd.nParticles, ro.nParticles)
a = sort_by_id(dev.download_particles()); b = sort_by_id(ora.download_particles())
bad = np.nonzero(a["tet"] != b["tet"])[0]
for i in bad:
    print(a["id"][i], a["cell"][i], b["cell"][i], a["tet"][i], b["tet"][i], a["pos"][i], b["pos"][i], a["pos"][i]-b["pos"][i], a["m"][i], b["m"][i])
for c in sorted(set(a["cell"][bad].tolist())):
    for name, p in (("dev", a), ("ora", b)):
        sel = np.nonzero(p["cell"] == c)[0]
        order = np.argsort(-(p["eta"][sel]*p["m"][sel]))
        print(name)
        for i in sel[order]:
            print("  id", p["id"][i], "eta*m %.17g" % (p["eta"][i]*p["m"][i]), "eta %.17g m %.17g" % (p["eta"][i], p["m"][i]), "Ux %.6f" % p["U"][i,0])
