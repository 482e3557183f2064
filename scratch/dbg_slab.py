import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from pdffoam_b200 import cases, meshgen
from pdffoam_b200.capi import DeviceLibrary
from oracle.oracle import OracleLib
from parity_util import sort_by_id
mesh = meshgen.slab_mesh(11, 5, lo=(0.0, 0.0), hi=(6.0, 2.0), thickness=2.0)
nc, nb = mesh.n_cells, mesh.n_boundary_faces
own = mesh.owner[mesh.n_internal_faces:]
U = np.tile([1.0, 0.2, 0.0], (nc, 1)); k = np.full(nc, 0.3)
R = np.zeros((nc, 6)); R[:, 0] = R[:, 3] = R[:, 5] = 0.2
Ub = U[own].copy(); Ub[:, 2] = 0
for name, comp in (("xmin", 0), ("xmax", 0), ("ymin", 1), ("ymax", 1)):
    Ub[mesh.boundary_slice(name), comp] = 0.0
fv = dict(U=U, Ub=Ub, UbFixed=np.zeros(nb, np.uint8), p=np.zeros(nc), pb=np.zeros(nb), k=k, kb=k[own].copy(),
          kbFixed=np.zeros(nb, np.uint8), epsilon=np.full(nc, 0.6), epsilonb=np.full(nb, 0.6), R=R, Rb=R[own].copy())
cc = mesh.cell_centres_simple()
z = cc[:, 0] / 6.0
cf = dict(rho=np.ones(nc), rhob=np.ones(nb), rhobFixed=np.zeros(nb, np.uint8), Phi=[z], Phib=[z[own].copy()], PhibFixed=[np.zeros(nb, np.uint8)])
prm = cases.default_params(1, particlesPerCell=30, mixed=[0], conserved=[0], CFL=0.3, averagingCoeff=10.0, DNum=0.05,
                           positionCorrection="limitedSimple", pcC=0.2, localTimeStepping="cell", kMin=1e-8,
                           relaxTimeU=1e-5, relaxTimek=1e-5, deltaT0=5e-3)
case = cases.Case("slab", mesh, prm, fv, cf)
dev = case.make_cloud(DeviceLibrary()); ora = case.make_cloud(OracleLib())
dev.seed_particles(); ora.seed_particles()
for s in range(5):
    rd = dev.evolve(1)[0]; ro = ora.evolve(1)[0]
    a = sort_by_id(dev.download_particles()); b = sort_by_id(ora.download_particles())
    print("step", s, "dt", rd.deltaT, ro.deltaT, "CoMax", rd.CoMax, ro.CoMax, "Ucorr", rd.UcorrMax, ro.UcorrMax, "eta", rd.etaMin, ro.etaMin, rd.etaMax, ro.etaMax)
    for col in ("pos","U","eta","Omega","m"):
        d = np.abs(a[col]-b[col]); 
        if d.ndim>1: d=d.max(1)
        bad = np.nonzero(d > 1e-9)[0]
        print("   ", col, "nbad", bad.size, "max", d.max(), "ids", a["id"][bad[:6]], "cells", a["cell"][bad[:6]])
    cd, co = dev.download_cell_fields(), ora.download_cell_fields()
    for kk in ("rho","pnd","U","k"):
        print("    cell", kk, np.abs(cd[kk]-co[kk]).max())
