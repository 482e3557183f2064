"""Case builders: synthetic frozen FV fields + numerics for the configurations
named in BASELINE.json / SURVEY.md §8(d) and Appendix E.

A `Case` bundles everything `mcParticleCloud`'s constructor borrows from the
OpenFOAM side (mcParticleCloud.C:301-735): the mesh, the FV fields U/p/k/epsilon/R
with their boundary values, rho and the scalar fields with their boundary
conditions, the mcSolution numerics and the model selection.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import meshgen
from .capi import MAX_ADDED, MAX_PHI, CloudHandle, Library, Params
from .meshgen import PolyMesh

# run-time selection names of the reference -> enum values of include/pdfb200.h
VELOCITY_MODELS = {"none": 0, "SLMFull": 1}                 # mcSLMFullVelocityModel.C:39-45
MIXING_MODELS = {"none": 0, "IEM": 1}                       # mcIEMMixingModel.C:37-43
OMEGA_MODELS = {"none": 0, "RAS": 1}                        # mcRASOmegaModel.C:63-69
REACTION_MODELS = {"none": 0, "cold": 1, "BurkeSchumann": 2, "SteadyFlamelet": 3}
POSITION_CORRECTIONS = {"none": 0, "limitedSimple": 1}
LOCAL_TIME_STEPPING = {"none": 0, "off": 0, "false": 0, "cell": 1, "particle": 2}   # mcLocalTimeStepping.C:56-97
INTERPOLATION_SCHEMES = {"cell": 0, "cellPointFace": 1}


def default_params(n_phi: int = 1, **kw) -> Params:
    """mcSolution defaults (mcSolution.C:44-57) + the model defaults cited in
    include/pdfb200.h. Keyword arguments override fields by name; model and
    scheme fields also accept the reference's run-time selection words."""
    p = Params()
    p.nPhi = n_phi
    p.nMixed = 0
    p.nConserved = 0
    p.CFL = 0.3
    p.averagingCoeff = 1000.0
    p.particlesPerCell = 40
    p.particleNumberControl = 0
    p.cloneAt, p.eliminateAt = 0.8, 1.2
    p.kMin = 100.0 * 1e-15
    p.DNum = 0.0
    p.relaxTimeU = p.relaxTimek = 1e15          # defaultRelaxationTime GREAT
    p.minRelaxTime = 10.0                       # FV deltaT (1) * minRelaxationFactor (10)
    for name in ("interpU", "interpk", "interpDiffU", "interpOmega", "interpUPosCorr", "interpZVar", "interpEta"):
        setattr(p, name, 1)                     # every tutorial selects cellPointFace
    p.velocityModel, p.C0, p.C1 = 1, 2.1, 1.0
    p.mixingModel, p.Cmix = 1, 2.0
    p.omegaModel, p.Omega0 = 1, 1.7976931348623157e308   # HUGE
    p.reactionModel, p.coldDensity = 1, 1.0
    p.zIdx, p.TIdx, p.chiIdx = 0, 1, 1
    p.Cchi = 2.0
    p.nAdded = 0
    p.positionCorrection, p.pcC, p.pcCFL = 0, 1e-2, 0.5
    p.localTimeStepping, p.ltsUpperBound = 0, 10.0
    p.k0 = 1e-15
    p.deltaT0 = 100.0 * 1e-15
    p.seed = 0x5EED
    p.rngMode = 0
    selectors = {"velocityModel": VELOCITY_MODELS, "mixingModel": MIXING_MODELS, "omegaModel": OMEGA_MODELS,
                 "reactionModel": REACTION_MODELS, "positionCorrection": POSITION_CORRECTIONS,
                 "localTimeStepping": LOCAL_TIME_STEPPING}
    for k, v in kw.items():
        if k in ("mixed", "conserved", "addedIdx"):
            arr = getattr(p, k)
            for i, x in enumerate(v):
                arr[i] = x
            setattr(p, {"mixed": "nMixed", "conserved": "nConserved", "addedIdx": "nAdded"}[k], len(v))
        elif k in selectors and isinstance(v, str):
            if v not in selectors[k]:
                raise KeyError(f"Unknown {k} type {v}; valid types are {sorted(selectors[k])}")
            setattr(p, k, selectors[k][v])
        elif k.startswith("interp") and isinstance(v, str):
            setattr(p, k, INTERPOLATION_SCHEMES[v])
        else:
            if not hasattr(p, k):
                raise KeyError(k)
            setattr(p, k, v)
    return p


@dataclass
class Case:
    name: str
    mesh: PolyMesh
    params: Params
    fv: Dict[str, np.ndarray]
    cf: Dict
    tables: Optional[Dict] = None     # {"chi","z","fields":[rho, added...]}
    info: Dict = field(default_factory=dict)

    def make_cloud(self, lib: Library, device: int = 0, capacity: int = 0) -> CloudHandle:
        """Construct the cloud and hand it every borrowed field (the work of
        mcParticleCloud's constructor up to checkMoments, .C:301-723)."""
        if capacity <= 0:
            capacity = int(self.mesh.n_cells * self.params.particlesPerCell * 1.5) + 4096
        h = CloudHandle(lib, self.mesh, self.params, device=device, capacity=capacity)
        if self.tables is not None:
            h.set_flamelet_tables(self.tables["chi"], self.tables["z"], self.tables["fields"])
        h.upload_fv_fields(self.fv)
        h.upload_cloud_fields(self.cf)
        h.init_moments()
        return h

    def random_particles(self, ppc: int, seed: int = 1234) -> Dict[str, np.ndarray]:
        """Host-side analogue of initReleaseParticles for axis-aligned hex
        meshes (used by parity tests that upload the same particles to the
        oracle and to the device)."""
        m = self.mesh
        rng = np.random.default_rng(seed)
        nc = m.n_cells
        n = nc * ppc
        cell = np.repeat(np.arange(nc, dtype=np.int32), ppc)
        lo, hi = self.info["cell_lo"], self.info["cell_hi"]
        u = rng.uniform(0.02, 0.98, size=(n, 3))
        pos = lo[cell] + u * (hi[cell] - lo[cell])
        for d in range(3):
            if m.geometric_d[d] < 0:
                pos[:, d] = 0.0
        k = self.fv["k"][cell]
        U = self.fv["U"].reshape(-1, 3)[cell] + rng.standard_normal((n, 3)) * np.sqrt(2.0 / 3.0 * k)[:, None]
        for d in range(3):
            if m.solution_d[d] < 0:
                U[:, d] = 0.0
        vol = self.info["cell_volumes"]
        rho = self.cf["rho"][cell]
        P = dict(pos=pos, U=U, m=rho * vol[cell] / ppc, Omega=(self.fv["epsilon"] / self.fv["k"])[cell].copy(),
                 rho=rho.copy(), eta=np.ones(n), Phi=[self.cf["Phi"][i][cell].copy() for i in range(self.params.nPhi)],
                 cell=cell, id=np.arange(n, dtype=np.uint32))
        return P


def _bnd_owner(mesh: PolyMesh) -> np.ndarray:
    return mesh.owner[mesh.n_internal_faces:]


def _flags(mesh: PolyMesh, fixed_patches: Sequence[str]) -> np.ndarray:
    f = np.zeros(mesh.n_boundary_faces, np.uint8)
    for name in fixed_patches:
        f[mesh.boundary_slice(name)] = 1
    return f


def _hex_info(mesh: PolyMesh) -> Dict:
    """cell bounding boxes and volumes of an axis-aligned structured hex mesh"""
    nx, ny, nz = mesh.shape
    pts = mesh.points
    ii = np.arange(mesh.n_cells)
    i, j, k = ii % nx, (ii // nx) % ny, ii // (nx * ny)
    p0 = i + (nx + 1) * (j + (ny + 1) * k)
    p1 = (i + 1) + (nx + 1) * ((j + 1) + (ny + 1) * (k + 1))
    lo, hi = pts[p0], pts[p1]
    return dict(cell_lo=lo, cell_hi=hi, cell_volumes=np.prod(hi - lo, axis=1))


def synthetic_box(nx=20, ny=10, nz=10, ppc=50, variant="stochastic", closed=False, n_phi=1,
                  reaction="cold", lts="cell", poscorr="limitedSimple", number_control=False,
                  tables: Optional[Dict] = None, seed=0x5EED, U0=10.0) -> Case:
    """The synthetic 3-D box of SURVEY §8(d): [0,2]x[0,1]x[0,1], analytic frozen
    FV fields, x faces inletOutlet (or slip if `closed`), y/z faces slip.

    variant="deterministic": C0=C1=0, uniform k and p (grad p* = 0), relaxation
    times 1e300, DNum=0, no position correction -> velocities frozen and the
    tracking purely geometric (bit-exact cell/tet parity case).
    """
    if closed:
        mesh = meshgen.box_mesh(nx, ny, nz, xmin=("xmin", meshgen.BC_SLIP), xmax=("xmax", meshgen.BC_SLIP))
        inlet, outlet = "xmin", "xmax"
    else:
        mesh = meshgen.box_mesh(nx, ny, nz)
        inlet, outlet = "inlet", "outlet"
    det = variant == "deterministic"
    cc = mesh.cell_centres_simple()
    fcb = mesh.face_centres_simple()[mesh.n_internal_faces:]
    own = _bnd_owner(mesh)

    def Ufun(x):
        return np.stack([U0 * (1 + 0.2 * np.sin(2 * np.pi * x[:, 1])), 0.05 * U0 * np.sin(np.pi * x[:, 0]),
                         0.05 * U0 * np.sin(2 * np.pi * x[:, 2])], axis=1)

    def kfun(x):
        if det:
            return np.full(x.shape[0], 0.015 * U0 ** 2)
        return 0.015 * U0 ** 2 * (1 + 0.5 * np.cos(2 * np.pi * x[:, 1]))

    def zfun(x):
        return 0.5 * (1.0 - np.tanh((x[:, 1] - 0.5) * 10.0))

    rho0 = 1.0
    U = Ufun(cc); k = kfun(cc)
    eps = 0.09 ** 0.75 * k ** 1.5 / 0.1
    # deterministic variant: p = 2/3 rho k makes p* = p - 2/3 rho k exactly zero
    p = (2. / 3. * rho0) * k if det else -rho0 * U0 ** 2 * (cc[:, 0] / 2)
    R = np.zeros((mesh.n_cells, 6)); R[:, 0] = R[:, 3] = R[:, 5] = 2.0 / 3.0 * k

    fixed_in = [] if closed else [inlet]
    UbFixed = _flags(mesh, fixed_in); kbFixed = _flags(mesh, fixed_in)
    Ub = U[own].copy(); kb = k[own].copy(); epsb = eps[own].copy(); pb = p[own].copy(); Rb = R[own].copy()
    if not closed:
        s = mesh.boundary_slice(inlet)
        Ub[s] = Ufun(fcb[s]); kb[s] = kfun(fcb[s])
        epsb[s] = 0.09 ** 0.75 * kb[s] ** 1.5 / 0.1
        Rb[s] = 0; Rb[s, 0] = Rb[s, 3] = Rb[s, 5] = 2.0 / 3.0 * kb[s]
        so = mesh.boundary_slice(outlet)
        Ub[so] = U[own][so]
    # slip walls: zero normal component of the boundary velocity (slip BC)
    for name, comp in (("ymin", 1), ("ymax", 1), ("zmin", 2), ("zmax", 2)):
        Ub[mesh.boundary_slice(name), comp] = 0.0
    if closed:
        Ub[mesh.boundary_slice("xmin"), 0] = 0.0
        Ub[mesh.boundary_slice("xmax"), 0] = 0.0

    fv = dict(U=U, Ub=Ub, UbFixed=UbFixed, p=p, pb=pb, k=k, kb=kb, kbFixed=kbFixed,
              epsilon=eps, epsilonb=epsb, R=R, Rb=Rb)

    rho = np.full(mesh.n_cells, rho0)
    Phi, Phib, PhibFixed = [], [], []
    zc = zfun(cc)
    for i in range(n_phi):
        if i == 0:
            c = zc.copy(); b = c[own].copy()
            if not closed:
                b[mesh.boundary_slice(inlet)] = (fcb[mesh.boundary_slice(inlet)][:, 1] < 0.5).astype(float)
        else:
            c = np.zeros(mesh.n_cells); b = np.zeros(mesh.n_boundary_faces)
        Phi.append(c); Phib.append(b); PhibFixed.append(_flags(mesh, fixed_in))
    cf = dict(rho=rho, rhob=rho[own].copy(), rhobFixed=_flags(mesh, fixed_in), Phi=Phi, Phib=Phib, PhibFixed=PhibFixed)

    kw = dict(particlesPerCell=ppc, mixed=[0], conserved=[0], CFL=0.3, averagingCoeff=1e3, Cmix=2.0,
              C0=2.1, C1=1.0, DNum=0.05, reactionModel=reaction, localTimeStepping=lts, ltsUpperBound=20.0,
              positionCorrection=poscorr, pcC=0.2, particleNumberControl=int(number_control),
              cloneAt=0.8, eliminateAt=1.2, kMin=1e-8, seed=seed, relaxTimeU=1e-5, relaxTimek=1e-5)
    if det:
        kw.update(C0=0.0, C1=0.0, DNum=0.0, relaxTimeU=1e300, relaxTimek=1e300, positionCorrection="none",
                  localTimeStepping="none", mixingModel="IEM")
    prm = default_params(n_phi, **kw)
    info = _hex_info(mesh)
    info.update(inlet=inlet, outlet=outlet, U0=U0)
    return Case(name=f"syntheticBox{nx}x{ny}x{nz}", mesh=mesh, params=prm, fv=fv, cf=cf, tables=tables, info=info)
