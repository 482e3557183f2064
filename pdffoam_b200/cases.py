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
MIXING_MODELS = {"none": 0, "IEM": 1, "modifiedCurl": 2}     # mcIEMMixingModel.C:37-43; modifiedCurl is new
OMEGA_MODELS = {"none": 0, "RAS": 1}                        # mcRASOmegaModel.C:63-69
REACTION_MODELS = {"none": 0, "cold": 1, "BurkeSchumann": 2, "SteadyFlamelet": 3, "naiveFlamelet": 4, "KulkarniAutoIgnition": 5}
POSITION_CORRECTIONS = {"none": 0, "limitedSimple": 1, "simple": 2, "external": 3}   # external: ellipticRelaxation / Muradoglu / integrated solved on the OpenFOAM side
INLET_RANDOM = {"inversion": 0, "sampling": 1}             # mcInletRandom.C:66-90
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
                 "reactionModel": REACTION_MODELS, "positionCorrection": POSITION_CORRECTIONS, "inletRandom": INLET_RANDOM,
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
        if self.tables is not None and "pv" in self.tables:
            h.set_autoignition_tables(self.tables["z"], self.tables["pv"], self.tables["cEq"], self.tables["fields"])
        elif self.tables is not None:
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
                 cell=cell, id=np.arange(n, dtype=np.uint64))
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


def with_scalars(case: Case, names: Sequence[str], reaction: str, tables: Optional[Dict] = None, **kw) -> Case:
    """Re-dress a case with the scalar set of a reacting tutorial:
    ("z","chi","T") + SteadyFlamelet (SandiaD, SydneyBluffBodyFlame:
    constant/thermophysicalProperties:11-20) or ("z","T") + BurkeSchumann
    (SandiaPropaneJet: constant/thermophysicalProperties:18-31)."""
    n_phi = len(names)
    nc, nb = case.mesh.n_cells, case.mesh.n_boundary_faces
    z, zb, zf = case.cf["Phi"][0], case.cf["Phib"][0], case.cf["PhibFixed"][0]
    Phi, Phib, PhibFixed = [z], [zb], [zf]
    for _ in names[1:]:
        Phi.append(np.zeros(nc)); Phib.append(np.zeros(nb)); PhibFixed.append(np.zeros(nb, np.uint8))
    case.cf.update(Phi=Phi, Phib=Phib, PhibFixed=PhibFixed)
    n_cov = n_phi * (n_phi + 1) // 2
    cov = [np.zeros(nc) for _ in range(n_cov)]
    cov[0] = 0.05 * z * (1 - z)                      # zzCov: some variance so that chi = Cchi*Omega*zVar varies
    case.cf.update(Cov=cov, Covb=[c[_bnd_owner(case.mesh)].copy() for c in cov], CovbFixed=[np.zeros(nb, np.uint8)] * n_cov)
    p = case.params
    base = {f[0]: getattr(p, f[0]) for f in p._fields_ if f[0] not in ("mixed", "conserved", "addedIdx", "nPhi", "nMixed", "nConserved", "nAdded")}
    base.update(reactionModel=reaction, zIdx=0)
    if reaction == "SteadyFlamelet":
        base.update(chiIdx=names.index("chi"), addedIdx=[names.index("T")], Cchi=2.0)
        case.tables = tables
    elif reaction == "BurkeSchumann":
        # tutorials/SandiaPropaneJet/constant/thermophysicalProperties:22-31 (stoichiometric T raised so that rho varies)
        base.update(TIdx=names.index("T"), bsRox=287.007, bsRfuel=188.554, bsRstoi=237.780, bsTox=294.0, bsTfuel=294.0,
                    bsTstoi=2000.0, bsZstoi=0.5)
        case.fv["p"] = case.fv["p"] + 1.0e5          # absolute pressure for rho = p/(R T)
        case.fv["pb"] = case.fv["pb"] + 1.0e5
    elif reaction == "naiveFlamelet":
        base.update(TIdx=names.index("T"))
    elif reaction == "KulkarniAutoIgnition":
        # tutorials/AutoIgnitionBurner/constant/thermophysicalProperties:18-25: scalarFields (z c pv T d), scalars (T)
        base.update(cIdx=names.index("c"), pvIdx=names.index("pv"), addedIdx=[names.index("T")])
        case.tables = tables
    base.update(kw)
    base.setdefault("mixed", [0]); base.setdefault("conserved", [0])
    case.params = default_params(n_phi, **base)
    return case


def load_autoignition_tables(path: str) -> Dict:
    """tests/golden/autoignition.npz -> {"z","pv","cEq","fields":[cdot, rho, T]} (rows = z, columns = pv,
    mcKulkarniAutoIgnition.C:111-335)"""
    t = np.load(path)
    return dict(z=t["z"], pv=t["pv"], cEq=t["cEq"], fields=[t["cdot"], t["rho"], t["T"]])


def load_flamelet_tables(path: str) -> Dict:
    """tests/golden/flamelet_*.npz -> {"chi","z","fields":[rho, T]} (rows = chi, columns = z,
    mcSteadyFlamelet.C:261-286)"""
    t = np.load(path)
    return dict(chi=t["chi"], z=t["z"], fields=[t["rho"], t["T"]])


def _rotation_tensor(n1, n2):
    """OpenFOAM rotationTensor(n1, n2) (transform.H)"""
    n1, n2 = np.asarray(n1, float), np.asarray(n2, float)
    s = n1 @ n2
    n3 = np.cross(n1, n2)
    return s * np.eye(3) + (1 - s) * np.outer(n3, n3) / (n3 @ n3 + 1e-300) + (np.outer(n2, n1) - np.outer(n1, n2))


def wedge_jet(nx=24, nr_jet=4, nr_co=10, length=0.3, r_min=2e-4, r_jet=3.6e-3, r_max=0.05, half_angle_deg=5.0, ppc=30,
              U_jet=50.0, U_co=1.0, number_control=True, lts="cell", poscorr="limitedSimple", seed=0x5EED) -> Case:
    """An axisymmetric jet in a coflow on a one-cell-thick wedge — the shape of all
    four tutorial configs: patches jet / coflow / outflow ->
    inletOutlet, axis / outerWall -> slip, front / back -> wedge (empty handler)."""
    xs = meshgen.graded(nx, 0.0, length, 4.0)
    rs = np.concatenate([meshgen.graded(nr_jet, r_min, r_jet, 1.0), meshgen.graded(nr_co, r_jet, r_max, 6.0)[1:]])
    mesh = meshgen.wedge_mesh(xs, rs, half_angle_deg, inlet_split=[("jet", r_jet), ("coflow", 1e30)])
    th = np.deg2rad(half_angle_deg)
    cc = mesh.cell_centres_simple()
    fcb = mesh.face_centres_simple()[mesh.n_internal_faces:]
    own = _bnd_owner(mesh)
    nc, nb = mesh.n_cells, mesh.n_boundary_faces

    def radius(x):
        return np.hypot(x[:, 1], x[:, 2])

    def Ufun(x):
        r = radius(x)
        spread = r_jet * (1 + 6.0 * x[:, 0] / length)
        ux = U_co + (U_jet - U_co) * (r_jet / spread) * np.exp(-(r / spread) ** 2)
        ur = 0.02 * ux * r / spread
        return np.stack([ux, ur, np.zeros_like(ux)], axis=1)

    def kfun(x):
        return 0.01 * Ufun(x)[:, 0] ** 2 + 0.05

    def zfun(x):
        r = radius(x)
        spread = r_jet * (1 + 6.0 * x[:, 0] / length)
        return np.clip((r_jet / spread) * np.exp(-(r / spread) ** 2), 0.0, 1.0)

    U = Ufun(cc); k = kfun(cc); eps = 0.09 ** 0.75 * k ** 1.5 / (0.5 * r_jet * 10)
    p = np.zeros(nc)
    R = np.zeros((nc, 6)); R[:, 0] = R[:, 3] = R[:, 5] = 2.0 / 3.0 * k
    Ub = U[own].copy(); kb = k[own].copy(); epsb = eps[own].copy(); pb = p[own].copy(); Rb = R[own].copy()
    fixed = ["jet", "coflow"]
    for name in fixed:
        s = mesh.boundary_slice(name)
        Ub[s] = Ufun(fcb[s]); Ub[s, 1] = 0.0
        kb[s] = kfun(fcb[s]); epsb[s] = 0.09 ** 0.75 * kb[s] ** 1.5 / (0.5 * r_jet * 10)
        Rb[s] = 0; Rb[s, 0] = Rb[s, 3] = Rb[s, 5] = 2.0 / 3.0 * kb[s]
    for name in ("axis", "outerWall"):
        Ub[mesh.boundary_slice(name), 1] = 0.0
    # wedge patches: wedgeFvPatchField::evaluate rotates the cell vector onto the patch
    for name, sign in (("back", -1.0), ("front", 1.0)):
        s = mesh.boundary_slice(name)
        n = np.array([0.0, -np.sin(th), np.cos(th)]) if sign > 0 else np.array([0.0, -np.sin(th), -np.cos(th)])   # outward patch normal
        cn = np.array([0.0, 0.0, 1.0]) if sign > 0 else np.array([0.0, 0.0, -1.0])
        T = _rotation_tensor(cn, n)
        Ub[s] = U[own][s] @ T.T
    fv = dict(U=U, Ub=Ub, UbFixed=_flags(mesh, fixed), p=p, pb=pb, k=k, kb=kb, kbFixed=_flags(mesh, fixed),
              epsilon=eps, epsilonb=epsb, R=R, Rb=Rb)
    rho = np.ones(nc)
    z = zfun(cc); zb = z[own].copy()
    zb[mesh.boundary_slice("jet")] = 1.0; zb[mesh.boundary_slice("coflow")] = 0.0
    cf = dict(rho=rho, rhob=rho[own].copy(), rhobFixed=_flags(mesh, fixed), Phi=[z], Phib=[zb], PhibFixed=[_flags(mesh, fixed)])
    prm = default_params(1, particlesPerCell=ppc, mixed=[0], conserved=[0], CFL=0.3, averagingCoeff=50.0, DNum=0.05,
                         localTimeStepping=lts, ltsUpperBound=20.0, positionCorrection=poscorr, pcC=5e-2,
                         particleNumberControl=int(number_control), kMin=1e-8, Omega0=1e5, seed=seed,
                         relaxTimeU=1e-5, relaxTimek=1e-5)
    # cell volumes of the wedge hexes: (r2^2 - r1^2)/2 * 2*theta ... computed by the library; keep the case light
    return Case(name=f"wedgeJet{nx}x{nr_jet + nr_co}", mesh=mesh, params=prm, fv=fv, cf=cf, info=dict(theta=th))


# ---------------------------------------------------------------------------
# The four tutorial configurations of BASELINE.json configs[0..3]:
# mesh shape, patch set, scalar set, model selection and numerics follow the cited tutorial
# files; the frozen FV fields are synthetic jet profiles (OpenFOAM's RAS solution is not
# available here), which is what "synthetic data of that shape" means for these cases.
# ---------------------------------------------------------------------------
def _streams_case(name: str, xs, r_blocks, half_angle_deg, streams, tables_path: Optional[str], scalars: Sequence[str],
                  reaction: str, numerics: Dict, golden_dir: Optional[str] = None, solenoidal: bool = False) -> Case:
    """streams: [(patch name, r_upper, U, z, rho, handler)] from the axis outwards;
    r_blocks: [(n cells, r_lo, r_hi, grading)] (blockMesh blocks in r).
    solenoidal: the frozen velocity field is the similarity solution of the round turbulent jet (Schlichting) in a
    uniform coflow, derived from a stream function, so that the volume flux is divergence-free and the cloud has a
    stationary state (the default field spreads outwards everywhere and slowly piles mass up at the outer wall)."""
    rs = np.concatenate([meshgen.graded(n, lo, hi, g)[(1 if i else 0):] for i, (n, lo, hi, g) in enumerate(r_blocks)])
    mesh = meshgen.wedge_mesh(xs, rs, half_angle_deg, inlet_split=[(s[0], s[1], s[5]) for s in streams])
    th = np.deg2rad(half_angle_deg)
    cc = mesh.cell_centres_simple()
    fcb = mesh.face_centres_simple()[mesh.n_internal_faces:]
    own = _bnd_owner(mesh)
    nc, nb = mesh.n_cells, mesh.n_boundary_faces
    length = float(xs[-1])
    inflow = [s for s in streams if s[5] == meshgen.BC_INLET_OUTLET]
    r_jet, U_jet, U_co = inflow[0][1], inflow[0][2], inflow[-1][2]

    def radius(x):
        return np.hypot(x[:, 1], x[:, 2])

    def stream_of(r):
        idx = np.searchsorted(np.array([s[1] for s in streams]), r, side="left")
        return np.minimum(idx, len(streams) - 1)

    sigma = 15.0
    x0 = r_jet * sigma / 1.29                 # virtual origin: half-width r_jet at the inlet
    Kj = (U_jet - U_co) * x0 / (2 * sigma ** 2)

    def Ufun(x):
        r = radius(x)
        if solenoidal:
            # psi = U_co r^2/2 + K (x+x0) xi^2/(1+xi^2/4), xi = sigma r/(x+x0): ux = psi_r / r, ur = -psi_x / r
            xx = x[:, 0] + x0
            xi = sigma * r / xx
            den = (1 + 0.25 * xi ** 2) ** 2
            ux = U_co + 2 * Kj * sigma ** 2 / xx / den
            ur = Kj * sigma * xi * (1 - 0.25 * xi ** 2) / (xx * den)
            return np.stack([ux, ur, np.zeros_like(ux)], axis=1)
        spread = r_jet * (1 + 8.0 * x[:, 0] / length)
        ux = U_co + (U_jet - U_co) * (r_jet / spread) * np.exp(-(r / spread) ** 2)
        for s in inflow[1:-1]:                                   # pilot-like annular streams decay into the coflow
            w = s[1] * (1 + 3.0 * x[:, 0] / length)
            ux = ux + (s[2] - U_co) * np.exp(-(r / w) ** 4) * np.exp(-6.0 * x[:, 0] / length) * (r > r_jet)
        ur = 0.02 * ux * r / spread
        return np.stack([ux, ur, np.zeros_like(ux)], axis=1)

    def kfun(x):
        return 0.01 * Ufun(x)[:, 0] ** 2 + 0.05

    def zfun(x):
        r = radius(x)
        spread = r_jet * (1 + 8.0 * x[:, 0] / length)
        return np.clip((r_jet / spread) * np.exp(-(r / spread) ** 2), 0.0, 1.0)

    lt = 0.5 * r_jet * 10
    U = Ufun(cc); k = kfun(cc); eps = 0.09 ** 0.75 * k ** 1.5 / lt
    p = np.zeros(nc)
    R = np.zeros((nc, 6)); R[:, 0] = R[:, 3] = R[:, 5] = 2.0 / 3.0 * k
    Ub = U[own].copy(); kb = k[own].copy(); epsb = eps[own].copy(); pb = p[own].copy(); Rb = R[own].copy()
    fixed = [s[0] for s in inflow]
    z = zfun(cc); zb = z[own].copy()
    rho_in = np.ones(nb)
    for s in streams:
        sl = mesh.boundary_slice(s[0])
        if s[5] == meshgen.BC_INLET_OUTLET:
            Ub[sl] = 0.0; Ub[sl, 0] = s[2]
            kb[sl] = 0.01 * s[2] ** 2 + 0.05; epsb[sl] = 0.09 ** 0.75 * kb[sl] ** 1.5 / lt
            Rb[sl] = 0; Rb[sl, 0] = Rb[sl, 3] = Rb[sl, 5] = 2.0 / 3.0 * kb[sl]
            zb[sl] = s[3]; rho_in[sl] = s[4]
        else:                                                    # bluff-body face: slip wall
            Ub[sl, 0] = 0.0
    for pname in ("axis", "outerWall"):
        Ub[mesh.boundary_slice(pname), 1] = 0.0
    for pname, sign in (("back", -1.0), ("front", 1.0)):
        sl = mesh.boundary_slice(pname)
        n = np.array([0.0, -np.sin(th), np.cos(th)]) if sign > 0 else np.array([0.0, -np.sin(th), -np.cos(th)])
        cn = np.array([0.0, 0.0, 1.0]) if sign > 0 else np.array([0.0, 0.0, -1.0])
        Ub[sl] = U[own][sl] @ _rotation_tensor(cn, n).T
    fv = dict(U=U, Ub=Ub, UbFixed=_flags(mesh, fixed), p=p, pb=pb, k=k, kb=kb, kbFixed=_flags(mesh, fixed),
              epsilon=eps, epsilonb=epsb, R=R, Rb=Rb)
    tables = None
    rho = np.ones(nc)
    if reaction == "SteadyFlamelet":
        tables = load_flamelet_tables(tables_path)
        rho = np.interp(z, tables["z"], tables["fields"][0][0])
    elif reaction == "cold":
        rho = np.full(nc, inflow[-1][4])
    rhob = rho[own].copy()
    for s in inflow:
        rhob[mesh.boundary_slice(s[0])] = s[4]
    n_phi = len(scalars)
    Phi, Phib, PhibFixed = [z], [zb], [_flags(mesh, fixed)]
    for _ in scalars[1:]:
        Phi.append(np.zeros(nc)); Phib.append(np.zeros(nb)); PhibFixed.append(np.zeros(nb, np.uint8))
    n_cov = n_phi * (n_phi + 1) // 2
    cov = [np.zeros(nc) for _ in range(n_cov)]
    cov[0] = 0.05 * z * (1 - z)
    cf = dict(rho=rho, rhob=rhob, rhobFixed=_flags(mesh, fixed), Phi=Phi, Phib=Phib, PhibFixed=PhibFixed,
              Cov=cov, Covb=[c[own].copy() for c in cov], CovbFixed=[np.zeros(nb, np.uint8)] * n_cov)
    kw = dict(mixed=[0], conserved=[0], reactionModel=reaction, zIdx=0, kMin=1e-8, DNum=0.05, Omega0=1e5,
              particleNumberControl=1, cloneAt=0.8, eliminateAt=1.2, localTimeStepping="cell", positionCorrection="limitedSimple",
              relaxTimeU=1e-5, relaxTimek=1e-5)
    if reaction == "SteadyFlamelet":
        kw.update(chiIdx=list(scalars).index("chi"), addedIdx=[list(scalars).index("T")], Cchi=2.0)
    elif reaction == "BurkeSchumann":
        # tutorials/SandiaPropaneJet/constant/thermophysicalProperties:22-31
        kw.update(TIdx=list(scalars).index("T"), bsRox=287.007, bsRfuel=188.554, bsRstoi=237.780, bsTox=294.0, bsTfuel=294.0,
                  bsTstoi=294.0, bsZstoi=0.5)
        fv["p"] = fv["p"] + 1.0e5; fv["pb"] = fv["pb"] + 1.0e5
        cf["rho"] = 1.0e5 / (287.007 * 294.0) * np.ones(nc); cf["rhob"] = cf["rho"][own].copy()
    kw.update(numerics)
    prm = default_params(n_phi, **kw)
    return Case(name=name, mesh=mesh, params=prm, fv=fv, cf=cf, tables=tables, info=dict(theta=th, r_jet=r_jet, streams=streams))


TUTORIALS = ("SydneyBluffBodyFlow", "SandiaD", "SydneyBluffBodyFlame", "SandiaPropaneJet4x")


def tutorial_case(name: str, ppc: Optional[int] = None, refine: int = 1, golden_dir: Optional[str] = None, solenoidal: bool = False) -> Case:
    """The four tutorial-shaped configurations of BASELINE.json (configs[0..3]); `refine` multiplies the cell
    counts of the blockMesh blocks in x and r, `ppc` overrides the particles per cell, `solenoidal` selects the
    divergence-free jet field (see _streams_case; used by the stationary-state validation)."""
    import os
    IO, SLIP = meshgen.BC_INLET_OUTLET, meshgen.BC_SLIP
    if golden_dir is None:
        golden_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    f = refine
    if name == "SandiaD":
        # tutorials/SandiaD/constant/polyMesh/blockMeshDict:22-48: blocks (1 54 8)(1 54 6)(1 54 20), grading x 15, r 2/2/8;
        # D = 7.2 mm, pilot to 18.2 mm; inlet z and rho from 0.org/z, 0.org/rho; system/mcThermoCloudSolution:11-53
        return _streams_case(
            "SandiaD", meshgen.graded(54 * f, 0.0, 0.8, 15.0),
            [(8 * f, 1.5e-4, 3.6e-3, 2.0), (6 * f, 3.6e-3, 9.1e-3, 2.0), (20 * f, 9.1e-3, 0.15, 8.0)], 5.0,
            [("jet", 3.6e-3, 49.6, 1.0, 1.08147, IO), ("pilot", 9.1e-3, 11.4, 0.268692, 0.178466, IO), ("coflow", 1e30, 0.9, 0.0, 1.21731, IO)],
            os.path.join(golden_dir, "flamelet_sandiaD.npz"), ("z", "chi", "T"), "SteadyFlamelet",
            dict(CFL=0.3, averagingCoeff=2e4, particlesPerCell=ppc or 40, ltsUpperBound=50.0, pcC=5e-2, Cmix=2.0, mixingModel="IEM"),
            solenoidal=solenoidal)
    if name == "SydneyBluffBodyFlame":
        # tutorials/SydneyBluffBodyFlame/constant/polyMesh/blockMeshDict:22-48: blocks (1 80 6)(1 80 40)(1 80 34);
        # jet r = 1.8 mm, bluff body r = 25 mm (slip), coflow to 150 mm; BASELINE configs[2]: modified-Curl mixing
        return _streams_case(
            "SydneyBluffBodyFlame", meshgen.graded(80 * f, 0.0, 0.3, 4.0),
            [(6 * f, 1.5e-4, 1.8e-3, 1.0), (40 * f, 1.8e-3, 25e-3, 1.0), (34 * f, 25e-3, 0.15, 6.0)], 5.0,
            [("jet", 1.8e-3, 108.0, 1.0, 0.84, IO), ("bluffBody", 25e-3, 0.0, 0.0, 1.0, SLIP), ("coflow", 1e30, 35.0, 0.0, 1.2, IO)],
            os.path.join(golden_dir, "flamelet_bluffbody.npz"), ("z", "chi", "T"), "SteadyFlamelet",
            dict(CFL=0.1, averagingCoeff=5e3, particlesPerCell=ppc or 40, ltsUpperBound=20.0, pcC=0.2, Cmix=2.0,
                 mixingModel="modifiedCurl", relaxTimeU=5e-5, relaxTimek=5e-5))
    if name == "SydneyBluffBodyFlow":
        # tutorials/SydneyBluffBodyFlow/flow/constant/polyMesh/blockMeshDict:46-48: blocks (1 64 6)(1 64 24)(1 64 34);
        # scalars case: (z), cold, IEM, ~100 particles per cell (BASELINE configs[0])
        return _streams_case(
            "SydneyBluffBodyFlow", meshgen.graded(64 * f, 0.0, 0.3, 4.0),
            [(6 * f, 1.5e-4, 1.8e-3, 1.0), (24 * f, 1.8e-3, 25e-3, 1.0), (34 * f, 25e-3, 0.15, 6.0)], 5.0,
            [("jet", 1.8e-3, 61.0, 1.0, 1.2, IO), ("bluffBody", 25e-3, 0.0, 0.0, 1.2, SLIP), ("coflow", 1e30, 20.0, 0.0, 1.2, IO)],
            None, ("z",), "cold",
            dict(CFL=0.1, averagingCoeff=1e3, particlesPerCell=ppc or 100, ltsUpperBound=20.0, pcC=0.2, Cmix=2.0, mixingModel="IEM",
                 coldDensity=1.2))
    if name == "SandiaPropaneJet4x":
        # tutorials/SandiaPropaneJet/constant/polyMesh/blockMeshDict:9-36: blocks (54 7 1)(54 5 1)(54 24 1), grading x 6,
        # half-angle 2.5 deg; BASELINE configs[3]: mesh refined 4x, 400 particles per cell
        g = 4 * f
        return _streams_case(
            "SandiaPropaneJet4x", meshgen.graded(54 * g, 0.0, 0.3, 6.0),
            [(7 * g, 1.0e-4, 2.6e-3, 1.0), (5 * g, 2.6e-3, 4.5e-3, 1.0), (24 * g, 4.5e-3, 0.04, 6.0)], 2.5,
            [("jet", 2.6e-3, 69.0, 1.0, 1.8, IO), ("bluffBody", 4.5e-3, 0.0, 0.0, 1.2, SLIP), ("coflow", 1e30, 9.2, 0.0, 1.2, IO)],
            None, ("z", "T"), "BurkeSchumann",
            dict(CFL=0.1, averagingCoeff=5e3, particlesPerCell=ppc or 400, ltsUpperBound=20.0, pcC=0.2, Cmix=2.0, mixingModel="IEM"))
    raise KeyError(f"unknown tutorial case {name}; valid: {TUTORIALS}")


def sandiaD_validation_case(seed: int = 0x5EED, rng_mode: int = 0) -> Case:
    """The case of the stationary-state validation (tests/test_gpu_sandiaD_profiles.py, tests/golden/
    make_sandiaD_ensemble.py): SandiaD-shaped (tutorial_case("SandiaD")) with the divergence-free jet field, a short
    averaging memory so that 4800 steps reach and sample the stationary state, and no position correction."""
    case = tutorial_case("SandiaD", solenoidal=True)
    case.params.averagingCoeff = 200.0
    case.params.positionCorrection = POSITION_CORRECTIONS["none"]
    case.params.seed = seed
    case.params.rngMode = rng_mode
    return case
