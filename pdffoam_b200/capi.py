"""ctypes binding of include/pdfb200.h.

`DeviceLibrary()` loads the CUDA library (pdffoam_b200/csrc/libpdfb200.so) and
fails loudly when it is missing: there is no CPU fallback in the product.
The same struct definitions are reused by oracle/oracle.py for the CPU oracle,
which exports the identical entry points with the prefix ``pdforacle_``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from .meshgen import PolyMesh

MAX_PHI, MAX_COV, MAX_ADDED = 6, 21, 4

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_uint32_p = C.POINTER(C.c_uint32)
c_uint64_p = C.POINTER(C.c_uint64)


class MeshStruct(C.Structure):
    _fields_ = [
        ("nPoints", C.c_int32), ("nFaces", C.c_int32), ("nInternalFaces", C.c_int32),
        ("nCells", C.c_int32), ("nPatches", C.c_int32),
        ("points", c_double_p), ("faceStart", c_int32_p), ("facePoints", c_int32_p),
        ("owner", c_int32_p), ("neighbour", c_int32_p),
        ("patchStart", c_int32_p), ("patchSize", c_int32_p), ("patchGeom", c_int32_p),
        ("patchHandler", c_int32_p), ("patchReflecting", c_int32_p),
        ("geometricD", C.c_int32 * 3), ("solutionD", C.c_int32 * 3),
    ]


class Params(C.Structure):
    _fields_ = [
        ("nPhi", C.c_int32),
        ("nMixed", C.c_int32), ("mixed", C.c_int32 * MAX_PHI),
        ("nConserved", C.c_int32), ("conserved", C.c_int32 * MAX_PHI),
        ("CFL", C.c_double), ("averagingCoeff", C.c_double),
        ("particlesPerCell", C.c_int32), ("particleNumberControl", C.c_int32),
        ("cloneAt", C.c_double), ("eliminateAt", C.c_double),
        ("kMin", C.c_double), ("DNum", C.c_double),
        ("relaxTimeU", C.c_double), ("relaxTimek", C.c_double), ("minRelaxTime", C.c_double),
        ("interpU", C.c_int32), ("interpk", C.c_int32), ("interpDiffU", C.c_int32), ("interpOmega", C.c_int32),
        ("interpUPosCorr", C.c_int32), ("interpZVar", C.c_int32), ("interpEta", C.c_int32),
        ("velocityModel", C.c_int32), ("C0", C.c_double), ("C1", C.c_double),
        ("mixingModel", C.c_int32), ("Cmix", C.c_double),
        ("omegaModel", C.c_int32), ("Omega0", C.c_double),
        ("reactionModel", C.c_int32),
        ("coldDensity", C.c_double),
        ("bsRfuel", C.c_double), ("bsRox", C.c_double), ("bsRstoi", C.c_double),
        ("bsTfuel", C.c_double), ("bsTox", C.c_double), ("bsTstoi", C.c_double), ("bsZstoi", C.c_double),
        ("zIdx", C.c_int32), ("TIdx", C.c_int32), ("chiIdx", C.c_int32), ("cIdx", C.c_int32), ("pvIdx", C.c_int32),
        ("Cchi", C.c_double),
        ("nAdded", C.c_int32), ("addedIdx", C.c_int32 * MAX_ADDED),
        ("positionCorrection", C.c_int32), ("pcC", C.c_double), ("pcCFL", C.c_double),
        ("localTimeStepping", C.c_int32), ("ltsUpperBound", C.c_double),
        ("k0", C.c_double), ("deltaT0", C.c_double),
        ("seed", C.c_uint64), ("rngMode", C.c_int32), ("inletRandom", C.c_int32),
    ]


class FvFields(C.Structure):
    _fields_ = [
        ("U", c_double_p), ("Ub", c_double_p), ("UbFixed", c_uint8_p),
        ("p", c_double_p), ("pb", c_double_p),
        ("k", c_double_p), ("kb", c_double_p), ("kbFixed", c_uint8_p),
        ("epsilon", c_double_p), ("epsilonb", c_double_p),
        ("R", c_double_p), ("Rb", c_double_p),
    ]


class PoscorrFields(C.Structure):
    _fields_ = [("G0", c_double_p), ("G0b", c_double_p), ("G1", c_double_p), ("G1b", c_double_p), ("G2", c_double_p), ("G2b", c_double_p),
                ("l", c_double_p), ("lb", c_double_p), ("zeta", c_double_p), ("zetab", c_double_p), ("scale", C.c_double)]


class CloudFields(C.Structure):
    _fields_ = [
        ("rho", c_double_p), ("rhob", c_double_p), ("rhobFixed", c_uint8_p),
        ("Phi", c_double_p * MAX_PHI), ("Phib", c_double_p * MAX_PHI), ("PhibFixed", c_uint8_p * MAX_PHI),
        ("Cov", c_double_p * MAX_COV), ("Covb", c_double_p * MAX_COV), ("CovbFixed", c_uint8_p * MAX_COV),
    ]


class ParticlesStruct(C.Structure):
    _fields_ = [
        ("pos", c_double_p), ("U", c_double_p),
        ("m", c_double_p), ("Omega", c_double_p), ("rho", c_double_p), ("eta", c_double_p),
        ("Phi", c_double_p * MAX_PHI),
        ("cell", c_int32_p), ("tet", c_int32_p), ("id", c_uint64_p),
    ]


class StepReport(C.Structure):
    _fields_ = [
        ("deltaT", C.c_double), ("deltaTNext", C.c_double), ("rhoRes", C.c_double),
        ("rhoErrMin", C.c_double), ("rhoErrMax", C.c_double),
        ("UMax", C.c_double), ("UcorrMax", C.c_double), ("etaMin", C.c_double), ("etaMax", C.c_double), ("CoMax", C.c_double),
        ("totalMass", C.c_double), ("totalParticleMass", C.c_double),
        ("deltaMassInst", C.c_double * (1 + MAX_PHI)),
        ("massInInst", C.c_double * (1 + MAX_PHI)),
        ("massOutInst", C.c_double * (1 + MAX_PHI)),
        ("nParticles", C.c_int64),
        ("nInjected", C.c_int64), ("nDeleted", C.c_int64), ("nCloned", C.c_int64),
        ("nEliminated", C.c_int64), ("nLost", C.c_int64),
        ("momentsAllreduceMs", C.c_double),
    ]

    def as_dict(self) -> Dict:
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class CellFieldsStruct(C.Structure):
    _fields_ = [
        ("rho", c_double_p), ("rhoInst", c_double_p), ("pnd", c_double_p), ("pndInst", c_double_p),
        ("U", c_double_p), ("Tau", c_double_p), ("k", c_double_p),
        ("Phi", c_double_p * MAX_PHI), ("Cov", c_double_p * MAX_COV), ("UPhi", c_double_p * MAX_PHI),
        ("PaNIC", c_double_p),
        ("mMom", c_double_p), ("VMom", c_double_p), ("UMom", c_double_p), ("UUMom", c_double_p),
        ("PhiMom", c_double_p * MAX_PHI), ("PhiPhiMom", c_double_p * MAX_COV), ("UPhiMom", c_double_p * MAX_PHI),
    ]


class MeshInfo(C.Structure):
    _fields_ = [
        ("nTets", C.c_int64), ("maxTetsPerCell", C.c_int32), ("isAxiSymmetric", C.c_int32),
        ("axis", C.c_double * 3), ("centrePlaneNormal", C.c_double * 3), ("openingAngle", C.c_double),
    ]


def _dp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def _ip(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


def _u8p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_uint8_p)


def _u64p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_uint64_p)


def jit_compile(lib: "Library", params: Params, axisym: bool, has_open: bool, tet_bits: int = 0,
                geometric_d=(1, 1, 1), solution_d=(1, 1, 1), cell_faces: int = 6, compiler: int = 0):
    """Compile the specialised kernel for these parameters the way the library does at the first step
    (embedded sources; compiler 0 = first available, 1 = NVRTC, 2 = nvcc); host only. -> (cubin size or -1, log)"""
    buf = C.create_string_buffer(1 << 16)
    g = np.asarray(geometric_d, np.int32); sd = np.asarray(solution_d, np.int32)
    n = lib.fn("jit_compile")(C.byref(params), int(axisym), int(has_open), int(tet_bits), int(cell_faces), _ip(g), _ip(sd), int(compiler), buf, len(buf))
    return int(n), buf.value.decode(errors="replace")


class PdfError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status (the reference would
    FatalErrorIn(...) << exit(FatalError))."""


# every symbol include/pdfb200.h declares (tests check the .so exports them all)
API_SYMBOLS = [
    "last_error", "create", "destroy", "set_params", "set_flamelet_tables", "set_autoignition_tables", "upload_fv_fields", "stage_fv_fields", "commit_fv_fields",
    "upload_cloud_fields", "upload_poscorr_fields", "init_moments", "seed_particles", "upload_particles", "num_particles",
    "download_particles", "evolve", "download_cell_fields", "download_cell_fields_async", "sync_downloads", "upload_moments", "step_counter", "set_step_counter", "id_counter", "set_id_counter", "delta_t", "set_delta_t",
    "comm_unique_id", "comm_init", "mesh_info_get", "download_tets", "download_geometry",
    "kernel_launches", "last_evolve_ms", "profile", "profile_ctas", "set_graph_mode", "jit_source", "jit_compile", "rng_probe",
]


class Library:
    """A shared library exporting the pdfb200.h entry points under `prefix`."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        self._declare()

    def fn(self, name: str):
        return getattr(self.lib, self.prefix + name)

    def has(self, name: str) -> bool:
        return hasattr(self.lib, self.prefix + name)

    def _declare(self):
        vp = C.c_void_p
        f = self.fn
        f("last_error").restype = C.c_char_p
        f("create").argtypes = [C.POINTER(MeshStruct), C.POINTER(Params), C.c_int, C.c_int64, C.POINTER(vp)]
        f("destroy").argtypes = [vp]; f("destroy").restype = None
        f("set_params").argtypes = [vp, C.POINTER(Params)]
        f("set_flamelet_tables").argtypes = [vp, C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_int32, C.POINTER(c_double_p)]
        f("upload_fv_fields").argtypes = [vp, C.POINTER(FvFields)]
        f("upload_cloud_fields").argtypes = [vp, C.POINTER(CloudFields)]
        f("init_moments").argtypes = [vp]
        f("seed_particles").argtypes = [vp]
        f("upload_particles").argtypes = [vp, C.c_int64, C.POINTER(ParticlesStruct)]
        f("num_particles").argtypes = [vp]; f("num_particles").restype = C.c_int64
        f("download_particles").argtypes = [vp, C.c_int64, C.POINTER(ParticlesStruct), C.POINTER(C.c_int64)]
        f("evolve").argtypes = [vp, C.c_int32, C.POINTER(StepReport)]
        f("download_cell_fields").argtypes = [vp, C.POINTER(CellFieldsStruct)]
        f("download_cell_fields_async").argtypes = [vp, C.POINTER(CellFieldsStruct)]
        f("sync_downloads").argtypes = [vp]
        f("set_autoignition_tables").argtypes = [vp, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p, C.c_int32, C.POINTER(c_double_p)]
        f("upload_moments").argtypes = [vp, C.POINTER(CellFieldsStruct)]
        f("stage_fv_fields").argtypes = [vp, C.POINTER(FvFields)]
        f("commit_fv_fields").argtypes = [vp]
        f("upload_poscorr_fields").argtypes = [vp, C.POINTER(PoscorrFields)]
        f("step_counter").argtypes = [vp]; f("step_counter").restype = C.c_int64
        f("set_step_counter").argtypes = [vp, C.c_int64]
        f("set_graph_mode").argtypes = [vp, C.c_int]
        f("id_counter").argtypes = [vp]; f("id_counter").restype = C.c_int64
        f("set_id_counter").argtypes = [vp, C.c_int64]
        f("delta_t").argtypes = [vp]; f("delta_t").restype = C.c_double
        f("set_delta_t").argtypes = [vp, C.c_double]
        f("mesh_info_get").argtypes = [vp, C.POINTER(MeshInfo)]
        f("download_tets").argtypes = [vp, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]
        f("download_geometry").argtypes = [vp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        if self.has("comm_unique_id"):
            f("comm_unique_id").argtypes = [C.c_char_p]
            f("comm_init").argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p]
        if self.has("kernel_launches"):
            f("kernel_launches").argtypes = [vp]; f("kernel_launches").restype = C.c_int64
            f("last_evolve_ms").argtypes = [vp]; f("last_evolve_ms").restype = C.c_double
            f("profile").argtypes = [vp, c_double_p, C.c_int]
        if self.has("rng_probe") and self.prefix == "pdfb200_":
            f("rng_probe").argtypes = [C.c_int, C.c_uint64, C.c_int64, C.POINTER(C.c_uint64), C.c_uint32, C.c_uint32, C.c_uint32, c_double_p]
        if self.has("jit_source"):
            f("jit_source").argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, c_int32_p, c_int32_p, C.c_char_p, C.c_long]
            f("jit_source").restype = C.c_long
        if self.has("jit_compile"):
            f("jit_compile").argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, c_int32_p, c_int32_p, C.c_int, C.c_char_p, C.c_long]
            f("jit_compile").restype = C.c_long

    def check(self, status: int, what: str):
        if status != 0:
            msg = self.fn("last_error")()
            raise PdfError(f"{self.prefix}{what} failed ({status}): {msg.decode() if msg else ''}")


def jit_source(lib: "Library", params: Params, axisym: bool, has_open: bool, tet_bits: int = 0,
               geometric_d=(1, 1, 1), solution_d=(1, 1, 1), cell_faces: int = 6) -> str:
    """The translation unit the library compiles at run time for these parameters
    and mesh flags (pdffoam_b200/csrc/jit.cu); host only."""
    buf = C.create_string_buffer(1 << 16)
    g = np.asarray(geometric_d, np.int32); sd = np.asarray(solution_d, np.int32)
    n = lib.fn("jit_source")(C.byref(params), int(axisym), int(has_open), int(tet_bits), int(cell_faces), _ip(g), _ip(sd), buf, len(buf))
    if n < 0:
        raise PdfError("jit_source: buffer too small")
    return buf.value.decode()


_DEVICE_LIB: Optional[Library] = None


def device_library_path() -> str:
    # PDFB200_LIB selects another build of the same library (kernel tuning experiments)
    return os.environ.get("PDFB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libpdfb200.so")


def DeviceLibrary() -> Library:
    """The CUDA library. No fallback: a missing build is an error."""
    global _DEVICE_LIB
    if _DEVICE_LIB is None:
        path = device_library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pdffoam_b200 has no CPU fallback)")
        _DEVICE_LIB = Library(path, "pdfb200_")
    return _DEVICE_LIB


class CloudHandle:
    """Thin object wrapper around one opaque cloud handle of `lib`."""

    def __init__(self, lib: Library, mesh: PolyMesh, params: Params, device: int = 0, capacity: int = 0):
        self.lib = lib
        self.mesh = mesh
        self.params = params
        self.n_phi = params.nPhi
        self.n_cov = params.nPhi * (params.nPhi + 1) // 2
        self._keep: List = []
        ms = self._mesh_struct(mesh)
        h = C.c_void_p()
        lib.check(lib.fn("create")(C.byref(ms), C.byref(params), device, capacity, C.byref(h)), "create")
        self.h = h

    def _mesh_struct(self, mesh: PolyMesh) -> MeshStruct:
        ms = MeshStruct()
        ms.nPoints, ms.nFaces, ms.nInternalFaces = mesh.n_points, mesh.n_faces, mesh.n_internal_faces
        ms.nCells, ms.nPatches = mesh.n_cells, len(mesh.patches)
        arrs = dict(
            points=np.ascontiguousarray(mesh.points, dtype=np.float64),
            faceStart=np.ascontiguousarray(mesh.face_start, dtype=np.int32),
            facePoints=np.ascontiguousarray(mesh.face_points, dtype=np.int32),
            owner=np.ascontiguousarray(mesh.owner, dtype=np.int32),
            neighbour=np.ascontiguousarray(mesh.neighbour, dtype=np.int32),
            patchStart=np.array([p.start for p in mesh.patches], dtype=np.int32),
            patchSize=np.array([p.size for p in mesh.patches], dtype=np.int32),
            patchGeom=np.array([p.geom for p in mesh.patches], dtype=np.int32),
            patchHandler=np.array([p.handler for p in mesh.patches], dtype=np.int32),
            patchReflecting=np.array([p.reflecting for p in mesh.patches], dtype=np.int32),
        )
        self._keep.append(arrs)
        ms.points = _dp(arrs["points"])
        for k in ("faceStart", "facePoints", "owner", "neighbour", "patchStart", "patchSize", "patchGeom", "patchHandler", "patchReflecting"):
            setattr(ms, k, _ip(arrs[k]))
        ms.geometricD = (C.c_int32 * 3)(*mesh.geometric_d)
        ms.solutionD = (C.c_int32 * 3)(*mesh.solution_d)
        return ms

    def close(self):
        if self.h:
            self.lib.fn("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration / fields ----
    def set_params(self, params: Params):
        self.lib.check(self.lib.fn("set_params")(self.h, C.byref(params)), "set_params")
        self.params = params

    def set_flamelet_tables(self, chi: np.ndarray, z: np.ndarray, fields: Sequence[np.ndarray]):
        chi = np.ascontiguousarray(chi, np.float64); z = np.ascontiguousarray(z, np.float64)
        fl = [np.ascontiguousarray(f, np.float64) for f in fields]
        for f in fl:
            assert f.shape == (chi.size, z.size)
        arr = (c_double_p * len(fl))(*[_dp(f) for f in fl])
        self.lib.check(self.lib.fn("set_flamelet_tables")(self.h, chi.size, z.size, _dp(chi), _dp(z), len(fl), arr), "set_flamelet_tables")

    def set_autoignition_tables(self, z: np.ndarray, pv: np.ndarray, cEq: np.ndarray, fields: Sequence[np.ndarray]):
        """constant/autoIgnition/* of mcKulkarniAutoIgnitionReactionModel: fields = [cdot, rho, added...], each [nZ][nPv]"""
        z = np.ascontiguousarray(z, np.float64); pv = np.ascontiguousarray(pv, np.float64); cEq = np.ascontiguousarray(cEq, np.float64)
        fs = [np.ascontiguousarray(f, np.float64) for f in fields]
        for f in fs:
            assert f.shape == (z.size, pv.size), f.shape
        arr = (c_double_p * len(fs))(*[_dp(f) for f in fs])
        self.lib.check(self.lib.fn("set_autoignition_tables")(self.h, z.size, pv.size, _dp(z), _dp(pv), _dp(cEq), len(fs), arr), "set_autoignition_tables")

    def upload_fv_fields(self, fv: Dict[str, np.ndarray]):
        """fv keys: U,Ub,UbFixed,p,pb,k,kb,kbFixed,epsilon,epsilonb,R,Rb."""
        s = FvFields()
        keep = {}
        for name, _ in FvFields._fields_:
            a = fv[name]
            if name.endswith("Fixed"):
                a = np.ascontiguousarray(a, np.uint8); setattr(s, name, _u8p(a))
            else:
                a = np.ascontiguousarray(a, np.float64); setattr(s, name, _dp(a))
            keep[name] = a
        self.lib.check(self.lib.fn("upload_fv_fields")(self.h, C.byref(s)), "upload_fv_fields")

    def stage_fv_fields(self, fv: Dict[str, np.ndarray]):
        """Start the asynchronous copy of the next FV state (pdfb200_stage_fv_fields); the arrays must be contiguous,
        of the right dtype (no conversion copy is made here) and stay untouched until commit_fv_fields() returns."""
        s = FvFields()
        for name, _ in FvFields._fields_:
            a = fv[name]
            want = np.uint8 if name.endswith("Fixed") else np.float64
            if a.dtype != want or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"stage_fv_fields: {name} must be a contiguous {want.__name__} array")
            setattr(s, name, _u8p(a) if want is np.uint8 else _dp(a))
        self._staged_keepalive = fv
        self.lib.check(self.lib.fn("stage_fv_fields")(self.h, C.byref(s)), "stage_fv_fields")

    def commit_fv_fields(self):
        self.lib.check(self.lib.fn("commit_fv_fields")(self.h), "commit_fv_fields")
        self._staged_keepalive = None

    def upload_cloud_fields(self, cf: Dict):
        """cf keys: rho,rhob,rhobFixed, Phi[list],Phib[list],PhibFixed[list], optional Cov,Covb,CovbFixed lists."""
        s = CloudFields()
        keep = []

        def d(a):
            a = np.ascontiguousarray(a, np.float64); keep.append(a); return _dp(a)

        def u(a):
            a = np.ascontiguousarray(a, np.uint8); keep.append(a); return _u8p(a)

        s.rho, s.rhob, s.rhobFixed = d(cf["rho"]), d(cf["rhob"]), u(cf["rhobFixed"])
        for i in range(self.n_phi):
            s.Phi[i], s.Phib[i], s.PhibFixed[i] = d(cf["Phi"][i]), d(cf["Phib"][i]), u(cf["PhibFixed"][i])
        cov = cf.get("Cov")
        for i in range(self.n_cov):
            if cov is not None and cov[i] is not None:
                s.Cov[i], s.Covb[i], s.CovbFixed[i] = d(cov[i]), d(cf["Covb"][i]), u(cf["CovbFixed"][i])
        self.lib.check(self.lib.fn("upload_cloud_fields")(self.h, C.byref(s)), "upload_cloud_fields")

    def upload_poscorr_fields(self, f: Dict):
        """Fields of an externally solved position correction (pdfb200_poscorr_fields): keys G0, G1, G2 ([nCells,3]),
        l, zeta ([nCells]) with their boundary values under the key + "b", and "scale"; missing pairs count as zero."""
        s = PoscorrFields()
        keep = []
        for k in ("G0", "G1", "G2", "l", "zeta"):
            if f.get(k) is not None:
                for kk in (k, k + "b"):
                    a = np.ascontiguousarray(f[kk], np.float64); keep.append(a); setattr(s, kk, _dp(a))
        s.scale = float(f.get("scale", 0.0))
        self.lib.check(self.lib.fn("upload_poscorr_fields")(self.h, C.byref(s)), "upload_poscorr_fields")

    def init_moments(self):
        self.lib.check(self.lib.fn("init_moments")(self.h), "init_moments")

    # ---- particles ----
    def seed_particles(self):
        self.lib.check(self.lib.fn("seed_particles")(self.h), "seed_particles")

    def num_particles(self) -> int:
        return int(self.lib.fn("num_particles")(self.h))

    def upload_particles(self, P: Dict[str, np.ndarray]):
        n = P["m"].shape[0]
        s = ParticlesStruct()
        keep = []

        def d(a):
            a = np.ascontiguousarray(a, np.float64); keep.append(a); return _dp(a)

        s.pos, s.U = d(P["pos"]), d(P["U"])
        s.m, s.Omega, s.rho, s.eta = d(P["m"]), d(P["Omega"]), d(P["rho"]), d(P["eta"])
        for i in range(self.n_phi):
            s.Phi[i] = d(P["Phi"][i])
        cell = np.ascontiguousarray(P["cell"], np.int32); keep.append(cell); s.cell = _ip(cell)
        if P.get("tet") is not None:
            tet = np.ascontiguousarray(P["tet"], np.int32); keep.append(tet); s.tet = _ip(tet)
        if P.get("id") is not None:
            pid = np.ascontiguousarray(P["id"], np.uint64); keep.append(pid); s.id = _u64p(pid)
        self.lib.check(self.lib.fn("upload_particles")(self.h, n, C.byref(s)), "upload_particles")

    def download_particles(self) -> Dict[str, np.ndarray]:
        n = self.num_particles()
        P = dict(pos=np.zeros((n, 3)), U=np.zeros((n, 3)), m=np.zeros(n), Omega=np.zeros(n), rho=np.zeros(n),
                 eta=np.zeros(n), Phi=[np.zeros(n) for _ in range(self.n_phi)], cell=np.zeros(n, np.int32),
                 tet=np.zeros(n, np.int32), id=np.zeros(n, np.uint64))
        s = ParticlesStruct()
        s.pos, s.U, s.m, s.Omega, s.rho, s.eta = _dp(P["pos"]), _dp(P["U"]), _dp(P["m"]), _dp(P["Omega"]), _dp(P["rho"]), _dp(P["eta"])
        for i in range(self.n_phi):
            s.Phi[i] = _dp(P["Phi"][i])
        s.cell, s.tet, s.id = _ip(P["cell"]), _ip(P["tet"]), _u64p(P["id"])
        nn = C.c_int64(0)
        self.lib.check(self.lib.fn("download_particles")(self.h, n, C.byref(s), C.byref(nn)), "download_particles")
        assert nn.value == n
        return P

    # ---- stepping ----
    def evolve(self, n_steps: int = 1) -> List[StepReport]:
        reps = (StepReport * n_steps)()
        self.lib.check(self.lib.fn("evolve")(self.h, n_steps, reps), "evolve")
        return list(reps)

    def delta_t(self) -> float:
        return float(self.lib.fn("delta_t")(self.h))

    def set_delta_t(self, dt: float):
        self.lib.check(self.lib.fn("set_delta_t")(self.h, dt), "set_delta_t")

    def download_rho(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Only the mean density, the one field fed back to the FV equations
        (pdfSimpleFoam/createFields.H:18); `out` may be a pinned host buffer."""
        if out is None:
            out = np.zeros(self.mesh.n_cells)
        s = CellFieldsStruct()
        s.rho = _dp(out)
        self.lib.check(self.lib.fn("download_cell_fields")(self.h, C.byref(s)), "download_cell_fields")
        return out

    def download_rho_async(self, out: np.ndarray):
        """pdfb200_download_cell_fields_async for the mean density: queued behind the steps issued so far, lands in `out`
        (a page-locked buffer) while the next evolve() runs; sync_downloads() waits for it."""
        s = CellFieldsStruct()
        s.rho = _dp(out)
        self.lib.check(self.lib.fn("download_cell_fields_async")(self.h, C.byref(s)), "download_cell_fields_async")

    def download_fields_async(self, out: Dict[str, np.ndarray]):
        """pdfb200_download_cell_fields_async for any of the plain cell fields (rho, rhoInst, pnd, pndInst, U, Tau, k,
        PaNIC): out maps the field name to a contiguous float64 buffer (page-locked for a truly asynchronous copy)."""
        s = CellFieldsStruct()
        for k, a in out.items():
            if k not in ("rho", "rhoInst", "pnd", "pndInst", "U", "Tau", "k", "PaNIC"):
                raise KeyError(k)
            setattr(s, k, _dp(a))
        self.lib.check(self.lib.fn("download_cell_fields_async")(self.h, C.byref(s)), "download_cell_fields_async")

    def profile_ctas(self) -> np.ndarray:
        """pdfb200_profile_ctas: [nCtas, 3] = start ns, end ns, SM of every CTA of the last particle-kernel launch
        (empty on the first call, which switches the recording on)."""
        out = np.zeros((4096, 3), np.int64)
        n = C.c_int64(0)
        f = self.lib.fn("profile_ctas")
        f.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        self.lib.check(f(self.h, 4096, out.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n)), "profile_ctas")
        return out[: n.value]

    def sync_downloads(self):
        self.lib.check(self.lib.fn("sync_downloads")(self.h), "sync_downloads")

    def download_cell_fields(self, moments: bool = True) -> Dict[str, np.ndarray]:
        nc = self.mesh.n_cells
        out = dict(rho=np.zeros(nc), rhoInst=np.zeros(nc), pnd=np.zeros(nc), pndInst=np.zeros(nc), U=np.zeros((nc, 3)),
                   Tau=np.zeros((nc, 6)), k=np.zeros(nc), Phi=[np.zeros(nc) for _ in range(self.n_phi)],
                   Cov=[np.zeros(nc) for _ in range(self.n_cov)], PaNIC=np.zeros(nc),
                   UPhi=[np.zeros((nc, 3)) for _ in range(self.n_phi)])
        if moments:
            out.update(mMom=np.zeros(nc), VMom=np.zeros(nc), UMom=np.zeros((nc, 3)), UUMom=np.zeros((nc, 6)),
                       PhiMom=[np.zeros(nc) for _ in range(self.n_phi)], PhiPhiMom=[np.zeros(nc) for _ in range(self.n_cov)],
                       UPhiMom=[np.zeros((nc, 3)) for _ in range(self.n_phi)])
        s = CellFieldsStruct()
        for k in ("rho", "rhoInst", "pnd", "pndInst", "U", "Tau", "k", "PaNIC"):
            setattr(s, k, _dp(out[k]))
        for i in range(self.n_phi):
            s.Phi[i] = _dp(out["Phi"][i])
            s.UPhi[i] = _dp(out["UPhi"][i])
        for i in range(self.n_cov):
            s.Cov[i] = _dp(out["Cov"][i])
        if moments:
            for k in ("mMom", "VMom", "UMom", "UUMom"):
                setattr(s, k, _dp(out[k]))
            for i in range(self.n_phi):
                s.PhiMom[i] = _dp(out["PhiMom"][i])
                s.UPhiMom[i] = _dp(out["UPhiMom"][i])
            for i in range(self.n_cov):
                s.PhiPhiMom[i] = _dp(out["PhiPhiMom"][i])
        self.lib.check(self.lib.fn("download_cell_fields")(self.h, C.byref(s)), "download_cell_fields")
        return out

    def upload_moments(self, cf: Dict[str, np.ndarray]):
        """Restart from the moment fields of a previous run (pdfb200_upload_moments); cf as returned by
        download_cell_fields(moments=True). Missing keys are passed as NULL."""
        s = CellFieldsStruct()
        keep = []

        def d(a):
            a = np.ascontiguousarray(a, np.float64); keep.append(a); return _dp(a)
        for k in ("mMom", "VMom", "UMom", "UUMom", "PaNIC"):
            if cf.get(k) is not None:
                setattr(s, k, d(cf[k]))
        for name, n in (("PhiMom", self.n_phi), ("PhiPhiMom", self.n_cov), ("UPhiMom", self.n_phi)):
            if cf.get(name) is not None:
                for i in range(n):
                    getattr(s, name)[i] = d(cf[name][i])
        self.lib.check(self.lib.fn("upload_moments")(self.h, C.byref(s)), "upload_moments")

    def step_counter(self) -> int:
        return int(self.lib.fn("step_counter")(self.h))

    def set_graph_mode(self, mode: int):
        """-1 automatic, 0 plain launches, 1 CUDA-graph replay of the step wherever possible (pdfb200_set_graph_mode)"""
        self.lib.check(self.lib.fn("set_graph_mode")(self.h, int(mode)), "set_graph_mode")

    def id_counter(self) -> int:
        return int(self.lib.fn("id_counter")(self.h))

    def set_id_counter(self, next_id: int):
        self.lib.check(self.lib.fn("set_id_counter")(self.h, int(next_id)), "set_id_counter")

    def set_step_counter(self, step: int):
        self.lib.check(self.lib.fn("set_step_counter")(self.h, int(step)), "set_step_counter")

    # ---- introspection ----
    def mesh_info(self) -> MeshInfo:
        mi = MeshInfo()
        self.lib.check(self.lib.fn("mesh_info_get")(self.h, C.byref(mi)), "mesh_info_get")
        return mi

    def download_tets(self) -> Dict[str, np.ndarray]:
        nt = int(self.mesh_info().nTets)
        out = dict(gradN=np.zeros((nt, 3, 3)), nbr=np.zeros((nt, 4), np.int32), face=np.zeros(nt, np.int32),
                   ptB=np.zeros(nt, np.int32), ptC=np.zeros(nt, np.int32), cell=np.zeros(nt, np.int32))
        self.lib.check(self.lib.fn("download_tets")(self.h, _dp(out["gradN"]), _ip(out["nbr"]), _ip(out["face"]),
                                                    _ip(out["ptB"]), _ip(out["ptC"]), _ip(out["cell"])), "download_tets")
        return out

    def download_geometry(self) -> Dict[str, np.ndarray]:
        nf, nc = self.mesh.n_faces, self.mesh.n_cells
        out = dict(faceCentres=np.zeros((nf, 3)), faceAreas=np.zeros((nf, 3)), cellCentres=np.zeros((nc, 3)),
                   cellVolumes=np.zeros(nc), courantCoeffs=np.zeros((nf, 3)), hNum=np.zeros(nc))
        self.lib.check(self.lib.fn("download_geometry")(self.h, _dp(out["faceCentres"]), _dp(out["faceAreas"]),
                                                        _dp(out["cellCentres"]), _dp(out["cellVolumes"]),
                                                        _dp(out["courantCoeffs"]), _dp(out["hNum"])), "download_geometry")
        return out

    def kernel_launches(self) -> int:
        return int(self.lib.fn("kernel_launches")(self.h))

    def last_evolve_ms(self) -> float:
        return float(self.lib.fn("last_evolve_ms")(self.h))

    def profile(self, reset: bool = False) -> Dict[str, float]:
        out = np.zeros(8)
        self.lib.check(self.lib.fn("profile")(self.h, _dp(out), int(reset)), "profile")
        names = ("prep_ms", "evolve_kernel_ms", "sort_ms", "moments_ms", "allreduce_ms", "finalise_ms", "steps", "particle_steps")
        return dict(zip(names, out.tolist()))

    def comm_init(self, n_ranks: int, rank: int, uid: bytes):
        self.lib.check(self.lib.fn("comm_init")(self.h, n_ranks, rank, uid), "comm_init")
