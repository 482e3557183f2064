"""OpenFOAM ASCII checkpoint files of the particle cloud: the per-field `lagrangian/<cloud>/*` files that
mcParticle::writeFields / readFields produce and consume (mcParticle/mcParticleIO.C:94-164 read, :167-234 write:
positions, origId, m, UParticle, Ucorrection, Omega, rho, eta and one file per scalar field) and the time-averaged
moment fields the cloud's constructor reads READ_IF_PRESENT (mcParticleCloud.C:339-517, checkMoments :739-820:
mMoment, VMoment, UMoment, UUMoment, <Phi>Moment, <Phi><Phi>Moment) as DimensionedField internal fields.

Inside OpenFOAM the adaptor of INTEGRATION.md lets OpenFOAM's own IOField / IOPosition do this and only moves the
columns through pdfb200_download_particles / pdfb200_upload_particles; this module is the same format for the
stand-alone driver and the tests, so that a run can be restarted from, or handed over to, a pdfFoam case directory.
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Optional, Sequence

import numpy as np

_HEADER = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    location    "{location}";
    object      {obj};
}}
"""


def _fmt(v: float) -> str:
    return repr(float(v))          # shortest string that round-trips the double


def _write_list(f, data: np.ndarray):
    data = np.asarray(data)
    f.write(f"{data.shape[0]}\n(\n")
    if data.ndim == 1:
        if data.dtype.kind in "iu":
            f.write("\n".join(str(int(x)) for x in data))
        else:
            f.write("\n".join(_fmt(x) for x in data))
    else:
        f.write("\n".join("(" + " ".join(_fmt(x) for x in row) + ")" for row in data))
    f.write("\n)\n")


_CLASSES = {1: "scalarField", 3: "vectorField", 6: "symmTensorField"}


def write_field(path: str, obj: str, data: np.ndarray, location: str, cls: Optional[str] = None):
    """An IOField<Type> file (scalarField / vectorField / symmTensorField / labelField)."""
    data = np.asarray(data)
    if cls is None:
        cls = "labelField" if data.dtype.kind in "iu" else _CLASSES[1 if data.ndim == 1 else data.shape[1]]
    with open(path, "w") as f:
        f.write(_HEADER.format(cls=cls, location=location, obj=obj))
        f.write("\n")
        _write_list(f, data)


def _strip_header(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    m = re.search(r"FoamFile\s*\{.*?\}", text, flags=re.S)
    return text[m.end():] if m else text


def _parse_list(body: str, integer=False) -> np.ndarray:
    m = re.search(r"(\d+)\s*\(", body)
    if not m:
        raise ValueError("no list found")
    n = int(m.group(1))
    inner = body[m.end():body.rindex(")")]
    if "(" in inner:
        rows = re.findall(r"\(([^()]*)\)", inner)
        a = np.array([[float(x) for x in r.split()] for r in rows], dtype=np.float64)
    else:
        a = np.array(inner.split(), dtype=np.int64 if integer else np.float64)
    if a.shape[0] != n:
        raise ValueError(f"list announces {n} entries, holds {a.shape[0]}")
    return a


def read_field(path: str) -> np.ndarray:
    text = open(path).read()
    integer = bool(re.search(r"class\s+labelField", text))
    return _parse_list(_strip_header(text), integer)


# ---------------------------------------------------------------------------
# lagrangian/<cloud>/*
# ---------------------------------------------------------------------------
def write_lagrangian(case_dir: str, time_name: str, P: Dict, scalar_names: Sequence[str], cloud_name: str = "thermoCloud") -> str:
    """mcParticle::writeFields (mcParticleIO.C:167-234) + Cloud<>::writePositions: P as returned by
    CloudHandle.download_particles(). Returns the directory written."""
    loc = f"{time_name}/lagrangian/{cloud_name}"
    d = os.path.join(case_dir, loc)
    os.makedirs(d, exist_ok=True)
    n = P["m"].shape[0]
    with open(os.path.join(d, "positions"), "w") as f:
        f.write(_HEADER.format(cls="Cloud<mcParticle>", location=loc, obj="positions"))
        f.write(f"\n{n}\n(\n")
        f.write("\n".join("(" + " ".join(_fmt(x) for x in P["pos"][i]) + f") {int(P['cell'][i])}" for i in range(n)))
        f.write("\n)\n")
    write_field(os.path.join(d, "origId"), "origId", np.asarray(P["id"], np.int64), loc)
    write_field(os.path.join(d, "origProcId"), "origProcId", np.zeros(n, np.int64), loc)
    write_field(os.path.join(d, "m"), "m", P["m"], loc)
    write_field(os.path.join(d, "UParticle"), "UParticle", P["U"], loc)
    write_field(os.path.join(d, "Ucorrection"), "Ucorrection", np.zeros((n, 3)), loc)   # written for post-processing only (:198-202)
    write_field(os.path.join(d, "Omega"), "Omega", P["Omega"], loc)
    write_field(os.path.join(d, "rho"), "rho", P["rho"], loc)
    write_field(os.path.join(d, "eta"), "eta", P["eta"], loc)
    if len(scalar_names) != len(P["Phi"]):
        raise ValueError("one name per particle scalar is needed (scalarFields of the cloud dictionary)")
    for name, col in zip(scalar_names, P["Phi"]):
        write_field(os.path.join(d, name), name, col, loc)
    return d


def read_lagrangian(case_dir: str, time_name: str, scalar_names: Sequence[str], cloud_name: str = "thermoCloud") -> Dict:
    """mcParticle::readFields (mcParticleIO.C:94-164): every field is MUST_READ; the transient members
    (shift, Co, ghost, nSteps, the reflection flags) restart from zero as there. The tet label is not stored
    (upload_particles locates it)."""
    d = os.path.join(case_dir, time_name, "lagrangian", cloud_name)
    body = _strip_header(open(os.path.join(d, "positions")).read())
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    rows = re.findall(r"\(([^()]*)\)\s*(-?\d+)", body[m.end():])
    if len(rows) != n:
        raise ValueError(f"positions announces {n} particles, holds {len(rows)}")
    pos = np.array([[float(x) for x in r[0].split()] for r in rows]).reshape(n, 3)
    cell = np.array([int(r[1]) for r in rows], dtype=np.int32)

    def must(name):
        p = os.path.join(d, name)
        if not os.path.exists(p):
            raise FileNotFoundError(f"cannot find file {p} (MUST_READ)")
        a = read_field(p)
        if a.shape[0] != n:
            raise ValueError(f"Size of {name} field {a.shape[0]} does not match the number of particles {n}")   # checkFieldIOobject
        return a
    P = dict(pos=pos, cell=cell, U=must("UParticle").reshape(n, 3), m=must("m"), Omega=must("Omega"), rho=must("rho"), eta=must("eta"),
             Phi=[must(name) for name in scalar_names])
    pid = os.path.join(d, "origId")
    if os.path.exists(pid):
        P["id"] = read_field(pid).astype(np.uint64)
    return P


# ---------------------------------------------------------------------------
# the moment fields (internal fields of DimensionedField<Type, volMesh>)
# ---------------------------------------------------------------------------
def _moment_files(scalar_names: Sequence[str]) -> List:
    """(key in download_cell_fields(), index or None, file name, class, dimensions)"""
    out = [("mMom", None, "mMoment", "volScalarField::DimensionedInternalField", "[1 0 0 0 0 0 0]"),
           ("VMom", None, "VMoment", "volScalarField::DimensionedInternalField", "[0 3 0 0 0 0 0]"),
           ("UMom", None, "UMoment", "volVectorField::DimensionedInternalField", "[1 1 -1 0 0 0 0]"),
           ("UUMom", None, "UUMoment", "volSymmTensorField::DimensionedInternalField", "[1 2 -2 0 0 0 0]")]
    n = len(scalar_names)
    pp = 0
    for i in range(n):
        out.append(("PhiMom", i, f"{scalar_names[i]}Moment", "volScalarField::DimensionedInternalField", "[1 0 0 0 0 0 0]"))
    for i in range(n):
        for j in range(i, n):
            # mcParticleCloud.C:1749-1760 names the covariance fields <a><b>Cov; their moments append "Moment" (:782)
            out.append(("PhiPhiMom", pp, f"{scalar_names[i]}{scalar_names[j]}CovMoment", "volScalarField::DimensionedInternalField", "[1 0 0 0 0 0 0]"))
            pp += 1
    for i in range(n):
        out.append(("UPhiMom", i, f"U{scalar_names[i]}Moment", "volVectorField::DimensionedInternalField", "[1 1 -1 0 0 0 0]"))
    return out


def write_moments(case_dir: str, time_name: str, cf: Dict, scalar_names: Sequence[str]):
    d = os.path.join(case_dir, time_name)
    os.makedirs(d, exist_ok=True)
    for key, idx, fname, cls, dims in _moment_files(scalar_names):
        a = cf[key] if idx is None else cf[key][idx]
        kind = {1: "scalar", 3: "vector", 6: "symmTensor"}[1 if np.ndim(a) == 1 else np.shape(a)[1]]
        with open(os.path.join(d, fname), "w") as f:
            f.write(_HEADER.format(cls=cls, location=time_name, obj=fname))
            f.write(f"\ndimensions      {dims};\n\nvalue           nonuniform List<{kind}>\n")
            _write_list(f, np.asarray(a))
            f.write(";\n")


def read_moments(case_dir: str, time_name: str, scalar_names: Sequence[str]) -> Optional[Dict]:
    """The moment fields of a time directory, or None when one of the reference's fields is missing — checkMoments()
    then re-initialises all of them ("Moments are missing. Forced re-initialization.", mcParticleCloud.C:816-819).
    The flux moments (not a reference field) are optional."""
    d = os.path.join(case_dir, time_name)
    n = len(scalar_names)
    cf: Dict = dict(PhiMom=[None] * n, PhiPhiMom=[None] * (n * (n + 1) // 2), UPhiMom=[None] * n)
    for key, idx, fname, cls, dims in _moment_files(scalar_names):
        p = os.path.join(d, fname)
        if not os.path.exists(p):
            if key == "UPhiMom":
                cf["UPhiMom"] = None
                continue
            return None
        body = _strip_header(open(p).read())
        a = _parse_list(body[body.index("nonuniform"):])
        if idx is None:
            cf[key] = a
        elif cf[key] is not None:
            cf[key][idx] = a
    return cf
