"""The OpenFOAM-side adaptor (integration/mcDeviceCloud.{H,C}: the class a maintainer adds to mcParticle/ so that
pdfSimpleFoam drives libpdfb200.so through the public surface of mcParticleCloud) compiled against stand-in OpenFOAM
headers and driven by integration/adaptorDemo.C the way mcThermo drives the cloud (mcThermo.C:114-143).

CPU: it compiles and links against libpdfb200.so, reads the reference's dictionaries (tutorial-style
<cloud>Properties / <cloud>Solution text) into exactly the parameters the Python binding uses, and fails with
OpenFOAM's FatalError text when there is no device. GPU: the run through the adaptor reproduces the run through the
Python binding bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from pdffoam_b200 import cases, meshgen
from pdffoam_b200.capi import Params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "integration", "_build", "adaptorDemo")

_INV = lambda d: {v: k for k, v in d.items() if k not in ("off", "false")}


def _build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "integration")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _dict_text(case, names):
    """<cloud>Properties and <cloud>Solution in the reference's dictionary syntax (tutorials/SandiaD/constant/
    thermophysicalProperties:11-77, system/mcThermoCloudSolution:11-53) from the Python-side parameters"""
    P = case.params
    r = repr
    scheme = _INV(cases.INTERPOLATION_SCHEMES)
    handler = {meshgen.BC_SLIP: "slip", meshgen.BC_EMPTY: "empty", meshgen.BC_OPEN: "open", meshgen.BC_INLET_OUTLET: "inletOutlet"}
    lts = {0: "off", 1: "cell", 2: "particle"}[P.localTimeStepping]
    thermo = [
        f"velocityModel      {_INV(cases.VELOCITY_MODELS)[P.velocityModel]};",
        f"positionCorrection {_INV(cases.POSITION_CORRECTIONS)[P.positionCorrection]};",
        f"localTimeStepping  {lts};  // comment",
        f"OmegaModel         {_INV(cases.OMEGA_MODELS)[P.omegaModel]};",
        f"mixingModel        {_INV(cases.MIXING_MODELS)[P.mixingModel]};",
        f"reactionModel      {_INV(cases.REACTION_MODELS)[P.reactionModel]};",
        "scalarFields       ( " + " ".join(names) + " );",
        "mixedScalars       ( " + " ".join(names[P.mixed[i]] for i in range(P.nMixed)) + " );",
        "conservedScalars   ( " + " ".join(names[P.conserved[i]] for i in range(P.nConserved)) + " );",
        f"coldReactionModelCoeffs {{ density {r(P.coldDensity)}; }}",
        "/* block comment */",
    ]
    if P.reactionModel == cases.REACTION_MODELS["SteadyFlamelet"]:
        thermo.append("SteadyFlameletReactionModelCoeffs\n{\n    scalars ( " + " ".join(names[P.addedIdx[i]] for i in range(P.nAdded)) + f" );\n    Cchi {r(P.Cchi)};\n}}")
        thermo.append(f"chiName {names[P.chiIdx]};")
    if P.reactionModel == cases.REACTION_MODELS["BurkeSchumann"]:
        thermo.append(f"BurkeSchumannReactionModelCoeffs {{ Rfuel {r(P.bsRfuel)}; Rox {r(P.bsRox)}; Rstoi {r(P.bsRstoi)}; Tfuel {r(P.bsTfuel)}; "
                      f"Tox {r(P.bsTox)}; Tstoi {r(P.bsTstoi)}; zstoi {r(P.bsZstoi)}; }}")
        thermo.append(f"TName {names[P.TIdx]};")
    bh = ["boundaryHandlers\n{"]
    for p in case.mesh.patches:
        body = f"type {handler[p.handler]};"
        if p.handler in (meshgen.BC_OPEN, meshgen.BC_INLET_OUTLET):
            body += f" reflecting {'on' if p.reflecting else 'off'};"
        if p.handler == meshgen.BC_INLET_OUTLET:
            body += f" randomGenerator {{ type {_INV(cases.INLET_RANDOM)[P.inletRandom]}; }}"
        bh.append(f"    {p.name} {{ {body} }}")
    bh.append("}")
    thermo += bh
    mix_coeffs = "modifiedCurlMixingModelCoeffs" if P.mixingModel == 2 else "IEMMixingModelCoeffs"
    pc_coeffs = "simplePositionCorrectionCoeffs" if P.positionCorrection == 2 else "limitedSimplePositionCorrectionCoeffs"
    lts_coeffs = "particleLocalTimeSteppingCoeffs" if P.localTimeStepping == 2 else "cellLocalTimeSteppingCoeffs"
    zz = names[P.zIdx] + names[P.zIdx] + "Cov"
    sol = [
        f"CFL {r(P.CFL)};", f"averagingCoeff {r(P.averagingCoeff)};", f"particlesPerCell {P.particlesPerCell};",
        f"particleNumberControl {'on' if P.particleNumberControl else 'off'};", f"cloneAt {r(P.cloneAt)};", f"eliminateAt {r(P.eliminateAt)};",
        f"kMin {r(P.kMin)};", f"DNum {r(P.DNum)};",
        f"SLMFullVelocityModelCoeffs {{ C0 {r(P.C0)}; C1 {r(P.C1)}; }}",
        f"{mix_coeffs} {{ Cmix {r(P.Cmix)}; }}",
        f"RASOmegaModelCoeffs {{ Omega0 {r(P.Omega0)}; }}",
        f"{lts_coeffs} {{ upperBound {r(P.ltsUpperBound)}; }}",
        f"{pc_coeffs} {{ C {r(P.pcC)}; CFL {r(P.pcCFL)}; }}",
        f"relaxationTimes {{ default 1e11; k {r(P.relaxTimek)}; U {r(P.relaxTimeU)}; }}",
        f"minRelaxationFactor {r(P.minRelaxTime)}; fvDeltaT 1;",
        "interpolationSchemes\n{\n" + "\n".join(f"    {k} {scheme[v]};" for k, v in (
            ("mcRASOmegaModel::Omega", P.interpOmega), ("U", P.interpU), ("k", P.interpk), ("SLMFullVelocityModel::diffU", P.interpDiffU),
            ("UPosCorr", P.interpUPosCorr), (zz, P.interpZVar), ("mcCellLocaltimeStepping::eta", P.interpEta))) + "\n}",
        f"deviceCoeffs {{ k0 {r(P.k0)}; deltaT0 {r(P.deltaT0)}; seed {P.seed}; }}",
    ]
    return "\n".join(thermo) + "\n", "\n".join(sol) + "\n"


def dump_case(case, d, names, n_steps):
    m = case.mesh
    os.makedirs(d, exist_ok=True)
    n_chi = case.tables["chi"].size if case.tables else 0
    n_z = case.tables["z"].size if case.tables else 0
    with open(os.path.join(d, "sizes.txt"), "w") as f:
        f.write(f"{m.n_points} {m.n_faces} {m.n_internal_faces} {m.n_cells} {len(m.patches)} {n_steps} {n_chi} {n_z}\n")
        f.write(" ".join(str(int(x)) for x in m.geometric_d) + "\n" + " ".join(str(int(x)) for x in m.solution_d) + "\n")
    np.ascontiguousarray(m.points, np.float64).tofile(os.path.join(d, "points.f64"))
    for name, a in (("faceStart", m.face_start), ("facePoints", m.face_points), ("owner", m.owner), ("neighbour", m.neighbour)):
        np.ascontiguousarray(a, np.int32).tofile(os.path.join(d, name + ".i32"))
    cls = {meshgen.PATCH_GENERIC: "generic", meshgen.PATCH_EMPTY: "empty", meshgen.PATCH_WEDGE: "wedge"}
    with open(os.path.join(d, "patches.txt"), "w") as f:
        for p in m.patches:
            f.write(f"{p.name} {p.start} {p.size} {cls[p.geom]}\n")
    for k, a in case.fv.items():
        np.ascontiguousarray(a, np.uint8 if k.endswith("Fixed") else np.float64).tofile(os.path.join(d, k + (".u8" if k.endswith("Fixed") else ".f64")))
    np.ascontiguousarray(case.cf["rho"]).tofile(os.path.join(d, "rho.f64")); np.ascontiguousarray(case.cf["rhob"]).tofile(os.path.join(d, "rhob.f64"))
    np.ascontiguousarray(case.cf["rhobFixed"], np.uint8).tofile(os.path.join(d, "rhobFixed.u8"))
    for i, n in enumerate(names):
        np.ascontiguousarray(case.cf["Phi"][i]).tofile(os.path.join(d, f"Phi_{n}.f64"))
        np.ascontiguousarray(case.cf["Phib"][i]).tofile(os.path.join(d, f"Phib_{n}.f64"))
        np.ascontiguousarray(case.cf["PhibFixed"][i], np.uint8).tofile(os.path.join(d, f"PhibFixed_{n}.u8"))
    if case.tables:
        case.tables["chi"].tofile(os.path.join(d, "chi.f64")); case.tables["z"].tofile(os.path.join(d, "z.f64"))
        with open(os.path.join(d, "tables.txt"), "w") as f:
            for i, t in enumerate(case.tables["fields"]):
                np.ascontiguousarray(t, np.float64).tofile(os.path.join(d, f"table{i}.f64")); f.write(f"table{i}.f64\n")
    thermo, sol = _dict_text(case, names)
    open(os.path.join(d, "thermoCloudProperties"), "w").write(thermo)
    open(os.path.join(d, "thermoCloudSolution"), "w").write(sol)


def _cases():
    c = cases.synthetic_box(6, 4, 3, ppc=12, closed=False, number_control=True)
    c.params.deltaT0 = 2e-3
    # the synthetic case carries no covariance field for zz: nothing to upload (the adaptor finds none in the registry)
    yield "box", c, ["z"]
    gold = os.path.join(ROOT, "tests", "golden")
    t = cases.tutorial_case("SandiaD", ppc=10)
    t.params.deltaT0 = 2e-6
    t.cf.pop("Cov"); t.cf.pop("Covb"); t.cf.pop("CovbFixed")
    yield "SandiaD-shaped", t, ["z", "chi", "T"]


def _read_params(d):
    out = {}
    for line in open(os.path.join(d, "params_read.txt")):
        k, v = line.split()
        out[k] = float(v)
    return out


@pytest.mark.parametrize("idx", [0, 1])
def test_adaptor_reads_reference_dictionaries_into_the_same_parameters(tmp_path, idx):
    _build()
    name, case, names = list(_cases())[idx]
    dump_case(case, str(tmp_path), names, 2)
    r = subprocess.run([DEMO, str(tmp_path)], capture_output=True, text=True)
    got = _read_params(str(tmp_path))
    P = case.params
    skip = {"mixed", "conserved", "addedIdx", "rngMode", "TIdx", "chiIdx", "cIdx", "pvIdx"}
    # coefficients of models the case does not select are not in its dictionaries
    if P.reactionModel != cases.REACTION_MODELS["SteadyFlamelet"]:
        skip |= {"Cchi", "nAdded"}
    if P.reactionModel != cases.REACTION_MODELS["BurkeSchumann"]:
        skip |= {f for f, _ in Params._fields_ if f.startswith("bs")}
    for fname, _ in Params._fields_:
        if fname in skip:
            continue
        assert got[fname] == float(getattr(P, fname)), f"{name}: {fname}: dictionary gave {got[fname]!r}, Python side has {getattr(P, fname)!r}"
    for arr, n in (("mixed", P.nMixed), ("conserved", P.nConserved), ("addedIdx", P.nAdded)):
        for i in range(n):
            assert got[f"{arr}{i}"] == getattr(P, arr)[i]
    import torch
    if not torch.cuda.is_available():
        # no device: the adaptor turns the library's status into OpenFOAM's fatal error
        assert r.returncode == 1 and "FOAM FATAL ERROR" in r.stderr and "mcDeviceCloud::createHandle()" in r.stderr


def test_adaptor_rejects_unknown_model_words_like_the_reference(tmp_path):
    _build()
    name, case, names = next(_cases())
    dump_case(case, str(tmp_path), names, 1)
    p = os.path.join(str(tmp_path), "thermoCloudProperties")
    text = open(p).read().replace("velocityModel      SLMFull;", "velocityModel      SLMFuller;")
    open(p, "w").write(text)
    r = subprocess.run([DEMO, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 1
    assert "Unknown mcVelocityModel type SLMFuller" in r.stderr and "Valid mcVelocityModel types are" in r.stderr and "SLMFull" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_run_through_the_adaptor_equals_run_through_the_binding(device_lib, tmp_path, idx):
    _build()
    name, case, names = list(_cases())[idx]
    n_steps = 4
    dump_case(case, str(tmp_path), names, n_steps)
    r = subprocess.run([DEMO, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Cloud thermoCloud" in r.stdout and "pmd relative error" in r.stdout
    h = case.make_cloud(device_lib)
    h.seed_particles()
    reps = h.evolve(n_steps)
    cf = h.download_cell_fields()
    rho = np.fromfile(os.path.join(str(tmp_path), "rho_out.f64"))
    U = np.fromfile(os.path.join(str(tmp_path), "U_out.f64")).reshape(-1, 3)
    assert np.array_equal(rho, cf["rho"]), f"{name}: rho through the adaptor differs"
    assert np.array_equal(U, cf["U"]), f"{name}: UCloudPDF through the adaptor differs"
    for i, n in enumerate(names):
        assert np.array_equal(np.fromfile(os.path.join(str(tmp_path), f"Phi_out_{n}.f64")), cf["Phi"][i]), n
    rows = [l.split() for l in open(os.path.join(str(tmp_path), "report.txt"))]
    assert len(rows) == n_steps
    for row, rep in zip(rows, reps):
        assert float(row[0]) == rep.rhoRes and float(row[2]) == rep.deltaTNext and int(row[3]) == rep.nParticles
