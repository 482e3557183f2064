"""CPU-side checks of the drop-in boundary and the host logic: the C-ABI library
loads and exports every symbol include/pdfb200.h declares (no compute calls
without a GPU), the ctypes mirrors have the C struct sizes, and the mesh
generators produce valid polyMeshes."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from pdffoam_b200 import capi, cases, meshgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pdfb200.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pdfb200_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_python_symbol_lists_agree():
    assert _declared_symbols() == sorted("pdfb200_" + s for s in capi.API_SYMBOLS)


def test_device_library_exports_every_declared_symbol():
    path = capi.device_library_path()
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libpdfb200.so does not export {s}"
    # the oracle mirrors the same entry points (minus the device-only ones)
    from oracle.oracle import OracleLib
    ol = OracleLib()
    for s in capi.API_SYMBOLS:
        if s in ("comm_unique_id", "comm_init", "kernel_launches", "last_evolve_ms", "profile", "profile_ctas", "jit_source", "jit_compile"):
            continue
        assert ol.has(s), f"oracle does not export pdforacle_{s}"


def test_create_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device create() returns an error status."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.DeviceLibrary()
    with pytest.raises(capi.PdfError, match="no CPU fallback|CUDA|device"):
        capi.CloudHandle(lib, meshgen.box_mesh(2, 2, 2), cases.default_params(1), capacity=100)


def test_ctypes_struct_sizes_match_c(tmp_path):
    src = tmp_path / "sizes.c"
    names = ["pdfb200_mesh", "pdfb200_params", "pdfb200_fv_fields", "pdfb200_cloud_fields", "pdfb200_particles",
             "pdfb200_step_report", "pdfb200_cell_fields", "pdfb200_mesh_info", "pdfb200_poscorr_fields"]
    body = "".join(f'printf("%zu\\n", sizeof({n}));' for n in names)
    src.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void){{{body} printf("%zu\\n", offsetof(pdfb200_params, seed)); return 0;}}\n'
                   .replace("#include <stdio.h>", "#include <stdio.h>\n#include <stddef.h>"))
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    py = [capi.MeshStruct, capi.Params, capi.FvFields, capi.CloudFields, capi.ParticlesStruct, capi.StepReport,
          capi.CellFieldsStruct, capi.MeshInfo, capi.PoscorrFields]
    assert out[:-1] == [C.sizeof(t) for t in py]
    assert out[-1] == capi.Params.seed.offset


def test_model_selection_by_reference_names():
    p = cases.default_params(3, velocityModel="SLMFull", mixingModel="IEM", omegaModel="RAS", reactionModel="SteadyFlamelet",
                             positionCorrection="limitedSimple", localTimeStepping="cell", interpU="cellPointFace",
                             mixed=[0], conserved=[0], addedIdx=[2], zIdx=0, chiIdx=1)
    assert (p.velocityModel, p.mixingModel, p.omegaModel, p.reactionModel, p.positionCorrection, p.localTimeStepping) == (1, 1, 1, 3, 1, 1)
    assert p.nAdded == 1 and p.addedIdx[0] == 2 and p.nMixed == 1 and p.nConserved == 1
    for word in ("off", "false", "none"):                   # mcLocalTimeStepping.C:56-97 accepts these
        assert cases.default_params(1, localTimeStepping=word).localTimeStepping == 0
    with pytest.raises(KeyError):
        cases.default_params(1, reactionModel="Arrhenius")


def _check_mesh(mesh):
    nI = mesh.n_internal_faces
    assert np.all(mesh.owner[:nI] < mesh.neighbour)                  # owner < neighbour
    key = mesh.owner[:nI].astype(np.int64) * mesh.n_cells + mesh.neighbour
    assert np.all(np.diff(key) > 0)                                  # upper-triangular order
    start = nI
    for p in mesh.patches:
        assert p.start == start
        start += p.size
    assert start == mesh.n_faces
    # closed cells: sum of outward face area vectors vanishes
    fp = mesh.face_points.reshape(-1, 4)
    P = mesh.points[fp]
    Sf = 0.5 * np.cross(P[:, 2] - P[:, 0], P[:, 3] - P[:, 1])
    acc = np.zeros((mesh.n_cells, 3))
    np.add.at(acc, mesh.owner, Sf)
    np.add.at(acc, mesh.neighbour, -Sf[:nI])
    assert np.abs(acc).max() < 1e-12 * np.abs(Sf).max() * 10
    # owner-outward normals: (Cf - Cc) . Sf > 0
    Cf = P.mean(axis=1)
    assert np.all(np.einsum("ij,ij->i", Cf - mesh.cell_centres_simple()[mesh.owner], Sf) > 0)


def test_box_slab_and_wedge_meshes_are_valid():
    _check_mesh(meshgen.box_mesh(5, 4, 3))
    _check_mesh(meshgen.slab_mesh(7, 3))
    w = meshgen.wedge_mesh(meshgen.graded(12, 0.0, 0.8, 15.0), np.concatenate([meshgen.graded(4, 1.5e-4, 3.6e-3), meshgen.graded(6, 3.6e-3, 0.15, 8.0)[1:]]),
                           half_angle_deg=5.0, inlet_split=[("jet", 3.6e-3), ("coflow", 1.0)])
    _check_mesh(w)
    assert [p.name for p in w.patches] == ["jet", "coflow", "outflow", "axis", "outerWall", "back", "front"]
    assert w.patch("jet").size == 4 and w.patch("coflow").size == 6
    assert w.geometric_d == (1, 1, -1) and w.solution_d == (1, 1, 1)
    # blockMesh grading: last/first cell width
    x = meshgen.graded(12, 0.0, 0.8, 15.0)
    assert (x[-1] - x[-2]) / (x[1] - x[0]) == pytest.approx(15.0, rel=1e-12)


def test_wedge_axisymmetry_detection(oracle_lib):
    """mcParticleCloud.C:661-717 + wedgePolyPatch::initTransforms"""
    w = meshgen.wedge_mesh(np.linspace(0, 0.1, 6), np.linspace(1e-3, 0.02, 5), half_angle_deg=5.0)
    h = capi.CloudHandle(oracle_lib, w, cases.default_params(1), capacity=10)
    mi = h.mesh_info()
    assert mi.isAxiSymmetric == 1
    assert mi.openingAngle == pytest.approx(np.deg2rad(10.0), rel=1e-12)
    np.testing.assert_allclose(np.abs(list(mi.axis)), [1, 0, 0], atol=1e-14)
    np.testing.assert_allclose(np.abs(list(mi.centrePlaneNormal)), [0, 0, 1], atol=1e-14)
    # the box is not axisymmetric
    assert capi.CloudHandle(oracle_lib, meshgen.box_mesh(2, 2, 2), cases.default_params(1), capacity=10).mesh_info().isAxiSymmetric == 0


def _jit_cases():
    gold = os.path.join(ROOT, "tests", "golden")
    yield "synthetic box (bench workload)", cases.synthetic_box(4, 2, 2, ppc=64, closed=False, number_control=True), False, True
    tab = cases.load_flamelet_tables(os.path.join(gold, "flamelet_sandiaD.npz"))
    yield "flamelet, 3 scalars", cases.with_scalars(cases.synthetic_box(4, 2, 2, ppc=8, closed=True), ["z", "chi", "T"],
                                                    "SteadyFlamelet", tab), False, False
    yield "BurkeSchumann wedge", cases.with_scalars(cases.wedge_jet(nx=4, nr_jet=2, nr_co=2, ppc=8), ["z", "T"], "BurkeSchumann"), True, True


@pytest.mark.parametrize("compiler", [1, 2])
@pytest.mark.parametrize("idx", [0, 1, 2])
def test_specialised_kernel_compiles_for_sm100a(idx, compiler):
    """The translation unit the library builds at run time (jit.cu) for three model selections, through the
    library's own compile path — kernel sources embedded in the .so, NVRTC in memory (compiler 1) and the
    nvcc route used when NVRTC is missing (compiler 2): both must produce a cubin for sm_100a that holds
    k_evolve_jit. (Running it needs a GPU: tests/test_gpu_jit.py.)"""
    name, case, axisym, has_open = list(_jit_cases())[idx]
    lib = C.CDLL(capi.device_library_path())
    holder = capi.Library.__new__(capi.Library)
    holder.lib, holder.prefix = lib, "pdfb200_"
    holder._declare()
    gd = case.mesh.geometric_d
    src = capi.jit_source(holder, case.params, axisym, has_open, 0, gd, case.mesh.solution_d)
    assert "struct JitP" in src and f"JIT_AXISYM {int(axisym)}" in src
    n, log = capi.jit_compile(holder, case.params, axisym, has_open, 0, gd, case.mesh.solution_d, compiler=compiler)
    assert n > 10000, f"{name}: {log[-3000:]}"


def test_jit_cache_directory_must_be_private(tmp_path, monkeypatch):
    """a cache directory that others can write to is not used (the cubin found there would run on the
    caller's data); the library then compiles in memory without caching. Checked through the exported symbol by
    pointing PDFB200_JIT_CACHE at a world-writable directory and at a symlink."""
    import stat
    lib = C.CDLL(capi.device_library_path())
    if not hasattr(lib, "pdfb200_jit_cache_dir"):
        pytest.skip("no cache introspection")
    lib.pdfb200_jit_cache_dir.restype = C.c_long
    lib.pdfb200_jit_cache_dir.argtypes = [C.c_char_p, C.c_long]
    buf = C.create_string_buffer(4096)

    def cache_dir():
        n = lib.pdfb200_jit_cache_dir(buf, len(buf))
        return buf.value.decode() if n > 0 else ""

    good = tmp_path / "good"
    monkeypatch.setenv("PDFB200_JIT_CACHE", str(good))
    assert cache_dir() == str(good) and stat.S_IMODE(os.stat(good).st_mode) == 0o700
    bad = tmp_path / "bad"; bad.mkdir(); os.chmod(bad, 0o777)
    monkeypatch.setenv("PDFB200_JIT_CACHE", str(bad))
    assert cache_dir() == ""
    link = tmp_path / "link"; os.symlink(good, link)
    monkeypatch.setenv("PDFB200_JIT_CACHE", str(link))
    assert cache_dir() == ""
