"""The CPU oracle pinned against the reference's own code: oracle/_ref/libpdfref.so holds
mcSteadyFlamelet.C:36-95 (findCoeffs, binterp), mcInletRandom + mcInversionInletRandom,
richTetrahedron and tetFacePointCellDecomposition compiled UNMODIFIED from /root/reference
against a stand-in for the OpenFOAM headers they include (oracle/ref/). Every comparison is
bit for bit unless stated. Skipped when the library has not been built (it is built here by
__graft_entry__.build(); the reference tree does not exist on the GPU box)."""
import ctypes as C
import os

import numpy as np
import pytest

from pdffoam_b200 import cases, meshgen
from pdffoam_b200.capi import CloudHandle
from oracle.oracle import RefLib, dp, ip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    r = RefLib()
    if r is None:
        pytest.skip("oracle/_ref/libpdfref.so not built (needs /root/reference)")
    return r


# ---------------------------------------------------------------------------
# mcSteadyFlamelet.C:44-93
# ---------------------------------------------------------------------------
def _find_coeffs(lib_fn, lst, v):
    i, w = C.c_int32(), C.c_double()
    rc = lib_fn(lst.size, dp(lst), float(v), C.byref(i), C.byref(w))
    return rc, i.value, w.value


@pytest.mark.parametrize("name", ["sandiaD", "bluffbody"])
def test_flamelet_lookup_matches_reference_code(oracle_lib, ref, name):
    """findCoeffs + binterp on the reference's own flamelet libraries (tests/golden/flamelet_*.npz): indices,
    weights and interpolated rho / T identical to the compiled reference for 20 000 (z, chi) queries, including
    values outside the table on both sides and exactly on its nodes."""
    t = np.load(os.path.join(ROOT, "tests", "golden", f"flamelet_{name}.npz"))
    z, chi = np.ascontiguousarray(t["z"]), np.ascontiguousarray(t["chi"])
    rng = np.random.default_rng(7)
    zq = np.concatenate([rng.uniform(-0.05, 1.05, 9000), z[rng.integers(0, z.size, 1000)]])
    cq = np.concatenate([np.exp(rng.uniform(np.log(1e-6), np.log(max(chi.max(), 1.0) * 5), 9000)), chi[rng.integers(0, chi.size, 1000)]])
    fo, fr = oracle_lib.fn("find_coeffs"), ref.lib.pdfref_find_coeffs
    for tabname in ("rho", "T"):
        tab = np.ascontiguousarray(t[tabname], np.float64)
        for k in range(0, zq.size, 1):
            rc1, iz1, wz1 = _find_coeffs(fo, z, zq[k]); rc2, iz2, wz2 = _find_coeffs(fr, z, zq[k])
            assert (rc1, iz1, wz1) == (rc2, iz2, wz2), f"z = {zq[k]!r}"
            if chi.size > 1:
                rc1, ic1, wc1 = _find_coeffs(fo, chi, cq[k]); rc2, ic2, wc2 = _find_coeffs(fr, chi, cq[k])
                assert (rc1, ic1, wc1) == (rc2, ic2, wc2), f"chi = {cq[k]!r}"
            else:
                ic1, wc1 = 0, 0.0      # single-row library: the row weight is 0 and the reference reads row ix+1 with weight 0
                continue
            a = oracle_lib.fn("binterp")(dp(tab), tab.shape[0], tab.shape[1], ic1, wc1, iz1, wz1)
            b = ref.lib.pdfref_binterp(dp(tab), tab.shape[0], tab.shape[1], ic1, wc1, iz1, wz1)
            assert a == b, f"{tabname}: z = {zq[k]!r}, chi = {cq[k]!r}: {a!r} vs {b!r}"
        if tabname == "rho":
            zq, cq = zq[:2000], cq[:2000]          # the second table repeats the index search: fewer queries


def test_find_coeffs_rejects_bad_weight_like_the_reference(oracle_lib, ref):
    """a list that is not increasing makes the weight leave [0, 1]: FatalError in the reference (.C:68-73)"""
    lst = np.array([0.0, 1.0, 0.5, 2.0])
    for v in (0.75, 1.5):
        rc_o, *_ = _find_coeffs(oracle_lib.fn("find_coeffs"), lst, v)
        rc_r, *_ = _find_coeffs(ref.lib.pdfref_find_coeffs, lst, v)
        assert (rc_o != 0) == (rc_r != 0)


# ---------------------------------------------------------------------------
# mcInletRandom / mcInversionInletRandom
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("UMean,uRms", [(2.0, 4.0), (50.0, 5.0), (0.3, 1.0), (-1.0, 2.0), (10.0, 0.05)])
def test_inlet_random_matches_reference_code(oracle_lib, ref, UMean, uRms):
    """Q, PDF, CDF and value(u) of the "inversion" generator. (2, 4) is the case of the reference's own
    tests/mcInletRandom/mcInletRandomTest.C. exp/erf/sqrt come from the same libm on both sides, so the
    results are identical; value() is Newton's iteration on them and is compared the same way."""
    rng = np.random.default_rng(11)
    u = np.ascontiguousarray(np.concatenate([rng.uniform(0, 1, 4000), [1e-12, 1 - 1e-12, 0.5]]))
    vals = np.zeros(u.size); out3 = np.zeros(3); warn = C.c_int32()
    for x in (0.1 * uRms, UMean + uRms, max(UMean, 0) + 3 * uRms):
        assert ref.lib.pdfref_inlet_random(UMean, uRms, x, dp(out3), u.size, dp(u), dp(vals), C.byref(warn)) == 0, ref.error()
        o = np.zeros(5)
        oracle_lib.fn("inlet_random")(UMean, uRms, x, 0.5, dp(o))
        assert o[0] == out3[0], "Q"
        assert o[2] == out3[1], "PDF"
        assert o[3] == out3[2], "CDF"
    mine = np.zeros(u.size)
    for k in range(u.size):
        o = np.zeros(5)
        oracle_lib.fn("inlet_random")(UMean, uRms, 1.0, u[k], dp(o))
        mine[k] = o[4]
    bad = np.nonzero(mine != vals)[0]
    assert bad.size == 0, f"value(u) differs for u = {u[bad[:5]]}: {mine[bad[:5]]} vs {vals[bad[:5]]}"
    assert np.all(vals > 0)                      # only inflowing velocities


def test_inlet_random_unknown_type_fails_like_the_reference(ref):
    assert ref.lib.pdfref_inlet_random_select(b"inversion") == 0
    assert ref.lib.pdfref_inlet_random_select(b"noSuchGenerator") != 0
    assert "Unknown mcInletRandom type" in ref.error()


# ---------------------------------------------------------------------------
# tetFacePointCellDecomposition<richTetPointRef>
# ---------------------------------------------------------------------------
def _meshes():
    yield "box", meshgen.box_mesh(4, 3, 2)
    m = meshgen.box_mesh(5, 4, 3)
    on_bnd = np.zeros(m.n_points, bool)
    for f in range(m.n_internal_faces, m.n_faces):
        on_bnd[m.face_points[m.face_start[f]:m.face_start[f + 1]]] = True
    d = np.random.default_rng(3).uniform(-0.04, 0.04, size=m.points.shape)
    m.points[~on_bnd] += d[~on_bnd]                # warped (non-planar) faces
    yield "warped box", m
    yield "wedge", meshgen.wedge_mesh(meshgen.graded(6, 0.0, 0.1, 3.0), meshgen.graded(5, 2e-4, 0.02, 4.0), 5.0)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_tet_decomposition_matches_reference_code(oracle_lib, ref, idx):
    """Order of the tets (cell, face, point), face label and point pair of each, richTetrahedron::update's
    Sa..Sd / mag / gradNi, cellTetrahedra, and inside() / find() for points in, on and outside the tets."""
    name, mesh = list(_meshes())[idx]
    h = CloudHandle(oracle_lib, mesh, cases.default_params(1), capacity=10)
    g = h.download_geometry()
    fc, cc = np.ascontiguousarray(g["faceCentres"]), np.ascontiguousarray(g["cellCentres"])
    pts = np.ascontiguousarray(mesh.points, np.float64)
    fs, fp = np.ascontiguousarray(mesh.face_start, np.int32), np.ascontiguousarray(mesh.face_points, np.int32)
    own, nei = np.ascontiguousarray(mesh.owner, np.int32), np.ascontiguousarray(mesh.neighbour, np.int32)
    d = ref.lib.pdfref_decomp_create(mesh.n_points, dp(pts), mesh.n_faces, ip(fs), ip(fp), ip(own), mesh.n_internal_faces, ip(nei),
                                     mesh.n_cells, dp(fc), dp(cc))
    assert d, ref.error()
    try:
        n_tets = ref.lib.pdfref_decomp_num_tets(d)
        assert n_tets == h.mesh_info().nTets
        to = h.download_tets()
        S_r, g_r, v_r = np.zeros(12), np.zeros(12), np.zeros(12)
        S_o, g_o, v_o = np.zeros(12), np.zeros(12), np.zeros(12)
        mag_r, mag_o = np.zeros(1), np.zeros(1)
        face_r, pair_r, pair_o = C.c_int32(), np.zeros(2, np.int32), np.zeros(2, np.int32)
        for t in range(n_tets):
            assert ref.lib.pdfref_decomp_tet(d, t, dp(S_r), dp(mag_r), dp(g_r), dp(v_r), C.byref(face_r), ip(pair_r)) == 0
            oracle_lib.check(oracle_lib.fn("tet_detail")(h.h, t, dp(S_o), dp(mag_o), dp(g_o), dp(v_o), ip(pair_o)), "tet_detail")
            assert face_r.value == to["face"][t], f"{name}: face of tet {t}"
            assert np.array_equal(pair_r, pair_o), f"{name}: point pair of tet {t}"
            assert np.array_equal(v_r, v_o), f"{name}: vertices of tet {t}"
            assert np.array_equal(S_r, S_o), f"{name}: face area vectors of tet {t}"
            assert mag_r[0] == mag_o[0], f"{name}: volume of tet {t}"
            assert np.array_equal(g_r, g_o), f"{name}: gradNi of tet {t}"
            assert np.array_equal(g_r[:9], to["gradN"][t].ravel()), f"{name}: downloaded gradN of tet {t}"
        # cellTetrahedra: consecutive labels per cell
        start = np.zeros(mesh.n_cells + 1, np.int32); lab = np.zeros(n_tets, np.int32)
        assert ref.lib.pdfref_decomp_cell_tets(d, ip(start), ip(lab)) == 0
        assert np.array_equal(lab, np.arange(n_tets, dtype=np.int32))
        assert np.array_equal(np.repeat(np.arange(mesh.n_cells), np.diff(start)), to["cell"])
        # inside() / find(): random points in the bounding box (some outside the mesh), cell centres, points, face centres
        rng = np.random.default_rng(5)
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        q = np.concatenate([lo + rng.uniform(-0.05, 1.05, size=(3000, 3)) * (hi - lo), cc, pts, fc])
        q = np.ascontiguousarray(q)
        hint = np.ascontiguousarray(rng.integers(-1, mesh.n_cells, size=q.shape[0]).astype(np.int32))
        out_r, out_o = np.zeros(q.shape[0], np.int32), np.zeros(q.shape[0], np.int32)
        assert ref.lib.pdfref_decomp_find(d, q.shape[0], dp(q), ip(hint), ip(out_r)) == 0
        oracle_lib.check(oracle_lib.fn("find_reference")(h.h, q.shape[0], dp(q), ip(hint), ip(out_o)), "find_reference")
        assert np.array_equal(out_r, out_o), f"{name}: find() differs for {np.count_nonzero(out_r != out_o)} points"
        assert (out_r >= 0).sum() > 1000 and (out_r < 0).sum() > 0
        for t in rng.integers(0, n_tets, 200):
            for k in rng.integers(0, q.shape[0], 20):
                p = np.ascontiguousarray(q[k])
                assert ref.lib.pdfref_decomp_inside(d, int(t), dp(p)) == oracle_lib.fn("tet_inside")(h.h, int(t), dp(p))
    finally:
        ref.lib.pdfref_decomp_destroy(d)


# ---------------------------------------------------------------------------
# mcSamplingInletRandom ("sampling"): stateful in the reference (batches of 1e6 deviates on the cloud's Random),
# stateless and keyed here - only the distribution can agree
# ---------------------------------------------------------------------------
def _ks_two_sample(a, b):
    a, b = np.sort(a), np.sort(b)
    allv = np.concatenate([a, b])
    d = np.abs(np.searchsorted(a, allv, side="right") / a.size - np.searchsorted(b, allv, side="right") / b.size).max()
    return d, d * np.sqrt(a.size * b.size / (a.size + b.size))


@pytest.mark.parametrize("UMean,uRms", [(2.0, 4.0), (20.0, 3.0), (0.0, 1.0)])
def test_sampling_generator_has_the_reference_distribution(oracle_lib, ref, UMean, uRms):
    """(2, 4) is the case of tests/mcInletRandom/mcInletRandomTest.C, which histograms both generators. Two-sample
    Kolmogorov-Smirnov between 2e5 values of the compiled reference class and of the oracle's keyed and sequential
    generators; and each against the analytic CDF (mcInletRandomI.H:34-38)."""
    n = 200000
    r = np.zeros(n)
    assert ref.lib.pdfref_inlet_sampling(UMean, uRms, 1, n, dp(r)) == 0, ref.error()
    keyed, seq = np.zeros(n), np.zeros(n)
    oracle_lib.fn("inlet_sampling")(UMean, uRms, 12345, 7, 3, n, dp(keyed), dp(seq))
    assert r.min() > 0 and keyed.min() > 0 and seq.min() > 0
    for name, x in (("keyed", keyed), ("sequential", seq)):
        d, stat = _ks_two_sample(r, x)
        assert stat < 1.95, f"{name}: KS distance {d:.4g} (scaled {stat:.3g}) against the reference sample"   # 0.1 % level
    # against the analytic CDF
    o = np.zeros(5)
    xs = np.sort(keyed)[:: n // 2000]
    cdf = np.array([(oracle_lib.fn("inlet_random")(UMean, uRms, float(x), 0.5, dp(o)), o[3])[1] for x in xs])
    emp = np.searchsorted(np.sort(keyed), xs, side="right") / n
    assert np.abs(emp - cdf).max() * np.sqrt(n) < 1.95
