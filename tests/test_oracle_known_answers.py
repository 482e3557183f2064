"""Pins the CPU oracle: the reference ships no numeric goldens,
so the oracle is checked against (a) the reference's own flamelet tables with an
independent numpy restatement of the lookup (tests/golden/make_golden.py),
(b) the analytic inflow PDF/CDF used by the reference's tests/mcInletRandom and
(c) closed-form known answers derived from the cited reference lines
(SURVEY.md §8c)."""
import ctypes as C
import os

import numpy as np
import pytest

from pdffoam_b200 import cases, meshgen
from oracle.oracle import dp, ip

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------------------
# flamelet lookup (mcSteadyFlamelet.C:44-93,304-329)
# ---------------------------------------------------------------------------
def _find_coeffs(lib, lst, v):
    i = C.c_int32(); w = C.c_double()
    lst = np.ascontiguousarray(lst, np.float64)
    rc = lib.fn("find_coeffs")(lst.size, dp(lst), float(v), C.byref(i), C.byref(w))
    assert rc == 0
    return i.value, w.value


def test_find_coeffs_edge_cases(oracle_lib):
    lst = np.array([0.0, 0.25, 0.5, 1.0])
    assert _find_coeffs(oracle_lib, lst, -1.0) == (0, 0.0)        # below first
    assert _find_coeffs(oracle_lib, lst, 0.0) == (0, 0.0)         # exactly first node: lower_bound == begin
    assert _find_coeffs(oracle_lib, lst, 2.0) == (2, 1.0)         # above last
    assert _find_coeffs(oracle_lib, lst, 0.25) == (0, 1.0)        # exact interior node belongs to the left interval
    i, w = _find_coeffs(oracle_lib, lst, 0.75)
    assert i == 2 and w == pytest.approx(0.5, abs=1e-15)
    assert _find_coeffs(oracle_lib, lst, 1.0) == (2, 1.0)


def test_binterp_transposed_weights_2x2(oracle_lib):
    """binterp weights v10 (next z column) with the chi weight and v01 (next chi
    row) with the z weight (mcSteadyFlamelet.C:84-92) — kept verbatim."""
    v = np.array([[1.0, 2.0], [10.0, 20.0]])   # rows = chi, columns = z
    wx, wy = 0.3, 0.8                           # wx = chi weight, wy = z weight
    got = oracle_lib.fn("binterp")(dp(v), 2, 2, 0, wx, 0, wy)
    quirky = 1.0 * 0.7 * 0.2 + 2.0 * 0.3 * 0.2 + 10.0 * 0.7 * 0.8 + 20.0 * 0.3 * 0.8
    textbook = 1.0 * 0.7 * 0.2 + 10.0 * 0.3 * 0.2 + 2.0 * 0.7 * 0.8 + 20.0 * 0.3 * 0.8
    assert got == pytest.approx(quirky, rel=1e-15)
    assert abs(got - textbook) > 1.0


@pytest.mark.parametrize("name", ["flamelet_sandiaD", "flamelet_bluffbody"])
def test_lookup_on_reference_tables(oracle_lib, name):
    t = np.load(os.path.join(GOLD, name + ".npz"))
    g = np.load(os.path.join(GOLD, "lookup_golden.npz"))
    z, chi = g[name + "_z"], g[name + "_chi"]
    rho = np.ascontiguousarray(t["rho"]); T = np.ascontiguousarray(t["T"])
    for k in range(z.size):
        iz, wz = _find_coeffs(oracle_lib, t["z"], z[k])
        ic, wc = _find_coeffs(oracle_lib, t["chi"], chi[k])
        r = oracle_lib.fn("binterp")(dp(rho), rho.shape[0], rho.shape[1], ic, wc, iz, wz)
        tt = oracle_lib.fn("binterp")(dp(T), T.shape[0], T.shape[1], ic, wc, iz, wz)
        assert r == pytest.approx(g[name + "_rho"][k], rel=1e-13)
        assert tt == pytest.approx(g[name + "_T"][k], rel=1e-13)


# ---------------------------------------------------------------------------
# inflow velocity PDF (mcInletRandom.C:92-108, mcInletRandomI.H:28-39,
# mcInversionInletRandom.C:63-119); parameters of tests/mcInletRandom
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["test", "fast", "back"])
def test_inlet_pdf_cdf_q(oracle_lib, tag):
    g = np.load(os.path.join(GOLD, "inlet_golden.npz"))
    a, urms = g[tag + "_params"]
    out = np.zeros(5)
    for x, pdf, cdf in zip(g[tag + "_x"], g[tag + "_pdf"], g[tag + "_cdf"]):
        oracle_lib.fn("inlet_random")(a, urms, float(x), 0.5, dp(out))
        assert out[2] == pytest.approx(pdf, rel=1e-12, abs=1e-300)
        assert out[3] == pytest.approx(cdf, rel=1e-9, abs=1e-12)
    assert out[0] == pytest.approx(g[tag + "_Q"][0], rel=1e-10)     # Q = mean inflow speed


def test_inlet_inversion_sampling_matches_cdf(oracle_lib):
    """1e5 samples with Umean=2, urms=4 (the reference test draws 1e6 and eyeballs
    a histogram); here a Kolmogorov-Smirnov bound against the analytic CDF."""
    n = 100000
    s = np.zeros(n)
    oracle_lib.fn("inlet_random_sample")(2.0, 4.0, 12345, n, dp(s))
    assert s.min() > 0
    s.sort()
    out = np.zeros(5)
    cdf = np.empty(200)
    xs = s[:: n // 200][:200]
    for i, x in enumerate(xs):
        oracle_lib.fn("inlet_random")(2.0, 4.0, float(x), 0.5, dp(out)); cdf[i] = out[3]
    emp = np.searchsorted(s, xs, side="right") / n
    assert np.max(np.abs(emp - cdf)) < 1.63 / np.sqrt(n) * 1.5      # KS 1% critical value with margin
    # inversion is exact: value(CDF(x)) == x
    for x in (0.5, 2.0, 7.0, 15.0):
        oracle_lib.fn("inlet_random")(2.0, 4.0, x, 0.5, dp(out))
        u = out[3]
        oracle_lib.fn("inlet_random")(2.0, 4.0, x, u, dp(out))
        assert out[4] == pytest.approx(x, abs=1e-7)


# ---------------------------------------------------------------------------
# tet decomposition (tetFacePointCellDecomposition.C:31-151, richTetrahedronI.H:107-134,
# richTetrahedron.C:40-91)
# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def unit_cube(oracle_lib):
    mesh = meshgen.box_mesh(1, 1, 1, lo=(0, 0, 0), hi=(1, 1, 1), xmin=("xmin", meshgen.BC_SLIP), xmax=("xmax", meshgen.BC_SLIP))
    from pdffoam_b200.capi import CloudHandle
    return CloudHandle(oracle_lib, mesh, cases.default_params(1), capacity=10)


def _tet_detail(lib, h, t):
    S = np.zeros((4, 3)); g = np.zeros((4, 3)); v = np.zeros((4, 3)); mag = C.c_double(); pp = np.zeros(2, np.int32)
    lib.fn("tet_detail")(h.h, t, dp(S), C.byref(mag), dp(g), dp(v), ip(pp))
    return S, mag.value, g, v, pp


def test_unit_cube_24_tets(oracle_lib, unit_cube):
    h = unit_cube
    assert h.mesh_info().nTets == 24 and h.mesh_info().maxTetsPerCell == 24
    vol = 0.0
    for t in range(24):
        S, mag, g, v, pp = _tet_detail(oracle_lib, h, t)
        assert mag == pytest.approx(1.0 / 24.0, rel=1e-14)            # 24 congruent tets
        vol += mag
        np.testing.assert_allclose(S.sum(axis=0), 0, atol=1e-16)      # closed surface
        np.testing.assert_allclose(g.sum(axis=0), 0, atol=1e-13)      # partition of unity
        np.testing.assert_allclose(v[3], [0.5, 0.5, 0.5], atol=1e-15)  # d = cell centre
        # gradN_i is the gradient of the barycentric coordinate of vertex i
        for i in range(4):
            for j in range(4):
                val = 1.0 + g[i] @ (v[j] - v[i])
                assert val == pytest.approx(1.0 if i == j else 0.0, abs=1e-13)
        # the cell owns all its faces: tetPoints = (i, rcIndex(i))
        assert pp[1] == (pp[0] - 1) % 4
    assert vol == pytest.approx(1.0, rel=1e-14)


def test_tet_adjacency_is_symmetric(oracle_lib):
    case = cases.synthetic_box(3, 3, 2, ppc=2, closed=True)
    h = case.make_cloud(oracle_lib)
    T = h.download_tets()
    nbr, face, cell = T["nbr"], T["face"], T["cell"]
    entry = {0: 0, 1: 2, 2: 1, 3: 3}
    for t in range(nbr.shape[0]):
        for k in range(4):
            n = nbr[t, k]
            if n < 0:
                assert k == 3 and face[t] >= case.mesh.n_internal_faces
                continue
            assert nbr[n, entry[k]] == t
            if k == 3:
                assert cell[n] != cell[t] and face[n] == face[t]
            else:
                assert cell[n] == cell[t]
                if k != 0:
                    assert face[n] == face[t]


def test_inside_and_find_reference_semantics(oracle_lib, unit_cube):
    h = unit_cube
    rng = np.random.default_rng(1)
    pts = rng.uniform(0.01, 0.99, size=(500, 3))
    cell = np.zeros(500, np.int32)
    fr = np.zeros(500, np.int32); lo = np.zeros(500, np.int32)
    oracle_lib.fn("find_reference")(h.h, 500, dp(pts), ip(cell), ip(fr))
    oracle_lib.fn("locate")(h.h, 500, dp(pts), ip(cell), ip(lo))
    assert np.all(fr >= 0)
    assert np.array_equal(fr, lo)          # generic points: find() == most-inside tet
    for k in range(0, 500, 25):
        assert oracle_lib.fn("tet_inside")(h.h, int(fr[k]), dp(pts[k])) == 1
    # a point on a shared tet face is inside both tets (tolerance SMALL); find() returns the first
    centre = np.array([[0.5, 0.5, 0.5]])
    oracle_lib.fn("find_reference")(h.h, 1, dp(centre), ip(cell[:1].copy()), ip(fr[:1]))
    assert fr[0] == 0
    outside = np.array([1.5, 0.5, 0.5])
    assert all(oracle_lib.fn("tet_inside")(h.h, t, dp(outside)) == 0 for t in range(24))


def test_linear_field_is_reproduced_exactly(oracle_lib):
    """cellPointFace interpolation and gradInterpolationConstantTet are exact for a
    linear field (gradInterpolationConstantTet.C:67-139 is the gradient of the
    B.2 interpolant)."""
    case = cases.synthetic_box(4, 3, 3, ppc=2, closed=True)
    h = case.make_cloud(oracle_lib)
    m = case.mesh
    a = np.array([1.5, -2.0, 0.75]); c0 = 0.3
    cc = m.cell_centres_simple()
    fcb = m.face_centres_simple()[m.n_internal_faces:]
    cellv = np.ascontiguousarray(cc @ a + c0); bndv = np.ascontiguousarray(fcb @ a + c0)
    rng = np.random.default_rng(3)
    cell = rng.integers(0, m.n_cells, 300).astype(np.int32)
    # interior cells only: boundary points use inverse-distance weights of fewer cells
    nx, ny, nz = m.shape
    i, j, k = cell % nx, (cell // nx) % ny, cell // (nx * ny)
    interior = (i > 0) & (i < nx - 1) & (j > 0) & (j < ny - 1) & (k > 0) & (k < nz - 1)
    cell = np.ascontiguousarray(cell[interior])
    lo, hi = case.info["cell_lo"][cell], case.info["cell_hi"][cell]
    pts = np.ascontiguousarray(lo + rng.uniform(0.05, 0.95, size=(cell.size, 3)) * (hi - lo))
    out = np.zeros(cell.size); grad = np.zeros((cell.size, 3))
    rc = oracle_lib.fn("interpolate_scalar")(h.h, dp(cellv), dp(bndv), cell.size, dp(pts), ip(cell), dp(out), dp(grad))
    assert rc == 0
    np.testing.assert_allclose(out, pts @ a + c0, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(grad, np.broadcast_to(a, grad.shape), rtol=1e-11, atol=1e-11)


def test_rotation_tensor(oracle_lib):
    n1 = np.array([0.0, 0.6, 0.8]); n2 = np.array([0.0, 0.0, 1.0])
    T = np.zeros(9)
    oracle_lib.fn("rotation_tensor")(dp(n1), dp(n2), dp(T))
    T = T.reshape(3, 3)
    np.testing.assert_allclose(T @ n1, n2, atol=1e-15)
    np.testing.assert_allclose(T @ T.T, np.eye(3), atol=1e-15)
    np.testing.assert_allclose(T @ np.array([1.0, 0, 0]), [1.0, 0, 0], atol=1e-15)   # axis is invariant


def test_philox_known_answer(oracle_lib):
    """Philox4x32-10 known-answer vectors of the Random123 distribution."""
    def run(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        oracle_lib.fn("philox")(c, k, o)
        return [hex(x) for x in o]
    assert run([0, 0, 0, 0], [0, 0]) == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]
