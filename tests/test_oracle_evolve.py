"""Closed-form checks of the oracle's evolve() against formulas derived from the
cited reference lines (SURVEY.md §8c list), plus invariants of the tet walk."""
import numpy as np
import pytest

from pdffoam_b200 import cases, meshgen
from pdffoam_b200.capi import CloudHandle
from oracle.oracle import dp, ip

from parity_util import sort_by_id


def _uniform_case(nx=4, ny=3, nz=3, U=(3.0, 1.0, -2.0), k=1.5, eps=4.5, **kw):
    """closed all-slip box with uniform FV fields"""
    case = cases.synthetic_box(nx, ny, nz, ppc=8, closed=True, **kw)
    nc, nb = case.mesh.n_cells, case.mesh.n_boundary_faces
    fv = case.fv
    fv["U"] = np.tile(np.array(U, float), (nc, 1)); fv["Ub"] = np.tile(np.array(U, float), (nb, 1))
    fv["k"] = np.full(nc, k); fv["kb"] = np.full(nb, k)
    fv["epsilon"] = np.full(nc, eps); fv["epsilonb"] = np.full(nb, eps)
    fv["p"] = np.zeros(nc); fv["pb"] = np.zeros(nb)
    R = np.zeros((nc, 6)); R[:, 0] = R[:, 3] = R[:, 5] = 2.0 / 3.0 * k
    fv["R"] = R; fv["Rb"] = R[case.mesh.owner[case.mesh.n_internal_faces:]].copy()
    return case


def _set(case, **kw):
    from pdffoam_b200.cases import default_params
    p = case.params
    base = {f[0]: getattr(p, f[0]) for f in p._fields_ if f[0] not in ("mixed", "conserved", "addedIdx")}
    base.pop("nPhi"); base.pop("nMixed"); base.pop("nConserved"); base.pop("nAdded")
    base.update(kw)
    case.params = default_params(p.nPhi, mixed=list(p.mixed[:p.nMixed]), conserved=list(p.conserved[:p.nConserved]), **base)
    return case


def test_iem_decay_factor(oracle_lib):
    """phi -= (1 - exp(-Cmix/2 Omega eta dt)) (phi - phi_cell)  (mcIEMMixingModel.C:68-83)"""
    case = _set(_uniform_case(), velocityModel="none", positionCorrection="none", localTimeStepping="none", DNum=0.0, Cmix=2.0)
    P = case.random_particles(8, seed=1)
    P["U"][:] = 0.0
    rng = np.random.default_rng(0)
    P["Phi"][0] = rng.uniform(0, 1, P["m"].size)
    h = case.make_cloud(oracle_lib)
    h.upload_particles(P)
    dt = 3e-2
    h.set_delta_t(dt)
    h.evolve(1)
    Q = sort_by_id(h.download_particles())
    Omega = 4.5 / 1.5
    np.testing.assert_allclose(Q["Omega"], Omega, rtol=1e-14)
    mean = case.cf["Phi"][0][P["cell"]]
    expect = P["Phi"][0] - (1.0 - np.exp(-0.5 * 2.0 * Omega * 1.0 * dt)) * (P["Phi"][0] - mean)
    np.testing.assert_allclose(Q["Phi"][0], expect, rtol=1e-13, atol=1e-15)
    assert np.array_equal(Q["cell"], P["cell"])
    np.testing.assert_array_equal(Q["pos"], P["pos"])          # U = 0: nothing moves


def test_slm_step_without_noise(oracle_lib):
    """mcSLMFullVelocityModel::correct with C0 = 0 and uniform fields (grad p* = 0,
    diffU = diffk = 0): U' = U + A (U - Ubar) dT (1 + A dT / 2), A = -(C1/2) Omega."""
    Ubar = np.array([3.0, 1.0, -2.0])
    case = _set(_uniform_case(U=tuple(Ubar)), C0=0.0, C1=1.0, positionCorrection="none", localTimeStepping="none",
                DNum=0.0, mixingModel="none")
    P = case.random_particles(8, seed=2)
    h = case.make_cloud(oracle_lib)
    h.upload_particles(P)
    dt = 1e-3
    h.set_delta_t(dt)
    rep = h.evolve(1)[0]
    Q = sort_by_id(h.download_particles())
    Omega = 3.0
    A = -(0.5 * 1.0 + 0.75 * 0.0) * Omega
    dU = (-(A * Ubar) + A * P["U"]) * dt
    expect = P["U"] + dU + 0.5 * A * dU * dt
    # slip walls may have reflected some particles: compare the unreflected ones
    same = np.all(np.sign(Q["U"] - Ubar) == np.sign(expect - Ubar), axis=1)
    assert same.mean() > 0.9
    np.testing.assert_allclose(Q["U"][same], expect[same], rtol=1e-12, atol=1e-13)
    assert rep.CoMax >= Omega                                    # Co = max(Co, Omega) (.C:184)
    assert rep.deltaTNext == pytest.approx(case.params.CFL / rep.CoMax, rel=1e-15)   # mcParticleCloud.C:1512


def test_moments_of_three_particles(oracle_lib):
    """updateCloudPDF (mcParticleCloud.C:933-1008) on a one-cell mesh with
    averagingCoeff = 1 (existWt = 0): the fields are the instantaneous moments."""
    mesh = meshgen.box_mesh(1, 1, 1, lo=(0, 0, 0), hi=(2, 1, 1), xmin=("xmin", meshgen.BC_SLIP), xmax=("xmax", meshgen.BC_SLIP))
    prm = cases.default_params(2, velocityModel="none", mixingModel="none", omegaModel="none", reactionModel="none",
                               positionCorrection="none", averagingCoeff=1.0, particlesPerCell=3)
    nb = mesh.n_boundary_faces
    fv = dict(U=np.zeros((1, 3)), Ub=np.zeros((nb, 3)), UbFixed=np.zeros(nb, np.uint8), p=np.zeros(1), pb=np.zeros(nb),
              k=np.ones(1), kb=np.ones(nb), kbFixed=np.zeros(nb, np.uint8), epsilon=np.ones(1), epsilonb=np.ones(nb),
              R=np.zeros((1, 6)), Rb=np.zeros((nb, 6)))
    cf = dict(rho=np.ones(1), rhob=np.ones(nb), rhobFixed=np.zeros(nb, np.uint8), Phi=[np.zeros(1)] * 2, Phib=[np.zeros(nb)] * 2,
              PhibFixed=[np.zeros(nb, np.uint8)] * 2)
    case = cases.Case("oneCell", mesh, prm, fv, cf)
    h = case.make_cloud(oracle_lib)
    m = np.array([0.5, 1.0, 2.0]); eta = np.array([1.0, 2.0, 0.5]); rho = np.array([1.0, 0.5, 4.0])
    U = np.array([[1.0, 0, 0], [0, 2.0, 0], [1.0, 1.0, 1.0]]) * 0.0   # keep them in place
    Uv = np.array([[1.0, 0.5, 0], [0, 2.0, -1.0], [1.0, 1.0, 1.0]])
    P = dict(pos=np.array([[0.3, 0.3, 0.3], [1.2, 0.6, 0.4], [1.7, 0.2, 0.8]]), U=U, m=m, Omega=np.ones(3), rho=rho, eta=eta,
             Phi=[np.array([0.2, 0.4, 1.0]), np.array([300.0, 1200.0, 600.0])], cell=np.zeros(3, np.int32))
    h.upload_particles(P)
    h.set_delta_t(1e-3)
    h.evolve(1)
    # velocityModel none keeps U: use non-zero velocities in a second cloud to check the velocity moments
    h2 = case.make_cloud(oracle_lib)
    P2 = dict(P); P2["U"] = Uv
    h2.upload_particles(P2)
    h2.set_delta_t(1e-9)   # tiny step: positions change by O(1e-9) only
    h2.evolve(1)
    for hh, Uc in ((h, U), (h2, Uv)):
        c = hh.download_cell_fields()
        # the default mcLocalTimeStepping::correct resets eta to 1 (mcLocalTimeStepping.C:101-104),
        # so the extraction weights eta*m are the masses
        w = 1.0 * m
        M, V = w.sum(), (w / rho).sum()
        assert c["PaNIC"][0] == 3
        assert c["mMom"][0] == pytest.approx(M, rel=1e-15)
        assert c["VMom"][0] == pytest.approx(V, rel=1e-15)
        assert c["rho"][0] == pytest.approx(M / V, rel=1e-15)
        assert c["pnd"][0] == pytest.approx(M / 2.0, rel=1e-15)          # cell volume 2
        np.testing.assert_allclose(c["U"][0], (w[:, None] * Uc).sum(0) / M, rtol=1e-14, atol=1e-16)
        z, T = P["Phi"]
        assert c["Phi"][0][0] == pytest.approx((w * z).sum() / M, rel=1e-14)
        assert c["Phi"][1][0] == pytest.approx((w * T).sum() / M, rel=1e-14)
        zm, Tm = (w * z).sum() / M, (w * T).sum() / M
        assert c["Cov"][0][0] == pytest.approx((w * z * z).sum() / M - zm * zm, rel=1e-12)     # zz
        assert c["Cov"][1][0] == pytest.approx((w * z * T).sum() / M - zm * Tm, rel=1e-12)     # zT
        assert c["Cov"][2][0] == pytest.approx((w * T * T).sum() / M - Tm * Tm, rel=1e-12)     # TT
        Um = (w[:, None] * Uc).sum(0) / M
        tau_xx = (w * Uc[:, 0] ** 2).sum() / M - Um[0] ** 2
        tau_xy = (w * Uc[:, 0] * Uc[:, 1]).sum() / M - Um[0] * Um[1]
        assert c["Tau"][0][0] == pytest.approx(tau_xx, rel=1e-12, abs=1e-15)
        assert c["Tau"][0][1] == pytest.approx(tau_xy, rel=1e-12, abs=1e-15)
        k = 0.5 * sum((w * Uc[:, d] ** 2).sum() / M - Um[d] ** 2 for d in range(3))
        assert c["k"][0] == pytest.approx(max(k, prm.kMin), rel=1e-12)


def test_slip_reflection_preserves_speed_and_walk_matches_find(oracle_lib):
    """Frozen velocities in a closed box: |U| is invariant under specular
    reflection (mcSlipBoundary.C:61-79); after every step the walk's tet equals the
    reference's brute-force find() (tetFacePointCellDecomposition.C:125-151) and
    the particle is inside its cell."""
    case = cases.synthetic_box(6, 4, 4, ppc=12, variant="deterministic", closed=True)
    P = case.random_particles(12, seed=4)
    h = case.make_cloud(oracle_lib)
    h.upload_particles(P)
    speed0 = np.linalg.norm(P["U"], axis=1)
    n = P["m"].size
    fr = np.zeros(n, np.int32)
    crossed = 0
    for s in range(15):
        r = h.evolve(1)[0]
        assert r.nLost == 0 and r.nParticles == n
        Q = sort_by_id(h.download_particles())
        np.testing.assert_allclose(np.linalg.norm(Q["U"], axis=1), speed0, rtol=1e-13)
        pts = np.ascontiguousarray(Q["pos"]); cell = np.ascontiguousarray(Q["cell"])
        oracle_lib.fn("find_reference")(h.h, n, dp(pts), ip(cell), ip(fr))
        assert np.array_equal(fr, Q["tet"]), f"step {s}: walk and find() disagree for {np.count_nonzero(fr != Q['tet'])} particles"
        lo, hi = case.info["cell_lo"][cell], case.info["cell_hi"][cell]
        assert np.all(pts >= lo - 1e-12) and np.all(pts <= hi + 1e-12)
        crossed += np.count_nonzero(cell != P["cell"])
    assert crossed > n       # particles did travel through many cells


def test_mass_conservation_closed_box(oracle_lib):
    """closed box, no number control: particle mass is conserved exactly and the
    step report's balance closes (mcParticleCloud.C:1238-1246,1440-1444)."""
    case = cases.synthetic_box(5, 4, 3, ppc=15, closed=True)
    P = case.random_particles(15, seed=6)
    h = case.make_cloud(oracle_lib)
    h.upload_particles(P)
    for _ in range(6):
        r = h.evolve(1)[0]
        assert abs(r.deltaMassInst[0]) < 1e-13 and r.massInInst[0] == 0 and r.massOutInst[0] == 0
    Q = h.download_particles()
    assert Q["m"].sum() == pytest.approx(P["m"].sum(), rel=1e-14)


def test_clone_conserves_mass_exactly_and_eliminate_in_expectation(oracle_lib):
    """particleNumberControl (mcParticleCloud.C:1014-1221)"""
    case = cases.synthetic_box(4, 3, 3, ppc=20, closed=True, number_control=True, lts="none", poscorr="none")
    _set(case, velocityModel="none", DNum=0.0)
    base = case.random_particles(20, seed=8)
    base["U"][:] = 0
    rng = np.random.default_rng(5)
    base["m"] = base["m"] * rng.uniform(0.5, 1.5, base["m"].size)
    # too few: keep 8 of 20 in every cell -> clone 8 per cell (n = min(N, Npc - N))
    keep = np.concatenate([np.arange(c * 20, c * 20 + 8) for c in range(case.mesh.n_cells)])
    few = {k: (v[keep] if k != "Phi" else [a[keep] for a in v]) for k, v in base.items()}
    h = case.make_cloud(oracle_lib)
    h.upload_particles(few)
    h.set_delta_t(1e-6)
    r = h.evolve(1)[0]
    assert r.nCloned == 8 * case.mesh.n_cells and r.nEliminated == 0
    Q = h.download_particles()
    assert Q["m"].sum() == pytest.approx(few["m"].sum(), rel=1e-15)
    for c in (0, 7, 20):
        assert np.count_nonzero(Q["cell"] == c) == 16
    # the 8 heaviest were split in two
    c0 = few["cell"] == 0
    heavy = np.sort(few["m"][c0])[::-1][:8]
    got = np.sort(Q["m"][Q["cell"] == 0])
    np.testing.assert_allclose(np.sort(np.repeat(heavy / 2, 2)), got, rtol=1e-15)

    # too many: 40 per cell -> eliminate; mass conserved in expectation over seeds
    many = {k: (np.concatenate([v, v]) if k != "Phi" else [np.concatenate([a, a]) for a in v]) for k, v in base.items()}
    many["id"] = np.arange(many["m"].size, dtype=np.uint64)
    ratios = []
    for seed in range(40):
        case.params.seed = 1000 + seed
        h = case.make_cloud(oracle_lib)
        h.upload_particles(many)
        h.set_delta_t(1e-6)
        r = h.evolve(1)[0]
        assert r.nEliminated > 0 and r.nCloned == 0
        Q = h.download_particles()
        assert abs(Q["m"].size - 20 * case.mesh.n_cells) < 6 * np.sqrt(20 * case.mesh.n_cells)
        ratios.append(Q["m"].sum() / many["m"].sum())
    ratios = np.array(ratios)
    assert abs(ratios.mean() - 1.0) < 4 * ratios.std() / np.sqrt(ratios.size) + 1e-12
    assert ratios.std() > 0        # it is a random scheme, not exact


def test_error_behaviour(oracle_lib, tmp_path):
    """bad selections fail with a message instead of aborting the process
    (FatalErrorIn(...) << exit(FatalError) in the reference)."""
    from pdffoam_b200.capi import PdfError
    with pytest.raises(KeyError, match="Unknown mixingModel type Curl"):
        cases.default_params(1, mixingModel="Curl")          # no Curl model exists in the reference
    mesh = meshgen.box_mesh(2, 2, 2)
    with pytest.raises(PdfError, match="cloneAt"):
        CloudHandle(oracle_lib, mesh, cases.default_params(1, cloneAt=1.5), capacity=10)   # mcSolution.C:146-152
    with pytest.raises(PdfError, match="eliminateAt"):
        CloudHandle(oracle_lib, mesh, cases.default_params(1, eliminateAt=0.9), capacity=10)
    h = CloudHandle(oracle_lib, mesh, cases.default_params(1), capacity=10)
    with pytest.raises(PdfError):
        h.evolve(1)                                            # no fields yet
    with pytest.raises(PdfError, match="monotonically"):
        h.set_flamelet_tables(np.array([1.0, 0.5]), np.array([0.0, 1.0]), [np.zeros((2, 2))])


def test_naive_flamelet_relation(oracle_lib):
    """mcNaiveFlameletModel.C:64-69 on the CPU restatement."""
    case = cases.with_scalars(cases.synthetic_box(4, 3, 3, ppc=12, closed=True), ["z", "T"], "naiveFlamelet")
    case.params.deltaT0 = 2e-3
    h = case.make_cloud(oracle_lib)
    h.upload_particles(case.random_particles(12, seed=2))
    h.evolve(3)
    Q = h.download_particles()
    z = Q["Phi"][0]
    np.testing.assert_allclose(Q["rho"], 1 - 3.2 * z * (1 - z), rtol=1e-15)
    np.testing.assert_allclose(Q["Phi"][1], 1e5 / Q["rho"], rtol=1e-15)


def test_oracle_ids_are_64_bits_wide(oracle_lib):
    """The particle id keys the random streams with both of its words (counter words 0 and 1 of Philox): ids that agree
    in the low word evolve differently; the id counter starts above the largest id handed in."""
    from pdffoam_b200 import cases
    case = cases.synthetic_box(4, 3, 2, ppc=12, closed=True, number_control=True)
    case.params.deltaT0 = 2e-3
    P = case.random_particles(12, seed=2)
    out = []
    for base in (0, 1 << 32, (1 << 40) + 17):
        Q = dict(P); Q["id"] = np.arange(P["m"].size, dtype=np.uint64) + np.uint64(base)
        h = case.make_cloud(oracle_lib)
        h.upload_particles(Q)
        assert h.id_counter() == base + P["m"].size
        h.evolve(2)
        R = h.download_particles()
        assert R["id"].dtype == np.uint64 and R["id"].min() >= base
        o = np.argsort(R["id"])
        out.append(R["U"][o][: P["m"].size // 2])
    assert not np.allclose(out[0], out[1]) and not np.allclose(out[1], out[2])
