"""Restart from the moment fields of a previous run (pdfb200_upload_moments; mcParticleCloud.C:339-517 reads them
READ_IF_PRESENT, checkMoments :739-820 accepts only a complete set) and the velocity-scalar flux moment
(new, BASELINE north_star "scalar fluxes"). The oracle half runs on the CPU, the device half needs a GPU."""
import numpy as np
import pytest

from pdffoam_b200 import capi, cases

from parity_util import compare_cell_fields, compare_particles, sort_by_id


def _case():
    case = cases.synthetic_box(6, 4, 3, ppc=16, closed=False, number_control=True, n_phi=2)
    case.params.deltaT0 = 2e-3
    case.params.averagingCoeff = 20.0
    return case


def _restart_roundtrip(lib, device=0):
    case = _case()
    a = case.make_cloud(lib, device=device)
    a.seed_particles()
    a.evolve(5)
    P, cf, dt, step, next_id = a.download_particles(), a.download_cell_fields(), a.delta_t(), a.step_counter(), a.id_counter()
    assert step == 5
    # the restarted cloud: same FV fields, the means the first run left (rho, Phi, Cov are cloud fields), its moments
    case_b = _case()
    own = case_b.mesh.owner[case_b.mesh.n_internal_faces:]
    nb = case_b.mesh.n_boundary_faces
    # (boundary values that are not fixed are re-derived from the cells; the fixed ones are the case's own)
    case_b.cf = dict(case_b.cf, rho=cf["rho"], Phi=cf["Phi"], Cov=cf["Cov"], Covb=[c[own].copy() for c in cf["Cov"]],
                     CovbFixed=[np.zeros(nb, np.uint8) for _ in cf["Cov"]])
    b = capi.CloudHandle(lib, case_b.mesh, case_b.params, device=device, capacity=int(P["m"].size * 1.5) + 4096)
    b.upload_fv_fields(case_b.fv); b.upload_cloud_fields(case_b.cf)
    b.upload_moments(cf)
    b.upload_particles(P)
    b.set_delta_t(dt); b.set_step_counter(step); b.set_id_counter(next_id)
    cb = b.download_cell_fields()
    for k in ("rho", "pnd", "U", "Tau", "k", "mMom", "VMom", "UMom", "UUMom"):
        np.testing.assert_allclose(cb[k], cf[k], rtol=1e-13, atol=1e-300, err_msg=f"derived from uploaded moments: {k}")
    for name in ("Phi", "Cov", "UPhi", "PhiMom", "PhiPhiMom", "UPhiMom"):
        for x, y in zip(cb[name], cf[name]):
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-14, err_msg=name)
    ra, rb = a.evolve(3), b.evolve(3)
    for x, y in zip(ra, rb):
        assert (x.nParticles, x.nInjected, x.nDeleted, x.nCloned, x.nEliminated) == (y.nParticles, y.nInjected, y.nDeleted, y.nCloned, y.nEliminated)
        assert x.deltaT == pytest.approx(y.deltaT, rel=1e-12)
    compare_particles(a.download_particles(), b.download_particles(), 1e-11, 1e-13, True, "restarted run")
    compare_cell_fields(a.download_cell_fields(), b.download_cell_fields(), 1e-11, 1e-13, "restarted run")
    # an incomplete set is refused, as checkMoments() does ("Moments are missing")
    part = dict(cf); part["UUMom"] = None
    with pytest.raises(capi.PdfError, match="Moments are missing"):
        b.upload_moments(part)


def _flux_definition(lib, device=0):
    """averagingCoeff = 1: the flux field of a step is the weighted covariance of the particles in each cell"""
    case = cases.synthetic_box(4, 3, 2, ppc=20, closed=True, number_control=False, lts="none", n_phi=2)
    case.params.averagingCoeff = 1.0
    case.params.deltaT0 = 1e-3
    h = case.make_cloud(lib, device=device)
    P = case.random_particles(20, seed=3)
    rng = np.random.default_rng(0)
    P["Phi"][1] = rng.uniform(0, 1, P["m"].size)
    h.upload_particles(P)
    h.evolve(1)
    Q = sort_by_id(h.download_particles())
    cf = h.download_cell_fields()
    w = Q["eta"] * Q["m"]
    for c in range(case.mesh.n_cells):
        sel = Q["cell"] == c
        if sel.sum() == 0:
            continue
        W = w[sel].sum()
        Ubar = (w[sel][:, None] * Q["U"][sel]).sum(axis=0) / W
        for k in range(2):
            pbar = (w[sel] * Q["Phi"][k][sel]).sum() / W
            flux = (w[sel][:, None] * Q["U"][sel] * Q["Phi"][k][sel][:, None]).sum(axis=0) / W - Ubar * pbar
            np.testing.assert_allclose(cf["UPhi"][k][c], flux, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(cf["UPhiMom"][k][c], (w[sel][:, None] * Q["U"][sel] * Q["Phi"][k][sel][:, None]).sum(axis=0), rtol=1e-12)


def test_oracle_restart_from_uploaded_moments(oracle_lib):
    _restart_roundtrip(oracle_lib)


def test_oracle_flux_moment_definition(oracle_lib):
    _flux_definition(oracle_lib)


@pytest.mark.gpu
def test_device_restart_from_uploaded_moments(device_lib):
    _restart_roundtrip(device_lib)


@pytest.mark.gpu
def test_device_flux_moment_definition(device_lib):
    _flux_definition(device_lib)
