"""the KulkarniAutoIgnition reaction model (mcKulkarniAutoIgnition.C:339-380) on the reference's own
auto-ignition library (tests/golden/autoignition.npz <- tutorials/AutoIgnitionBurner/constant/autoIgnition) and the
"sampling" inflow generator (mcSamplingInletRandom.C:37-102)."""
import importlib.util
import os

import numpy as np
import pytest

from pdffoam_b200 import cases

from parity_util import run_both, sort_by_id

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
_mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mg)            # the independent numpy restatement of findCoeffs / binterp


def _tables():
    return cases.load_autoignition_tables(os.path.join(ROOT, "tests", "golden", "autoignition.npz"))


def _kulkarni_case(**kw):
    case = cases.synthetic_box(5, 4, 3, ppc=16, **kw)
    return cases.with_scalars(case, ["z", "c", "pv", "T"], "KulkarniAutoIgnition", _tables(), mixed=[0, 1])


def _check_against_numpy(lib, device=0):
    """one step with everything but the reaction model switched off: c, pv, rho and T follow
    mcKulkarniAutoIgnition.C:339-380 evaluated with numpy on the same library"""
    case = _kulkarni_case(closed=True, lts="none", poscorr="none")
    case.params.velocityModel = 0; case.params.mixingModel = 0; case.params.DNum = 0.0
    t = _tables()
    P = case.random_particles(16, seed=4)
    rng = np.random.default_rng(1)
    n = P["m"].size
    P["U"][:] = 0.0
    P["Phi"][0] = rng.uniform(-0.02, 1.02, n)
    P["Phi"][1] = rng.uniform(0.0, 0.25, n)              # some above cEq: limited
    h = case.make_cloud(lib, device=device)
    h.upload_particles(P)
    dt = 2e-5
    h.set_delta_t(dt)
    rep = h.evolve(1)[0]
    Q = sort_by_id(h.download_particles())
    z, c0 = P["Phi"][0], P["Phi"][1]
    cdot, rho, T = t["fields"]
    cdot_max = 0.0
    for i in range(n):
        iz, wz = _mg.find_coeffs(t["z"], z[i])
        cEq = t["cEq"][iz] * (1 - wz) + t["cEq"][iz + 1] * wz
        c = min(c0[i], cEq)
        pv = c / cEq if cEq > 1e-5 * t["cEq"].max() else 0.0
        ic, wc = _mg.find_coeffs(t["pv"], pv)
        cd = _mg.binterp(cdot, iz, wz, ic, wc)
        cdot_max = max(cdot_max, cd)
        assert Q["Phi"][2][i] == pytest.approx(pv, rel=1e-13, abs=1e-300)
        assert Q["Phi"][1][i] == pytest.approx(c + cd * 1.0 * dt, rel=1e-13, abs=1e-300)
        assert Q["rho"][i] == pytest.approx(_mg.binterp(rho, iz, wz, ic, wc), rel=1e-13)
        assert Q["Phi"][3][i] == pytest.approx(_mg.binterp(T, iz, wz, ic, wc), rel=1e-13)
    assert rep.CoMax >= cdot_max * (1 - 1e-12)            # p.Co() = max(p.Co(), cdot) (.C:379)


def test_oracle_kulkarni_matches_numpy_restatement(oracle_lib):
    _check_against_numpy(oracle_lib)


def test_kulkarni_without_library_is_refused(oracle_lib):
    from pdffoam_b200 import capi
    case = _kulkarni_case(closed=True)
    case.tables = None
    h = case.make_cloud(oracle_lib)
    h.upload_particles(case.random_particles(16, seed=1))
    with pytest.raises(capi.PdfError, match="auto-ignition tables"):
        h.evolve(1)


@pytest.mark.gpu
def test_device_kulkarni_matches_numpy_restatement(device_lib):
    _check_against_numpy(device_lib)


@pytest.mark.gpu
def test_device_kulkarni_matches_oracle(device_lib, oracle_lib):
    case = _kulkarni_case(closed=False, number_control=True)
    case.params.deltaT0 = 1e-4
    # three steps: the tabulated source is stiff (cdot up to 1.5e4 1/s on c ~ 1e-4, d cdot / d pv ~ 1e5), so a last-bit
    # difference in c grows by a factor of about ten per step - after the fifth step it shows at 1e-9
    run_both(case, device_lib, oracle_lib, None, n_steps=3, seed_on_device=True, rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_device_kulkarni_without_library_is_refused(device_lib):
    from pdffoam_b200 import capi
    case = _kulkarni_case(closed=True)
    case.tables = None
    h = case.make_cloud(device_lib)
    with pytest.raises(capi.PdfError, match="auto-ignition tables"):
        h.seed_particles()


@pytest.mark.gpu
def test_device_sampling_inflow_generator_matches_oracle(device_lib, oracle_lib):
    case = cases.synthetic_box(7, 4, 3, ppc=16, closed=False, number_control=True)
    case.params.inletRandom = cases.INLET_RANDOM["sampling"]
    case.params.deltaT0 = 2e-3
    dev, ora, reps = run_both(case, device_lib, oracle_lib, None, n_steps=6, seed_on_device=True, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nInjected for r in reps) > 50
