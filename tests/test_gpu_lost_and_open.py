"""GPU parity on the paths a healthy run never takes ("lost-particle path is dead
code as far as the tests know"): particles lost at an `empty` handler of a 3-D mesh and after
more than 1000 tet crossings (mcParticle.C:147-158, notifyLostParticle mcParticleCloud.C:2088-2092),
the re-spread of their mass over the survivors of the cell in the next step (:1363-1369), and the
open boundary with `reflecting off` / the plain `open` handler (mcOpenBoundary.C:159-205)."""
import numpy as np
import pytest

from pdffoam_b200 import cases, meshgen

from parity_util import run_both
from test_oracle_lost_and_open import lost_case, open_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("number_control", [False, True])
def test_lost_at_empty_patch_and_mass_respread(device_lib, oracle_lib, number_control):
    case = lost_case(number_control=number_control)
    P = case.random_particles(12, seed=3)
    P["U"][:, 2] = np.abs(P["U"][:, 2]) + 8.0
    m0 = P["m"].sum()
    # run_both compares the reports (nLost included), every particle column (the masses carry the re-spread
    # shares from the second step on) and the cell fields after every step
    dev, ora, reps = run_both(case, device_lib, oracle_lib, P, n_steps=5, rtol=1e-10, atol=1e-12)
    lost = [r[0].nLost for r in reps]
    assert sum(lost) > 0 and lost[0] > 0
    Q = dev.download_particles()
    assert Q["m"].sum() <= m0 * (1 + 1e-12)
    # the step after a loss starts from masses that include the shares: the device hands them out in the particle
    # kernel, the oracle at the end of the step (mcParticleCloud.C:1363-1369) — the downloads above agreed


def test_lost_after_1000_crossings(device_lib, oracle_lib):
    """A few particles fast enough to cross more than 1000 tets in one step (bouncing between the slip walls)."""
    case = cases.synthetic_box(6, 4, 3, ppc=10, variant="deterministic", closed=True)
    P = case.random_particles(10, seed=9)
    fast = np.arange(0, P["m"].size, 97)
    P["U"][fast] *= 400.0
    dev = case.make_cloud(device_lib, device=0); ora = case.make_cloud(oracle_lib)
    dev.upload_particles(P); ora.upload_particles(P)
    dev.set_delta_t(0.5); ora.set_delta_t(0.5)
    rd, ro = dev.evolve(1)[0], ora.evolve(1)[0]
    assert rd.nLost == ro.nLost and rd.nLost >= fast.size // 2
    assert rd.nParticles == ro.nParticles == P["m"].size - rd.nLost
    from parity_util import compare_particles, compare_cell_fields
    compare_particles(dev.download_particles(), ora.download_particles(), 1e-10, 1e-12, True, "after the loss")
    rd, ro = dev.evolve(1)[0], ora.evolve(1)[0]
    compare_particles(dev.download_particles(), ora.download_particles(), 1e-10, 1e-12, True, "step after the loss")
    compare_cell_fields(dev.download_cell_fields(), ora.download_cell_fields(), 1e-10, 1e-12, "step after the loss")


@pytest.mark.parametrize("handler", [meshgen.BC_INLET_OUTLET, meshgen.BC_OPEN])
def test_reflecting_off_outlet(device_lib, oracle_lib, handler):
    case = open_case(0, handler)
    P = case.random_particles(12, seed=5)
    dev, ora, reps = run_both(case, device_lib, oracle_lib, P, n_steps=6, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nDeleted for r in reps) > 0
    assert sum(r[0].massOutInst[0] for r in reps) < 0


def test_open_handler_reflecting_on(device_lib, oracle_lib):
    """the plain `open` handler with the default reflecting switch: no injection through that patch"""
    case = open_case(1, meshgen.BC_OPEN)
    P = case.random_particles(12, seed=6)
    dev, ora, reps = run_both(case, device_lib, oracle_lib, P, n_steps=6, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nInjected for r in reps) > 0      # the inlet still injects
