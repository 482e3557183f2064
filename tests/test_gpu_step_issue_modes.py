"""How evolve() issues its steps must not change what they compute: one step per call (a host synchronisation after
every step), many steps per call (reports fetched from the device ring once per batch of 64), and CUDA-graph replay of
the captured step (pdfb200_set_graph_mode) give bit-identical particles, cell fields and reports."""
import numpy as np
import pytest

from pdffoam_b200 import cases

from parity_util import sort_by_id

pytestmark = pytest.mark.gpu


def _cases():
    c = cases.synthetic_box(7, 4, 3, ppc=14, closed=False, number_control=True)
    c.params.deltaT0 = 2e-3
    yield "open box", c
    t = cases.tutorial_case("SydneyBluffBodyFlame", ppc=8)      # modified Curl, flamelet, wedge, bluff-body slip patch
    t.params.deltaT0 = 2e-6
    yield "bluff-body flame", t


def _run(device_lib, case, mode, n_steps, per_call):
    h = case.make_cloud(device_lib)
    h.set_graph_mode(mode)
    h.seed_particles()
    reps = []
    done = 0
    while done < n_steps:
        k = min(per_call, n_steps - done)
        reps += h.evolve(k)
        done += k
    return sort_by_id(h.download_particles()), h.download_cell_fields(), reps, h.kernel_launches()


@pytest.mark.parametrize("idx", [0, 1])
def test_graph_replay_and_batched_issue_are_bit_identical(device_lib, idx):
    name, case = list(_cases())[idx]
    n = 70                                                       # more than one ring of 64 reports
    ref = _run(device_lib, case, 0, n, 1)
    for mode, per_call in ((0, n), (1, n), (1, 7)):
        got = _run(device_lib, case, mode, n, per_call)
        for k in ("pos", "U", "m", "Omega", "rho", "eta", "cell", "tet", "id"):
            assert np.array_equal(got[0][k], ref[0][k]), f"{name}: mode {mode}, {per_call} steps per call: particle column {k}"
        for a, b in zip(got[0]["Phi"], ref[0]["Phi"]):
            assert np.array_equal(a, b)
        for k in ("rho", "U", "Tau", "k", "PaNIC", "mMom", "UUMom"):
            assert np.array_equal(got[1][k], ref[1][k]), f"{name}: mode {mode}: cell field {k}"
        for ra, rb in zip(got[2], ref[2]):
            for f in ("deltaT", "deltaTNext", "rhoRes", "UMax", "CoMax", "totalParticleMass", "nParticles", "nInjected", "nDeleted", "nCloned",
                      "nEliminated", "nLost"):
                assert getattr(ra, f) == getattr(rb, f), f"{name}: mode {mode}: report.{f}"
            assert list(ra.deltaMassInst) == list(rb.deltaMassInst) and list(ra.massInInst) == list(rb.massInInst)
        assert got[3] == ref[3], "the launch count must not depend on how the steps are issued"
