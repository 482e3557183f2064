"""The four tutorial-shaped configurations of BASELINE.json configs[0..3] (pdffoam_b200.cases.tutorial_case,
SURVEY Appendix E) on the device against the oracle: axisymmetric wedge with jet / pilot / coflow (or bluff body)
blocks, graded cells, the reference's flamelet libraries, modified-Curl mixing, BurkeSchumann."""
import pytest

from pdffoam_b200 import cases

from parity_util import run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,ppc", [("SydneyBluffBodyFlow", 12), ("SandiaD", 16), ("SydneyBluffBodyFlame", 12), ("SandiaPropaneJet4x", 4)])
def test_tutorial_shape_matches_oracle(device_lib, oracle_lib, name, ppc):
    case = cases.tutorial_case(name, ppc=ppc)
    case.params.deltaT0 = 2e-6
    dev, ora, reps = run_both(case, device_lib, oracle_lib, None, n_steps=4, seed_on_device=True, rtol=1e-10, atol=1e-11)
    assert sum(r[0].nInjected for r in reps) > 0
    assert sum(r[0].nLost for r in reps) == 0
    assert dev.mesh_info().isAxiSymmetric == 1
