"""Further parity cases of the CUDA path against the CPU restatement: warped
(non-planar) faces, the largest scalar count, a crowded cell (every per-cell
kernel leaves its register fast path) and the `cell` interpolation scheme."""
import numpy as np
import pytest

from pdffoam_b200 import cases

from parity_util import run_both

pytestmark = pytest.mark.gpu


def _perturb_interior_points(mesh, amplitude, seed=3):
    """Move every point that is not on a boundary face by up to `amplitude` of the
    smallest cell edge: the quad faces are no longer planar, which is what the
    face-centre tet decomposition is for (tetFacePointCellDecomposition.C:31-120)."""
    on_bnd = np.zeros(mesh.n_points, bool)
    for f in range(mesh.n_internal_faces, mesh.n_faces):
        on_bnd[mesh.face_points[mesh.face_start[f]:mesh.face_start[f + 1]]] = True
    nx, ny, nz = mesh.shape
    ext = mesh.points.max(axis=0) - mesh.points.min(axis=0)
    h = float(np.min(ext / np.array([nx, ny, nz])))
    rng = np.random.default_rng(seed)
    d = rng.uniform(-amplitude * h, amplitude * h, size=mesh.points.shape)
    mesh.points[~on_bnd] += d[~on_bnd]


def test_warped_hex_mesh_deterministic_and_stochastic(device_lib, oracle_lib):
    case = cases.synthetic_box(7, 5, 4, ppc=20, variant="deterministic", closed=True)
    _perturb_interior_points(case.mesh, 0.15)
    dev, ora, reps = run_both(case, device_lib, oracle_lib, None, n_steps=10, seed_on_device=True, rtol=1e-10)
    assert sum(r[0].nLost for r in reps) == 0
    case = cases.synthetic_box(7, 5, 4, ppc=20, closed=False, number_control=True)
    _perturb_interior_points(case.mesh, 0.15)
    case.params.deltaT0 = 2e-3
    dev, ora, reps = run_both(case, device_lib, oracle_lib, None, n_steps=6, seed_on_device=True, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nInjected for r in reps) > 0 and sum(r[0].nLost for r in reps) == 0


@pytest.mark.parametrize("n_phi", [2, 4, 5, 6])
def test_many_scalars(device_lib, oracle_lib, n_phi):
    """every instantiation of the moments kernel beyond the common ones (one group of accumulators up to nPhi = 4, two
    groups from 5 on); nPhi = PDFB200_MAX_PHI = 6 is the widest particle kernel, 21 covariances"""
    case = cases.synthetic_box(6, 4, 3, ppc=16, closed=False, n_phi=n_phi)
    case.params.deltaT0 = 2e-3
    P = case.random_particles(16, seed=12)
    run_both(case, device_lib, oracle_lib, P, n_steps=5, rtol=1e-10, atol=1e-12)


def test_crowded_cell(device_lib, oracle_lib):
    """One cell holds > 256 local particles: the counting sort's ordering kernel, number
    control and the moments all take their large-cell paths."""
    case = cases.synthetic_box(4, 3, 2, ppc=40, closed=True, number_control=True, lts="none")
    case.params.deltaT0 = 2e-3
    P = case.random_particles(40, seed=6)
    rng = np.random.default_rng(1)
    idx = np.nonzero(P["cell"] == 5)[0]
    dup = np.tile(idx, 8)                                   # 40 + 320 particles in cell 5
    Q = {k: (np.concatenate([v, v[dup]]) if k != "Phi" else [np.concatenate([a, a[dup]]) for a in v]) for k, v in P.items()}
    Q["id"] = np.arange(Q["m"].shape[0], dtype=np.uint64)
    lo, hi = case.info["cell_lo"][5], case.info["cell_hi"][5]
    Q["pos"] = Q["pos"].copy()
    Q["pos"][-dup.size:] = lo + rng.uniform(0.05, 0.95, size=(dup.size, 3)) * (hi - lo)
    Q["m"] = Q["m"] * rng.uniform(0.5, 1.5, size=Q["m"].shape[0])
    dev, ora, reps = run_both(case, device_lib, oracle_lib, Q, n_steps=4, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nEliminated for r in reps) > 0


def test_cell_interpolation_scheme_everywhere(device_lib, oracle_lib):
    """interpolationSchemes default `cell` (mcSolution.C:49) for every field."""
    case = cases.synthetic_box(6, 4, 4, ppc=18, closed=False)
    for name in ("interpU", "interpk", "interpDiffU", "interpOmega", "interpUPosCorr", "interpZVar", "interpEta"):
        setattr(case.params, name, 0)
    case.params.deltaT0 = 2e-3
    P = case.random_particles(18, seed=4)
    run_both(case, device_lib, oracle_lib, P, n_steps=5, rtol=1e-10, atol=1e-12)
