"""GPU parity tests: the CUDA library (through the C ABI) against the CPU oracle
on the same seeded inputs.

Bars (BASELINE.json north_star): particle-to-cell/tet indices bit-exact on
deterministic runs; deterministic model outputs and cell moments within 1e-10
relative in fp64. Because the device and the oracle draw from the same
counter-based Philox streams (keyed by particle id), the stochastic runs are
compared at the same tolerance as well; the statistical test against the
oracle's sequential-stream mode lives in test_gpu_statistics.py.
"""
import numpy as np
import pytest

from pdffoam_b200 import cases

from parity_util import compare_cell_fields, compare_particles, run_both

pytestmark = pytest.mark.gpu

RTOL = 1e-10   # fp64 tolerance stated by BASELINE.json


def test_geometry_and_tets_match_oracle(device_lib, oracle_lib):
    case = cases.synthetic_box(6, 5, 4, ppc=4, closed=True)
    dev = case.make_cloud(device_lib); ora = case.make_cloud(oracle_lib)
    gd, go = dev.download_geometry(), ora.download_geometry()
    for k in gd:
        if k == "hNum":   # cbrt differs in the last ulp between libm and CUDA
            np.testing.assert_allclose(gd[k], go[k], rtol=1e-15)
        else:
            assert np.array_equal(gd[k], go[k]), f"geometry {k} not bit-identical"
    td, to = dev.download_tets(), ora.download_tets()
    assert dev.mesh_info().nTets == ora.mesh_info().nTets == case.mesh.n_cells * 24
    for k in ("nbr", "face", "ptB", "ptC", "cell"):
        assert np.array_equal(td[k], to[k]), f"tet table {k} differs"
    assert np.array_equal(td["gradN"], to["gradN"]), "shape-function gradients not bit-identical"


def test_deterministic_tracking_bit_exact(device_lib, oracle_lib):
    """Zero-noise, frozen-velocity run in a closed all-slip box: tracking is
    purely geometric (tet walk + specular reflection)."""
    case = cases.synthetic_box(10, 6, 5, ppc=20, variant="deterministic", closed=True)
    P = case.random_particles(20, seed=11)
    dev, ora, reps = run_both(case, device_lib, oracle_lib, P, n_steps=12, rtol=RTOL)
    pd_, po = dev.download_particles(), ora.download_particles()
    # positions follow the same arithmetic: expect bit-identical coordinates too
    o1, o2 = np.argsort(pd_["id"]), np.argsort(po["id"])
    assert np.array_equal(pd_["pos"][o1], po["pos"][o2])
    assert sum(r[0].nLost for r in reps) == 0


def test_stochastic_closed_box_same_streams(device_lib, oracle_lib):
    """Full model chain (RAS Omega, IEM, cold reaction, SLM with noise, cell local
    time stepping, limitedSimple position correction, numerical diffusion)."""
    case = cases.synthetic_box(8, 5, 4, ppc=25, closed=True)
    P = case.random_particles(25, seed=5)
    run_both(case, device_lib, oracle_lib, P, n_steps=8, rtol=1e-10, atol=1e-12)


def test_open_box_injection_and_outflow(device_lib, oracle_lib):
    """inletOutlet patches: injection (inversion generator), outflow cull,
    deletion at inflow faces, open-boundary reflection."""
    case = cases.synthetic_box(8, 4, 4, ppc=20, closed=False)
    P = case.random_particles(20, seed=3)
    dev, ora, reps = run_both(case, device_lib, oracle_lib, P, n_steps=8, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nInjected for r in reps) > 0
    assert sum(r[0].nDeleted for r in reps) > 0


def test_particle_number_control(device_lib, oracle_lib):
    """clone / eliminate with id-keyed draws: identical populations.

    Cloning ranks a cell's particles by eta*m. Wherever the interpolated eta
    field is locally constant (e.g. the first step, deltaT = 1e-13, rescales eta
    to 1 + O(1e-13) everywhere) particles of equal mass have eta*m values that
    differ only in their last bits, and the ranking then depends on 1-ulp
    differences between CPU and GPU interpolation arithmetic — an inherent
    sensitivity of the reference algorithm (std::sort on near-ties), not a
    defect. The test gives the particles distinct masses and switches local time
    stepping off (eta == 1), so that eta*m values are either separated far beyond
    round-off or tie exactly (a clone and its source), which is broken by id."""
    case = cases.synthetic_box(6, 4, 4, ppc=20, closed=True, number_control=True, lts="none")
    case.params.deltaT0 = 2e-3
    P = case.random_particles(20, seed=9)
    # unbalance the population: drop most particles of some cells, triple others
    keep = np.ones(P["m"].shape[0], bool)
    rng = np.random.default_rng(0)
    for c in range(0, case.mesh.n_cells, 5):
        idx = np.nonzero(P["cell"] == c)[0]
        keep[idx[rng.random(idx.size) < 0.6]] = False
    Q = {k: (v[keep] if k != "Phi" else [a[keep] for a in v]) for k, v in P.items()}
    dup = np.nonzero(np.isin(Q["cell"], np.arange(1, case.mesh.n_cells, 7)))[0]
    R = {k: (np.concatenate([v, v[dup]]) if k != "Phi" else [np.concatenate([a, a[dup]]) for a in v]) for k, v in Q.items()}
    R["id"] = np.arange(R["m"].shape[0], dtype=np.uint64)
    R["pos"] = R["pos"].copy()
    R["pos"][-dup.size:] += 1e-4 * (rng.random((dup.size, 3)) - 0.5)
    R["m"] = R["m"] * rng.uniform(0.5, 1.5, size=R["m"].shape[0])
    dev, ora, reps = run_both(case, device_lib, oracle_lib, R, n_steps=5, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nCloned for r in reps) > 0
    assert sum(r[0].nEliminated for r in reps) > 0


def test_dense_cells_sorted_by_cell_and_tet(device_lib, oracle_lib):
    """>= 256 particles/cell switches the device sort key from cell to (cell, local tet);
    open box with number control so that every consumer of the sorted order runs."""
    case = cases.synthetic_box(4, 2, 2, ppc=256, closed=False, number_control=True, lts="none")
    case.params.deltaT0 = 2e-3
    P = case.random_particles(256, seed=21)
    rng = np.random.default_rng(4)
    P["m"] = P["m"] * rng.uniform(0.5, 1.5, size=P["m"].shape[0])
    dev, ora, reps = run_both(case, device_lib, oracle_lib, P, n_steps=4, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nInjected for r in reps) > 0


def test_seed_on_device_matches_oracle(device_lib, oracle_lib):
    case = cases.synthetic_box(6, 4, 3, ppc=16, closed=True)
    run_both(case, device_lib, oracle_lib, None, n_steps=3, seed_on_device=True, rtol=1e-10, atol=1e-12)


def test_product_does_not_load_oracle(device_lib):
    """The product path must not route through the CPU oracle."""
    with open("/proc/self/maps") as f:
        maps = f.read()
    assert "libpdfb200.so" in maps
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); from pdffoam_b200 import cases; from pdffoam_b200.capi import DeviceLibrary;"
            "c = cases.synthetic_box(4,3,3,ppc=8,closed=True); h = c.make_cloud(DeviceLibrary()); h.seed_particles(); h.evolve(2);"
            "m = open('/proc/self/maps').read(); assert 'libpdfb200.so' in m; assert 'pdforacle' not in m; print('ok')" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_midsize_open_box_jit_kernel_and_counting_sort(device_lib, oracle_lib):
    """40 x 20 x 20 cells x 64 particles (1.0e6): cells >> CTAs, so the run-time specialised kernel, the counting
    sort by cell, the Morton processing order, injection, outflow and number control are all compared with the
    oracle at a size where a CTA no longer sees a whole cell row."""
    case = cases.synthetic_box(40, 20, 20, ppc=64, closed=False, number_control=True)
    case.params.deltaT0 = 1e-3
    dev, ora, reps = run_both(case, device_lib, oracle_lib, None, n_steps=3, seed_on_device=True, rtol=1e-10, atol=1e-12)
    assert sum(r[0].nInjected for r in reps) > 0 and sum(r[0].nDeleted for r in reps) > 0
    assert sum(r[0].nCloned + r[0].nEliminated for r in reps) > 0
