"""Size-independent properties of the CUDA path at sizes the CPU oracle cannot
reach: BASELINE.json configs[4] per-GPU shape (2e6 cells x 64 particles/cell =
1.28e8 particles) for the report-level checks, 250k cells x 64 for the checks
that download the particles.

Checked: run-to-run determinism (bit-identical reports and density field), the
population balance of every step, PaNIC summing to the particle count, id/mass
preservation in a closed box, every particle inside its cell and inside a tet
of that cell, and the cell-summed mass moment matching the reported total.
"""
import numpy as np
import pytest

from pdffoam_b200 import cases

pytestmark = pytest.mark.gpu

REPORT_FIELDS = ("deltaT", "deltaTNext", "rhoRes", "rhoErrMin", "rhoErrMax", "UMax", "UcorrMax", "etaMin", "etaMax",
                 "CoMax", "totalMass", "totalParticleMass", "nParticles", "nInjected", "nDeleted", "nCloned",
                 "nEliminated", "nLost")


def _run_open(device_lib, nx, ny, nz, steps):
    case = cases.synthetic_box(nx, ny, nz, ppc=64, closed=False, number_control=True)
    h = case.make_cloud(device_lib, capacity=int(nx * ny * nz * 64 * 1.25))
    h.seed_particles()
    n0 = h.num_particles()
    reps = [r.as_dict() for r in h.evolve(steps)]
    rho = h.download_rho().copy()
    cf = h.download_cell_fields(moments=False)
    h.close()
    return n0, reps, rho, cf


def test_baseline_size_determinism_and_balance(device_lib):
    """2e6 cells x 64 ppc, inletOutlet x faces, number control (the bench workload)."""
    nx, ny, nz, steps = 200, 100, 100, 6
    n0, repsA, rhoA, cfA = _run_open(device_lib, nx, ny, nz, steps)
    assert n0 == nx * ny * nz * 64
    n = n0
    for r in repsA:
        n = n + r["nInjected"] - r["nDeleted"] + r["nCloned"] - r["nEliminated"] - r["nLost"]
        assert r["nParticles"] == n, "population balance broken"
        assert r["nLost"] == 0
        assert np.isfinite(r["rhoRes"]) and r["CoMax"] > 0 and r["deltaTNext"] > 0
    assert sum(r["nInjected"] for r in repsA) > 0 and sum(r["nDeleted"] for r in repsA) > 0
    assert sum(r["nCloned"] for r in repsA) > 0
    # PaNIC (after number control) counts exactly the live particles
    assert int(round(cfA["PaNIC"].sum())) == repsA[-1]["nParticles"]
    assert np.all(np.isfinite(rhoA)) and rhoA.min() > 0
    del cfA
    # second run from the same seed: bit-identical reports and density field
    _, repsB, rhoB, _ = _run_open(device_lib, nx, ny, nz, steps)
    for a, b in zip(repsA, repsB):
        for k in REPORT_FIELDS:
            assert a[k] == b[k], f"report field {k} differs between identical runs"
        assert a["deltaMassInst"] == b["deltaMassInst"] and a["massInInst"] == b["massInInst"]
    assert np.array_equal(rhoA, rhoB), "density field differs between identical runs"


def test_large_closed_box_particles_stay_consistent(device_lib):
    """250k cells x 64 ppc, closed slip box, no number control: ids and masses are
    preserved, every particle lies inside its cell and its tet belongs to the cell."""
    nx, ny, nz = 100, 50, 50
    case = cases.synthetic_box(nx, ny, nz, ppc=64, closed=True, number_control=False)
    h = case.make_cloud(device_lib)
    h.seed_particles()
    P0 = h.download_particles()
    n0 = P0["id"].shape[0]
    assert n0 == nx * ny * nz * 64
    o0 = np.argsort(P0["id"])
    assert np.array_equal(P0["id"][o0], np.arange(n0, dtype=np.uint64))
    m0 = P0["m"][o0].copy()
    del P0
    reps = h.evolve(5)
    assert all(r.nLost == 0 and r.nDeleted == 0 and r.nInjected == 0 for r in reps)
    P = h.download_particles()
    o = np.argsort(P["id"])
    assert np.array_equal(P["id"][o], np.arange(n0, dtype=np.uint64)), "ids not preserved"
    assert np.array_equal(P["m"][o], m0), "particle masses changed in a closed box without number control"
    lo, hi = case.info["cell_lo"][P["cell"]], case.info["cell_hi"][P["cell"]]
    tol = 1e-12
    assert np.all(P["pos"] >= lo - tol) and np.all(P["pos"] <= hi + tol), "particle outside its cell"
    assert np.array_equal(P["tet"] // 24, P["cell"]), "tet does not belong to the particle's cell"
    # the moments saw every particle: cell-summed instantaneous mass == sum of eta*m
    cf = h.download_cell_fields(moments=False)
    vol = case.info["cell_volumes"]
    np.testing.assert_allclose((cf["pndInst"] * vol).sum(), (P["eta"] * P["m"]).sum(), rtol=1e-11)
    np.testing.assert_allclose(reps[-1].totalParticleMass, (P["eta"] * P["m"]).sum(), rtol=1e-11)
    assert int(round(cf["PaNIC"].sum())) == n0
    h.close()
