"""Helpers shared by the parity tests: run the same case on the CUDA library and
on the CPU oracle and compare particle columns (matched by particle id) and
cell fields."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def sort_by_id(P: Dict) -> Dict:
    o = np.argsort(P["id"], kind="stable")
    out = {}
    for k, v in P.items():
        if k == "Phi":
            out[k] = [a[o] for a in v]
        else:
            out[k] = v[o]
    return out


def compare_particles(pd_: Dict, po: Dict, rtol=1e-10, atol=1e-12, exact_index=True, what=""):
    pd_, po = sort_by_id(pd_), sort_by_id(po)
    assert pd_["id"].shape == po["id"].shape, f"{what}: particle counts differ {pd_['id'].shape} vs {po['id'].shape}"
    assert np.array_equal(pd_["id"], po["id"]), f"{what}: particle id sets differ"
    if exact_index:
        bad = np.nonzero(pd_["cell"] != po["cell"])[0]
        assert bad.size == 0, f"{what}: {bad.size} cell labels differ (first ids {pd_['id'][bad[:5]]})"
        bad = np.nonzero(pd_["tet"] != po["tet"])[0]
        assert bad.size == 0, f"{what}: {bad.size} tet labels differ (first ids {pd_['id'][bad[:5]]})"
    for k in ("pos", "U", "m", "Omega", "rho", "eta"):
        np.testing.assert_allclose(pd_[k], po[k], rtol=rtol, atol=atol, err_msg=f"{what}: column {k}")
    for i, (a, b) in enumerate(zip(pd_["Phi"], po["Phi"])):
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=f"{what}: Phi[{i}]")


def compare_cell_fields(cd: Dict, co: Dict, rtol=1e-10, atol=1e-12, what=""):
    for k in ("rho", "rhoInst", "pnd", "pndInst", "U", "Tau", "k", "PaNIC", "mMom", "VMom", "UMom", "UUMom"):
        np.testing.assert_allclose(cd[k], co[k], rtol=rtol, atol=atol, err_msg=f"{what}: cell field {k}")
    for name in ("Phi", "Cov", "PhiMom", "PhiPhiMom", "UPhi", "UPhiMom"):
        for i, (a, b) in enumerate(zip(cd[name], co[name])):
            np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=f"{what}: cell field {name}[{i}]")


def compare_reports(rd, ro, rtol=1e-10, what=""):
    for k in ("deltaT", "deltaTNext", "rhoRes", "rhoErrMin", "rhoErrMax", "UMax", "UcorrMax", "etaMin", "etaMax", "CoMax",
              "totalMass", "totalParticleMass"):
        a, b = getattr(rd, k), getattr(ro, k)
        assert abs(a - b) <= rtol * max(abs(b), 1e-300) + 1e-14, f"{what}: report.{k} {a} vs {b}"
    for k in ("nParticles", "nInjected", "nDeleted", "nCloned", "nEliminated", "nLost"):
        assert getattr(rd, k) == getattr(ro, k), f"{what}: report.{k} {getattr(rd, k)} vs {getattr(ro, k)}"


def run_both(case, device_lib, oracle_lib, particles: Optional[Dict], n_steps: int, seed_on_device=False,
             rtol=1e-10, atol=1e-12, exact_index=True, check_every_step=True, mass_rtol=1e-9):
    """Evolve the case on both implementations for n_steps and compare after every step."""
    dev = case.make_cloud(device_lib, device=0)
    ora = case.make_cloud(oracle_lib)
    if seed_on_device:
        dev.seed_particles(); ora.seed_particles()
        compare_particles(dev.download_particles(), ora.download_particles(), rtol, atol, exact_index, "after seeding")
    else:
        dev.upload_particles(particles); ora.upload_particles(particles)
    reports = []
    for s in range(n_steps):
        rd = dev.evolve(1)[0]; ro = ora.evolve(1)[0]
        reports.append((rd, ro))
        if check_every_step or s == n_steps - 1:
            compare_reports(rd, ro, max(rtol, 1e-10), f"step {s}")
            for k in range(1 + case.params.nConserved):
                for name in ("deltaMassInst", "massInInst", "massOutInst"):
                    a, b = getattr(rd, name)[k], getattr(ro, name)[k]
                    scale = max(abs(ro.totalParticleMass), 1e-300)
                    assert abs(a - b) <= mass_rtol * scale, f"step {s}: report.{name}[{k}] {a} vs {b}"
            compare_particles(dev.download_particles(), ora.download_particles(), rtol, atol, exact_index, f"step {s}")
            compare_cell_fields(dev.download_cell_fields(), ora.download_cell_fields(), rtol, atol, f"step {s}")
    return dev, ora, reports
