#!/usr/bin/env python
"""Generates tests/golden/sandiaD_ensemble.npz: the "reference ensemble" of the stationary-state validation
(BASELINE north_star: "reproducing the reference's converged SandiaD mean and RMS profiles within the stated
statistical tolerance"; observables of tutorials/SandiaD/system/case.ini:80-140).

pdfFoam itself cannot run here (it needs OpenFOAM), so the ensemble comes from the CPU oracle in rngMode = 1: one
sequential random stream consumed in particle-list order, OpenFOAM's Random (48-bit LCG + polar Gaussian with cache)
- the reference's behaviour, different streams from the device's keyed Philox. Run in the build container:
    python tests/golden/make_sandiaD_ensemble.py            (8 members in parallel, about 8 minutes on 8 cores)

Case: pdffoam_b200.cases.sandiaD_validation_case() - the SandiaD tutorial's mesh (jet / pilot / coflow blocks, graded,
5 degree wedge), scalars (z chi T), SteadyFlamelet on the reference's flamelet library, IEM, RAS, cell local time
stepping, number control, inletOutlet patches, 40 particles per cell; frozen divergence-free jet field; no position
correction (without the FV coupling the density fields are not consistent and the correction has no fixed point).
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

N_MEMBERS = 8
N_STEPS = 4800
PROBE_STEPS = (3600, 4200, 4800)


def observables(c):
    """the fields of case.ini:80-140 per cell"""
    return dict(Ux=c["U"][:, 0], Ur=c["U"][:, 1], uRMS=np.sqrt(np.maximum(c["Tau"][:, 0], 0)), vRMS=np.sqrt(np.maximum(c["Tau"][:, 3], 0)),
                uv=c["Tau"][:, 1], z=c["Phi"][0], zRMS=np.sqrt(np.maximum(c["Cov"][0], 0)), T=c["Phi"][2],
                TRMS=np.sqrt(np.maximum(c["Cov"][5], 0)))


def member(seed):
    from pdffoam_b200 import cases
    from oracle.oracle import OracleLib
    case = cases.sandiaD_validation_case(seed=seed, rng_mode=1)
    h = case.make_cloud(OracleLib())
    h.seed_particles()
    probes, done = [], 0
    for s in PROBE_STEPS:
        h.evolve(s - done); done = s
        o = observables(h.download_cell_fields())
        probes.append({k: float(v.mean()) for k, v in o.items()})
    return observables(h.download_cell_fields()), probes


def main():
    t0 = time.time()
    with mp.get_context("fork").Pool(min(N_MEMBERS, os.cpu_count() or 1)) as pool:
        res = pool.map(member, [9000 + 17 * i for i in range(N_MEMBERS)])
    out = {"n_members": np.array(N_MEMBERS), "n_steps": np.array(N_STEPS)}
    for k in res[0][0]:
        a = np.stack([r[0][k] for r in res])
        out["mean_" + k] = a.mean(0); out["var_" + k] = a.var(0, ddof=1); out["domain_" + k] = a.mean(1)
        out["probe_" + k] = np.array([[p[k] for p in r[1]] for r in res])      # [member, probe step]: stationarity evidence
    out["probe_steps"] = np.array(PROBE_STEPS)
    np.savez_compressed(os.path.join(HERE, "sandiaD_ensemble.npz"), **out)
    print("wrote sandiaD_ensemble.npz in %.0f s" % (time.time() - t0))
    for k in ("Ux", "z", "T", "uRMS"):
        print(k, "ensemble-mean of the domain mean at steps", PROBE_STEPS, out["probe_" + k].mean(0), "member spread", out["probe_" + k].std(0, ddof=1))


if __name__ == "__main__":
    main()
