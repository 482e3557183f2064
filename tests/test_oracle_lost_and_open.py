"""Oracle-level checks (CPU) of the paths a healthy run never takes: particles lost at an
`empty` patch of a 3-D mesh (mcEmptyBoundary.C:61-104 aborts the reference; here the particle
is dropped through notifyLostParticle, mcParticleCloud.C:2088-2092) with the re-spread of
their mass (mcParticleCloud.C:1363-1369), and the open boundary with `reflecting off` /
the plain `open` handler (mcOpenBoundary.C:159-205)."""
import numpy as np

from pdffoam_b200 import cases, meshgen


def lost_case(nx=6, ny=4, nz=3, ppc=12, number_control=False):
    """closed box whose zmax wall carries the `empty` handler on a generic patch of a 3-D mesh:
    every particle that reaches it is lost"""
    case = cases.synthetic_box(nx, ny, nz, ppc=ppc, closed=True, number_control=number_control, lts="none")
    case.mesh.patch("zmax").handler = meshgen.BC_EMPTY
    case.params.deltaT0 = 4e-3
    return case


def open_case(reflecting: int, handler: int, nx=8, ny=4, nz=3, ppc=12):
    case = cases.synthetic_box(nx, ny, nz, ppc=ppc, closed=False, number_control=False)
    pout = case.mesh.patch(case.info["outlet"])
    pout.reflecting = reflecting
    pout.handler = handler
    case.params.deltaT0 = 4e-3
    return case


def test_lost_particles_mass_is_respread(oracle_lib):
    case = lost_case()
    P = case.random_particles(12, seed=3)
    P["U"][:, 2] = np.abs(P["U"][:, 2]) + 8.0          # everything drifts towards zmax
    h = case.make_cloud(oracle_lib)
    h.upload_particles(P)
    m0 = P["m"].sum()
    n = P["m"].size
    lost = 0
    for _ in range(4):
        r = h.evolve(1)[0]
        lost += r.nLost
        n -= r.nLost
        Q = h.download_particles()
        assert Q["m"].size == n == r.nParticles
        # the mass of the lost particles went to the survivors of their cells (cells left empty keep it out of the cloud)
        assert Q["m"].sum() <= m0 * (1 + 1e-12)
    assert lost > 0
    assert Q["m"].sum() > 0.5 * m0


def test_reflecting_off_outlet_deletes_and_counts_mass(oracle_lib):
    for handler in (meshgen.BC_INLET_OUTLET, meshgen.BC_OPEN):
        case = open_case(0, handler)
        P = case.random_particles(12, seed=5)
        h = case.make_cloud(oracle_lib)
        h.upload_particles(P)
        deleted, mass_out = 0, 0.0
        for _ in range(5):
            r = h.evolve(1)[0]
            deleted += r.nDeleted
            mass_out += r.massOutInst[0]
            assert r.nLost == 0
        assert deleted > 0 and mass_out < 0           # mcOpenBoundary.C:187-193: massOut -= m
        # the reflecting outlet keeps those particles instead
        case_r = open_case(1, handler)
        hr = case_r.make_cloud(oracle_lib)
        hr.upload_particles(P)
        deleted_r = sum(hr.evolve(1)[0].nDeleted for _ in range(5))
        assert deleted_r < deleted
