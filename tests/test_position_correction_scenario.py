"""The reference's own integration scenario for position correction
(tests/positionCorrectionTest: Allrun:16-30, simple/constant/polyMesh/blockMeshDict:33,
initPositionCorrectionTest.C:74-86, simple/system/mcSolution): a 6x2x2 box of 33x11x1
cells with open x faces, slip walls and empty front/back, U = (1,0,0), k = 1.5,
epsilon = 0.75, 30 particles per cell, the masses in the cell containing (1,1,1)
doubled, 200 steps with frozen FV fields. The reference only eyeballs the result;
here the density bump must relax with limitedSimple and must not without it.
"""
import numpy as np
import pytest

from pdffoam_b200 import cases, meshgen


def scenario(poscorr):
    mesh = meshgen.slab_mesh(33, 11, lo=(0.0, 0.0), hi=(6.0, 2.0), thickness=2.0,
                             xmin=("open0", meshgen.BC_INLET_OUTLET), xmax=("open1", meshgen.BC_INLET_OUTLET))
    nc, nb = mesh.n_cells, mesh.n_boundary_faces
    own = mesh.owner[mesh.n_internal_faces:]
    U = np.tile([1.0, 0.0, 0.0], (nc, 1))
    k = np.full(nc, 1.5)
    R = np.zeros((nc, 6)); R[:, 0] = R[:, 3] = R[:, 5] = 1.0                 # 2/3 k
    fv = dict(U=U, Ub=U[own].copy(), UbFixed=np.zeros(nb, np.uint8), p=np.zeros(nc), pb=np.zeros(nb), k=k, kb=k[own].copy(),
              kbFixed=np.zeros(nb, np.uint8), epsilon=np.full(nc, 0.75), epsilonb=np.full(nb, 0.75), R=R, Rb=R[own].copy())
    cf = dict(rho=np.ones(nc), rhob=np.ones(nb), rhobFixed=np.zeros(nb, np.uint8), Phi=[], Phib=[], PhibFixed=[])
    prm = cases.default_params(0, particlesPerCell=30, CFL=0.2, averagingCoeff=100.0, DNum=0.0, kMin=1e-8,
                               positionCorrection=poscorr, pcC=0.1, pcCFL=0.5, localTimeStepping="none",
                               relaxTimeU=1e-5, relaxTimek=1e-5, Omega0=1e5, deltaT0=1e-3, particleNumberControl=0)
    return cases.Case("positionCorrectionTest", mesh, prm, fv, cf)


def run(lib, poscorr, steps=200):
    case = scenario(poscorr)
    h = case.make_cloud(lib)
    h.seed_particles()
    P = h.download_particles()
    cc = case.mesh.cell_centres_simple()
    target = int(np.argmin(np.linalg.norm(cc[:, :2] - [1.0, 1.0], axis=1)))
    P["m"] = np.where(P["cell"] == target, 2 * P["m"], P["m"])
    h.upload_particles(P)
    err, res, lost = [], [], 0
    for s in range(steps):
        r = h.evolve(1)[0]
        lost += r.nLost
        if s >= 100 and s % 10 == 9:
            c = h.download_cell_fields(moments=False)
            err.append(float(c["pnd"][target] / c["rho"][target] - 1))
            res.append(r.rhoRes)
    h.close()
    return np.array(err), np.array(res), lost


def check(with_pc, without_pc):
    e1, r1, lost1 = with_pc
    e0, r0, lost0 = without_pc
    assert lost1 == 0 and lost0 == 0
    # the time-averaged particle number density of the doubled cell is back at the mean density
    assert abs(e1.mean()) < 0.04, e1
    # without position correction the excess is still there
    assert e0.mean() > 0.06, e0
    # and the density residual of the whole field (rhoRes of evolve) is lower with correction
    assert r1.mean() < 0.85 * r0.mean(), (r1.mean(), r0.mean())


def test_density_bump_relaxes_on_the_oracle(oracle_lib):
    check(run(oracle_lib, "limitedSimple"), run(oracle_lib, "none"))


@pytest.mark.gpu
def test_density_bump_relaxes_on_the_device(device_lib):
    check(run(device_lib, "limitedSimple"), run(device_lib, "none"))
