"""Checkpoint files in the reference's format (pdffoam_b200/foamio.py): lagrangian/<cloud>/* as
mcParticle::writeFields / readFields lay them out (mcParticleIO.C:94-234) and the moment fields the cloud
reads READ_IF_PRESENT (mcParticleCloud.C:339-517). Round trip through the files, then a run continued from them
(oracle on the CPU; the same through the device when a GPU is there)."""
import os

import numpy as np
import pytest

from pdffoam_b200 import capi, cases, foamio

from parity_util import compare_cell_fields, compare_particles

NAMES = ("z", "T")


def _case():
    case = cases.synthetic_box(5, 4, 3, ppc=12, closed=False, number_control=True, n_phi=2)
    case.params.deltaT0 = 2e-3
    case.params.averagingCoeff = 20.0
    return case


def _continue_from_files(lib, tmp_path, device=0):
    case = _case()
    a = case.make_cloud(lib, device=device)
    a.seed_particles()
    a.evolve(4)
    P, cf = a.download_particles(), a.download_cell_fields()
    d = foamio.write_lagrangian(str(tmp_path), "0.004", P, NAMES)
    foamio.write_moments(str(tmp_path), "0.004", cf, NAMES)
    assert sorted(os.listdir(d)) == sorted(["positions", "origId", "origProcId", "m", "UParticle", "Ucorrection", "Omega", "rho", "eta", "z", "T"])
    head = open(os.path.join(d, "UParticle")).read(400)
    assert "class       vectorField;" in head and 'location    "0.004/lagrangian/thermoCloud";' in head
    Q = foamio.read_lagrangian(str(tmp_path), "0.004", NAMES)
    for k in ("pos", "U", "m", "Omega", "rho", "eta"):
        assert np.array_equal(Q[k], P[k]), k                      # repr() round-trips every double
    assert np.array_equal(Q["cell"], P["cell"]) and np.array_equal(Q["id"], P["id"])
    assert all(np.array_equal(x, y) for x, y in zip(Q["Phi"], P["Phi"]))
    mom = foamio.read_moments(str(tmp_path), "0.004", NAMES)
    for k in ("mMom", "VMom", "UMom", "UUMom"):
        assert np.array_equal(mom[k], cf[k]), k
    for name in ("PhiMom", "PhiPhiMom", "UPhiMom"):
        assert all(np.array_equal(x, y) for x, y in zip(mom[name], cf[name])), name
    # continue from the files
    cb = _case()
    own = cb.mesh.owner[cb.mesh.n_internal_faces:]
    cb.cf = dict(cb.cf, rho=cf["rho"], Phi=cf["Phi"], Cov=cf["Cov"], Covb=[c[own].copy() for c in cf["Cov"]],
                 CovbFixed=[np.zeros(cb.mesh.n_boundary_faces, np.uint8) for _ in cf["Cov"]])
    b = capi.CloudHandle(lib, cb.mesh, cb.params, device=device, capacity=int(P["m"].size * 1.5) + 4096)
    b.upload_fv_fields(cb.fv); b.upload_cloud_fields(cb.cf)
    b.upload_moments(mom)
    b.upload_particles(Q)                                          # no tet column in the files: located on upload
    b.set_delta_t(a.delta_t()); b.set_step_counter(a.step_counter()); b.set_id_counter(a.id_counter())
    a.evolve(2); b.evolve(2)
    compare_particles(a.download_particles(), b.download_particles(), 1e-11, 1e-13, True, "continued from files")
    compare_cell_fields(a.download_cell_fields(), b.download_cell_fields(), 1e-11, 1e-13, "continued from files")
    # a time directory without one of the reference's moment fields -> None (forced re-initialisation)
    os.remove(os.path.join(str(tmp_path), "0.004", "UUMoment"))
    assert foamio.read_moments(str(tmp_path), "0.004", NAMES) is None
    # a field file of the wrong length is refused (Cloud::checkFieldIOobject)
    foamio.write_field(os.path.join(d, "eta"), "eta", P["eta"][:-1], "0.004/lagrangian/thermoCloud")
    with pytest.raises(ValueError, match="does not match"):
        foamio.read_lagrangian(str(tmp_path), "0.004", NAMES)


def test_oracle_continues_from_checkpoint_files(oracle_lib, tmp_path):
    _continue_from_files(oracle_lib, tmp_path)


@pytest.mark.gpu
def test_device_continues_from_checkpoint_files(device_lib, tmp_path):
    _continue_from_files(device_lib, tmp_path)
