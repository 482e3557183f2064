"""Error behaviour of the C ABI on a device: misuse must come back as a status with a message
(the reference aborts with FatalError), never as an out-of-bounds access that kills the CUDA
context."""
import os

import numpy as np
import pytest

from pdffoam_b200 import capi, cases
from parity_util import compare_particles

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_steady_flamelet_without_tables_is_refused(device_lib):
    """mcSteadyFlamelet's constructor fails when its library is missing (mcSteadyFlamelet.C:120-286)."""
    tab = cases.load_flamelet_tables(os.path.join(ROOT, "tests", "golden", "flamelet_sandiaD.npz"))
    case = cases.with_scalars(cases.synthetic_box(4, 3, 2, ppc=8, closed=True), ["z", "chi", "T"], "SteadyFlamelet", tab)
    case.tables = None                                   # "forgot" pdfb200_set_flamelet_tables
    h = case.make_cloud(device_lib)
    with pytest.raises(capi.PdfError, match="flamelet tables"):
        h.seed_particles()
    h.upload_particles(case.random_particles(8, seed=1))
    with pytest.raises(capi.PdfError, match="flamelet tables"):
        h.evolve(1)
    # a library with rho only, but one added scalar selected
    h.set_flamelet_tables(tab["chi"], tab["z"], tab["fields"][:1])
    with pytest.raises(capi.PdfError, match="fewer tables"):
        h.evolve(1)
    # the context is alive: with the full library the step runs
    h.set_flamelet_tables(tab["chi"], tab["z"], tab["fields"])
    assert h.evolve(1)[0].nParticles == case.mesh.n_cells * 8


@pytest.mark.parametrize("order", ["set_then_upload", "upload_then_set"])
def test_set_delta_t_survives_upload_and_seed(device_lib, oracle_lib, order):
    case = cases.synthetic_box(4, 3, 2, ppc=8, closed=True)
    P = case.random_particles(8, seed=2)
    for lib in (device_lib, oracle_lib):
        h = case.make_cloud(lib)
        if order == "set_then_upload":
            h.set_delta_t(3e-3); h.upload_particles(P)
        else:
            h.upload_particles(P); h.set_delta_t(3e-3)
        assert h.delta_t() == 3e-3
        assert h.evolve(1)[0].deltaT == 3e-3
    h = case.make_cloud(device_lib)
    h.set_delta_t(2e-3); h.seed_particles()
    assert h.evolve(1)[0].deltaT == 2e-3


def test_upload_with_bad_labels_is_refused_and_context_survives(device_lib):
    case = cases.synthetic_box(4, 3, 2, ppc=8, closed=True)
    h = case.make_cloud(device_lib)
    P = case.random_particles(8, seed=3)
    bad = dict(P); bad["cell"] = P["cell"].copy(); bad["cell"][5] = case.mesh.n_cells + 7
    with pytest.raises(capi.PdfError, match="cell label"):
        h.upload_particles(bad)
    bad["cell"][5] = -3
    with pytest.raises(capi.PdfError, match="cell label"):
        h.upload_particles(bad)
    bad = dict(P); bad["tet"] = np.zeros_like(P["cell"])         # tet 0 belongs to cell 0 only
    with pytest.raises(capi.PdfError, match="tet label"):
        h.upload_particles(bad)
    h.upload_particles(P)
    assert h.evolve(2)[-1].nParticles == P["m"].size


def test_default_ids_and_next_id(device_lib):
    """ids given on upload need not be 0..n-1 (a restart after number control): clones must not reuse one"""
    case = cases.synthetic_box(4, 3, 2, ppc=10, closed=True, number_control=True)
    case.params.deltaT0 = 2e-3
    P = case.random_particles(10, seed=4)
    P["id"] = (np.arange(P["m"].size, dtype=np.uint64) * 7 + 1000).astype(np.uint64)
    # thin out a few cells so that number control clones
    keep = ~np.isin(P["cell"], [1, 2, 3]) | (np.arange(P["m"].size) % 3 == 0)
    Q = {k: (v[keep] if k != "Phi" else [a[keep] for a in v]) for k, v in P.items()}
    h = case.make_cloud(device_lib)
    h.upload_particles(Q)
    reps = h.evolve(3)
    assert sum(r.nCloned for r in reps) > 0
    ids = h.download_particles()["id"]
    assert np.unique(ids).size == ids.size
    new = np.setdiff1d(ids, Q["id"])
    assert new.size > 0 and new.min() > Q["id"].max()
    # no ids given: running index
    R = dict(Q); R.pop("id")
    h2 = case.make_cloud(device_lib)
    h2.upload_particles(R)
    assert np.array_equal(np.sort(h2.download_particles()["id"]), np.arange(Q["m"].size, dtype=np.uint64))


def test_stage_and_commit_equal_a_plain_upload(device_lib):
    """pdfb200_stage_fv_fields + pdfb200_commit_fv_fields (asynchronous copy into a second buffer set, then a pointer
    swap) must leave the cloud in the state pdfb200_upload_fv_fields leaves it in, also when the staged copy is
    issued while an evolve() of the previous fields is what runs next"""
    case = cases.synthetic_box(6, 4, 3, ppc=12, closed=False, number_control=True)
    case.params.deltaT0 = 2e-3
    fv2 = {k: (v.copy() if v.dtype == np.uint8 else np.ascontiguousarray(v * (1.1 if k in ("U", "Ub", "k", "kb") else 1.0))) for k, v in case.fv.items()}
    fv1 = {k: np.ascontiguousarray(v) for k, v in case.fv.items()}
    a = case.make_cloud(device_lib); b = case.make_cloud(device_lib)
    a.seed_particles(); b.seed_particles()
    with pytest.raises(capi.PdfError, match="nothing staged"):
        b.commit_fv_fields()
    for step in range(4):
        fields = fv2 if step % 2 else fv1
        a.upload_fv_fields(fields)
        if step == 0:
            b.stage_fv_fields(fields)
        b.commit_fv_fields()
        b.stage_fv_fields(fv1 if step % 2 else fv2)       # the next state, copied while this step runs
        ra, rb = a.evolve(1)[0], b.evolve(1)[0]
        assert ra.deltaTNext == rb.deltaTNext and ra.nParticles == rb.nParticles and ra.UMax == rb.UMax
    ca, cb = a.download_cell_fields(), b.download_cell_fields()
    for k in ("rho", "U", "k", "Tau"):
        assert np.array_equal(ca[k], cb[k]), k


def test_ids_beyond_32_bits(device_lib, oracle_lib):
    """Particle ids are 64 bits wide (a 1e9-particle cloud with number control runs through 2^32 ids in a
    few hundred steps). Ids above 2^32 key their own random streams - the same ones on the device and in the oracle -
    survive the step, and the particles created afterwards are numbered above them."""
    case = cases.synthetic_box(6, 4, 3, ppc=16, closed=False, number_control=True)
    case.params.deltaT0 = 2e-3
    P = case.random_particles(16, seed=3)
    base = (1 << 33) + 12345
    P["id"] = (np.arange(P["m"].size, dtype=np.uint64) + np.uint64(base))
    dev, ora = case.make_cloud(device_lib, device=0), case.make_cloud(oracle_lib)
    dev.upload_particles(P); ora.upload_particles(P)
    assert dev.id_counter() == ora.id_counter() == base + P["m"].size
    for step in range(4):
        rd, ro = dev.evolve(1)[0], ora.evolve(1)[0]
        compare_particles(dev.download_particles(), ora.download_particles(), rtol=1e-10, atol=1e-12, what=f"step {step}")
    Q = dev.download_particles()
    assert Q["id"].dtype == np.uint64 and Q["id"].min() >= base
    assert np.setdiff1d(Q["id"], P["id"]).size > 0, "the open box injects and clones: new ids expected"
    assert dev.id_counter() == ora.id_counter() > base + P["m"].size
    # the high word is part of the stream key: the same particles with the low words alone evolve differently
    low = case.make_cloud(device_lib, device=0)
    P2 = dict(P); P2["id"] = P["id"] - np.uint64(1 << 33)
    low.upload_particles(P2); low.evolve(1)
    hi1 = case.make_cloud(device_lib, device=0); hi1.upload_particles(P); hi1.evolve(1)
    a, b = low.download_particles(), hi1.download_particles()
    n = min(a["U"].shape[0], b["U"].shape[0])
    assert not np.allclose(np.sort(a["U"][:n, 0]), np.sort(b["U"][:n, 0]))


def test_async_download_is_the_state_after_its_step(device_lib):
    """pdfb200_download_cell_fields_async: the read-back queued after step n must deliver step n's fields although
    step n+1 is issued (and overwrites the fields on the device) before the host waits for the copy"""
    import torch
    case = cases.synthetic_box(10, 6, 5, ppc=20, closed=False, number_control=True)
    case.params.deltaT0 = 2e-3
    a, b = case.make_cloud(device_lib), case.make_cloud(device_lib)
    a.seed_particles(); b.seed_particles()
    nc = case.mesh.n_cells
    pin = lambda *shape: torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    bufs = [dict(rho=pin(nc), U=pin(nc, 3), pndInst=pin(nc)) for _ in range(2)]
    for mode in (0, 1):                 # plain launches, then the CUDA-graph replay
        a.set_graph_mode(mode); b.set_graph_mode(mode)
        ref = []
        for n in range(6):
            a.evolve(1); cf = a.download_cell_fields(moments=False); ref.append({k: cf[k].copy() for k in bufs[0]})
        got = []
        for n in range(6):
            b.evolve(1)
            if n > 0:
                b.sync_downloads(); got.append({k: v.copy() for k, v in bufs[(n - 1) % 2].items()})
            b.download_fields_async(bufs[n % 2])
        b.sync_downloads(); got.append({k: v.copy() for k, v in bufs[5 % 2].items()})
        for n in range(6):
            for k in bufs[0]:
                assert np.array_equal(ref[n][k], got[n][k]), f"mode {mode}, step {n}, field {k}"
        assert not np.array_equal(ref[0]["U"], ref[5]["U"]) and not np.array_equal(ref[4]["pndInst"], ref[5]["pndInst"])


def test_profile_ctas_reports_every_cta_of_the_particle_kernel(device_lib):
    """pdfb200_profile_ctas: the first call arms the hook, later calls return start, end and SM of every CTA of the last
    launch of the particle kernel; arming it must not change the results"""
    case = cases.synthetic_box(12, 8, 6, ppc=32, closed=False, number_control=True)
    case.params.deltaT0 = 2e-3
    a, b = case.make_cloud(device_lib), case.make_cloud(device_lib)
    a.seed_particles(); b.seed_particles()
    assert b.profile_ctas().shape[0] == 0
    a.evolve(3); b.evolve(3)
    t = b.profile_ctas()
    assert t.shape[0] > 0 and np.all(t[:, 1] >= t[:, 0]) and np.all(t[:, 2] >= 0) and np.all(t[:, 2] < 1024)
    Pa, Pb = a.download_particles(), b.download_particles()
    oa, ob = np.argsort(Pa["id"]), np.argsort(Pb["id"])
    for k in ("pos", "U", "m"):
        assert np.array_equal(Pa[k][oa], Pb[k][ob]), k
