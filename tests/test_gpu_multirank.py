"""Two-GPU test of the particle-sharded step: each rank seeds its share of every
cell, the per-cell moment block is summed with the library's NCCL all-reduce
(pdfb200_comm_init), and the union of both shards must reproduce the single-GPU
run (same global particle ids -> same RNG streams). Needs 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _make_case():
    from pdffoam_b200 import cases
    case = cases.synthetic_box(8, 5, 4, ppc=24, closed=True, number_control=False)
    case.params.deltaT0 = 2e-3
    return case


def _worker(rank, world, port, n_steps, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pdffoam_b200.capi import DeviceLibrary
    lib = DeviceLibrary()
    case = _make_case()
    h = case.make_cloud(lib, device=rank)
    uid = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        lib.check(lib.fn("comm_unique_id")(buf), "comm_unique_id")
        uid[0] = buf.raw
    dist.broadcast_object_list(uid, src=0)
    h.comm_init(world, rank, uid[0])
    h.seed_particles()
    reps = h.evolve(n_steps)
    P = h.download_particles(); cf = h.download_cell_fields()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=P["pos"], U=P["U"], m=P["m"], z=P["Phi"][0], cell=P["cell"], tet=P["tet"],
             id=P["id"], rho=cf["rho"], Ucell=cf["U"], k=cf["k"], zc=cf["Phi"][0], PaNIC=cf["PaNIC"], dt=h.delta_t(),
             ar_ms=sum(r.momentsAllreduceMs for r in reps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_run_matches_single_gpu(device_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    n_steps = 4
    case = _make_case()
    single = case.make_cloud(device_lib, device=0)
    single.seed_particles()
    single.evolve(n_steps)
    Ps = single.download_particles(); cs = single.download_cell_fields()
    mp.start_processes(_worker, args=(2, _free_port(), n_steps, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    for key, ref in (("rho", cs["rho"]), ("Ucell", cs["U"]), ("k", cs["k"]), ("zc", cs["Phi"][0]), ("PaNIC", cs["PaNIC"])):
        np.testing.assert_array_equal(r[0][key], r[1][key])
        np.testing.assert_allclose(r[0][key], ref, rtol=1e-11, atol=1e-13, err_msg=key)
    assert r[0]["dt"] == r[1]["dt"] == pytest.approx(single.delta_t(), rel=1e-13)
    ids = np.concatenate([r[0]["id"], r[1]["id"]])
    assert np.unique(ids).size == ids.size == Ps["id"].size
    o = np.argsort(ids); os_ = np.argsort(Ps["id"])
    for key, ref in (("cell", Ps["cell"]), ("tet", Ps["tet"])):
        assert np.array_equal(np.concatenate([r[0][key], r[1][key]])[o], ref[os_]), key
    for key, ref in (("pos", Ps["pos"]), ("U", Ps["U"]), ("m", Ps["m"]), ("z", Ps["Phi"][0])):
        np.testing.assert_allclose(np.concatenate([r[0][key], r[1][key]])[o], ref[os_], rtol=1e-10, atol=1e-12, err_msg=key)
    assert r[0]["ar_ms"] > 0
