"""World-size-2 test of the particle-sharded step on CPU (gloo): each rank owns a
shard of the particles (mesh, fields replicated), the per-cell instantaneous
moment block is summed across ranks between evolve_begin and evolve_end, the
particle maxima are max-reduced. The union of the shards must
reproduce the single-rank run: identical particles (same counter-based RNG
streams, keyed by global particle id) and cell fields equal up to summation
order."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _make_case(closed=True):
    from pdffoam_b200 import cases
    case = cases.synthetic_box(6, 4, 3, ppc=16, closed=closed, number_control=False)
    case.params.deltaT0 = 2e-3     # a realistic first step (the default 1e-13 injects nothing)
    return case


def _worker(rank, world, port, n_steps, out_dir, closed):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pdffoam_b200.capi import StepReport
    from oracle.oracle import OracleLib, dp
    lib = OracleLib()
    case = _make_case(closed)
    h = case.make_cloud(lib)
    lib.fn("set_rank")(h.h, world, rank)
    h.seed_particles()
    nblk = int(lib.fn("moment_block_size")(h.h))
    block = np.zeros(nblk)
    rep = StepReport()
    for _ in range(n_steps):
        lib.check(lib.fn("evolve_begin")(h.h, C.byref(rep), dp(block)), "evolve_begin")
        tb = torch.from_numpy(block)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)                       # the moment all-reduce
        mx = torch.tensor([rep.CoMax, rep.UMax, rep.UcorrMax, rep.etaMax, -rep.etaMin], dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        rep.CoMax, rep.UMax, rep.UcorrMax, rep.etaMax, rep.etaMin = mx[0].item(), mx[1].item(), mx[2].item(), mx[3].item(), -mx[4].item()
        lib.check(lib.fn("evolve_end")(h.h, C.byref(rep), dp(block)), "evolve_end")
    P = h.download_particles()
    cf = h.download_cell_fields()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=P["pos"], U=P["U"], m=P["m"], z=P["Phi"][0], cell=P["cell"], tet=P["tet"],
             id=P["id"], rho=cf["rho"], Ucell=cf["U"], k=cf["k"], zc=cf["Phi"][0], PaNIC=cf["PaNIC"], dt=h.delta_t())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_open_box_injection_is_shared_between_ranks(oracle_lib, tmp_path):
    """inletOutlet faces are dealt round-robin to the ranks; the per-face RNG key does
    not depend on the rank, so after one step the union of both ranks' particles
    (old + injected) and the cell fields equal the single-rank run. Later steps
    differ only in the ids (hence noise streams) of the injected particles."""
    case = _make_case(closed=False)
    single = case.make_cloud(oracle_lib)
    single.seed_particles()
    rep = single.evolve(1)[0]
    assert rep.nInjected > 0
    Ps = single.download_particles(); cs = single.download_cell_fields()
    mp.start_processes(_worker, args=(2, _free_port(), 1, str(tmp_path), False), nprocs=2, join=True, start_method="fork")
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    # same particles were injected (the per-face streams do not depend on the rank); their ids
    # differ between the runs, hence their Langevin noise, so only noise-free quantities are exact
    assert r[0]["id"].size + r[1]["id"].size == Ps["id"].size
    m = np.sort(np.concatenate([r[0]["m"], r[1]["m"]]))
    np.testing.assert_allclose(m, np.sort(Ps["m"]), rtol=1e-13)
    np.testing.assert_array_equal(r[0]["PaNIC"], r[1]["PaNIC"])
    assert r[0]["PaNIC"].sum() == cs["PaNIC"].sum()
    np.testing.assert_allclose(r[0]["Ucell"], cs["U"], rtol=0, atol=5e-3)


def test_two_rank_sharded_oracle_matches_single_rank(oracle_lib, tmp_path):
    n_steps = 4
    case = _make_case()
    single = case.make_cloud(oracle_lib)
    single.seed_particles()
    single.evolve(n_steps)
    Ps = single.download_particles(); cs = single.download_cell_fields()

    port = _free_port()
    mp.start_processes(_worker, args=(2, port, n_steps, str(tmp_path), True), nprocs=2, join=True, start_method="fork")
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    # cell fields are identical on both ranks and equal the single-rank fields up to summation order
    for key, ref in (("rho", cs["rho"]), ("Ucell", cs["U"]), ("k", cs["k"]), ("zc", cs["Phi"][0]), ("PaNIC", cs["PaNIC"])):
        np.testing.assert_array_equal(r[0][key], r[1][key])
        np.testing.assert_allclose(r[0][key], ref, rtol=1e-11, atol=1e-13, err_msg=key)
    assert r[0]["dt"] == r[1]["dt"] == pytest.approx(single.delta_t(), rel=1e-13)
    # the union of the shards is the single-rank population
    ids = np.concatenate([r[0]["id"], r[1]["id"]])
    assert np.unique(ids).size == ids.size == Ps["id"].size
    o = np.argsort(ids); os_ = np.argsort(Ps["id"])
    assert np.array_equal(ids[o], Ps["id"][os_])
    for key, ref in (("cell", Ps["cell"]), ("tet", Ps["tet"])):
        assert np.array_equal(np.concatenate([r[0][key], r[1][key]])[o], ref[os_]), key
    for key, ref in (("pos", Ps["pos"]), ("U", Ps["U"]), ("m", Ps["m"]), ("z", Ps["Phi"][0])):
        np.testing.assert_allclose(np.concatenate([r[0][key], r[1][key]])[o], ref[os_], rtol=1e-10, atol=1e-12, err_msg=key)
    # both ranks hold about half of the particles
    assert abs(r[0]["id"].size - r[1]["id"].size) < 0.1 * ids.size


def _worker_unbalanced(rank, world, port, n_steps, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pdffoam_b200 import cases
    from pdffoam_b200.capi import StepReport
    from oracle.oracle import OracleLib, dp
    lib = OracleLib()
    case = cases.synthetic_box(6, 4, 3, ppc=16, closed=True, number_control=True)
    case.params.deltaT0 = 2e-3
    h = case.make_cloud(lib)
    lib.fn("set_rank")(h.h, world, rank)
    # 24 particles per cell (1.5 x the target: every cell eliminates), three quarters of them on rank 0
    P = case.random_particles(24, seed=5)
    mine = (np.arange(P["m"].size) % 4 != 0) if rank == 0 else (np.arange(P["m"].size) % 4 == 0)
    Q = {k: (v[mine] if k != "Phi" else [a[mine] for a in v]) for k, v in P.items()}
    h.upload_particles(Q)
    m0 = float(Q["m"].sum()); n0 = int(mine.sum())
    nblk = int(lib.fn("moment_block_size")(h.h))
    block = np.zeros(nblk)
    rep = StepReport()
    for _ in range(n_steps):
        lib.check(lib.fn("evolve_begin")(h.h, C.byref(rep), dp(block)), "evolve_begin")
        tb = torch.from_numpy(block)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        mx = torch.tensor([rep.CoMax, rep.UMax, rep.UcorrMax, rep.etaMax, -rep.etaMin], dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        rep.CoMax, rep.UMax, rep.UcorrMax, rep.etaMax, rep.etaMin = mx[0].item(), mx[1].item(), mx[2].item(), mx[3].item(), -mx[4].item()
        lib.check(lib.fn("evolve_end")(h.h, C.byref(rep), dp(block)), "evolve_end")
    P1 = h.download_particles()
    np.savez(os.path.join(out_dir, f"unb{rank}.npz"), n0=n0, n1=P1["m"].size, m0=m0, m1=float(P1["m"].sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_number_control_pulls_the_ranks_towards_equal_shares(oracle_lib, tmp_path):
    """Particle-sharded number control: every rank steers its holding of a cell towards 1/nRanks of the cell's target.
    (Shares in proportion to the local counts keep an imbalance for ever and let it grow by drift: DESIGN.md 5.)
    Start: 24 particles per cell against a target of 16, 18 of them on rank 0 and 6 on rank 1."""
    mp.start_processes(_worker_unbalanced, args=(2, _free_port(), 6, str(tmp_path)), nprocs=2, join=True, start_method="fork")
    r = [np.load(tmp_path / f"unb{k}.npz") for k in range(2)]
    a0, b0, a1, b1 = int(r[0]["n0"]), int(r[1]["n0"]), int(r[0]["n1"]), int(r[1]["n1"])
    assert a0 == 3 * b0
    # the population is back in the band of the target (72 cells x 16; one rank alone settles at 0.85 of it on this
    # case, and the rank below its share removes nothing while the other comes down to its share) and the split far
    # more even than 3 : 1
    assert 0.6 * 72 * 16 < a1 + b1 < 1.1 * 72 * 16, (a1, b1)
    assert abs(a1 - b1) < 0.25 * (a1 + b1), (a1, b1)
    # elimination conserves mass in expectation only; rank by rank it stays within a few per cent here
    assert abs((r[0]["m1"] + r[1]["m1"]) / (r[0]["m0"] + r[1]["m0"]) - 1) < 0.1
