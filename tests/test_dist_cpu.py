"""CPU, gloo, world_size 2: the host-side multi-rank plumbing (unique-id exchange, max-over-ranks
timing reduction, block-column ownership map).  The NCCL data path itself needs GPUs (test_gpu_multi)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpb200 import dist as gd
    uid = gd.broadcast_bytes(bytes(range(128)) if rank == 0 else None, src=0)
    m = gd.max_over_ranks(10.0 + rank)

    class FakeEngine:                      # records what the real Engine would receive
        def nccl_unique_id(self):
            return bytes([7] * 128)

        def comm_init(self, n, r, u):
            self.args = (n, r, u)

    e = FakeEngine()
    w, r = gd.init_engine_comm(e, p2p=False)
    out[rank] = (uid == bytes(range(128)), m, w, r, e.args[0], e.args[1], e.args[2] == bytes([7] * 128))
    dist.destroy_process_group()


def test_gloo_world2_plumbing():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gaussianprocesses.jl_b200"))
    world, port = 2, _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        uid_ok, m, w, r, n_, r_, id_ok = out[rank]
        assert uid_ok and id_ok
        assert m == 11.0                       # max over ranks of 10 + rank
        assert (w, r, n_, r_) == (2, rank, 2, rank)


def test_block_column_ownership_is_balanced():
    # mirror of cholesky_dist's owner(b) = b % R over Npad / dist_nb block columns
    for Np, nb, R in [(32768, 1024, 8), (32768, 1024, 2), (4096, 128, 4)]:
        nblk = Np // nb
        work = np.zeros(R)
        for b in range(nblk):
            work[b % R] += (Np - b * nb) * nb          # trapezoid area owned
        assert work.max() / work.mean() < 1.25


def _predict_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gpb200
    from test_host_cpu import OracleEngineCV, make_data

    class Eng(OracleEngineCV):                 # counts the rows this rank was asked to predict
        rows = 0

        def predict(self, xs, alpha=None, want_var=True, full_cov=False):
            Eng.rows += xs.shape[0]
            return super().predict(xs, alpha, want_var, full_cov)

        def storage_info(self):
            return dict(sharded=False)

    X, y, Xs = make_data(60, 2, 5, m=11)       # 11 test points over 3 ranks: slices 3 / 4 / 4
    eng = Eng()
    gp = gpb200.GPE.__new__(gpb200.GPE)
    eng.bind(gp)
    gpb200.GPE.__init__(gp, X.T, y, gpb200.MeanConst(0.3), gpb200.SEIso(0.2, 0.1), -1.0, engine=eng)
    gp._world = world
    mu, s2 = gp.predict_f(Xs.T, partition=True)
    rows_part = Eng.rows
    mu1, s21 = gp.predict_f(Xs.T)              # every rank predicts everything (the replicated default)
    few = gp.predict_f(Xs[:2].T, partition=True)          # fewer points than ranks: one rank has an empty slice
    few1 = gp.predict_f(Xs[:2].T)
    err = None
    try:
        gp.predict_f(Xs.T, full_cov=True, partition=True)
    except ValueError as e:
        err = str(e)
    same = lambda a, b: bool(a.shape == b.shape and np.allclose(a, b, rtol=1e-11, atol=1e-13))   # numpy BLAS on a slice vs on
    out[rank] = (rows_part, same(mu, mu1) and same(s2, s21),                                      # the whole block: last-bit noise
                 same(few[0], few1[0]) and same(few[1], few1[1]), err is not None)
    dist.destroy_process_group()


def test_partitioned_predict_f_gloo_world3():
    """predict_f(partition=True): each rank predicts only its slice of the test points and every rank ends up with the same,
    complete mean / variance vectors as an unpartitioned call (host-side slicing + all-gather; engine stubbed by the oracle)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "gaussianprocesses.jl_b200"))
    from gpb200.dist import partition_bounds
    assert partition_bounds(11, 3) == [0, 3, 7, 11] and partition_bounds(2, 3) == [0, 0, 1, 2] and partition_bounds(0, 4) == [0] * 5
    world, port = 3, _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_predict_worker, args=(world, port, out), nprocs=world, join=True)
    assert [out[r][0] for r in range(world)] == [3, 4, 4]
    for r in range(world):
        assert out[r][1] and out[r][2] and out[r][3], (r, out[r])
