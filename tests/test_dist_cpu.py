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
