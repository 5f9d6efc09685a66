"""World-size-2 check of the multi-GPU host logic on CPU (gloo): batch slices are independent units, so
gathering per-rank results must reproduce the full-batch result bit for bit, and the bench's timing
reduction is a MAX over ranks.  The per-rank compute here is the oracle (no GPU in this test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cspn_monodepth_b200 import sharding
from tests.util import make_inputs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    from oracle import c_oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g, d, s = make_inputs(11, 5, 8, 1, 13, 17, density=0.05)      # 5 images over 2 ranks: 3 + 2
        sl = sharding.batch_slice(5, world, rank)
        y = torch.from_numpy(c_oracle.forward(g[sl], d[sl], s[sl], 24, 3, 0, threads=1))
        parts = [None] * world
        dist.all_gather_object(parts, y)
        t = torch.tensor([1.0 + rank])                                # pretend per-rank step time
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            full = c_oracle.forward(g, d, s, 24, 3, 0, threads=1)
            ret["equal"] = bool(np.array_equal(torch.cat(parts).numpy(), full))
            ret["tmax"] = float(t)
    finally:
        dist.destroy_process_group()


def test_two_rank_batch_sharding_is_exact():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["equal"] is True
    assert ret["tmax"] == 2.0
