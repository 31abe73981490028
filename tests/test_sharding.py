"""N > 1 path on CPU: the batch shards by contiguous column ranges with no collective.
Two gloo ranks each evaluate their shard (the oracle stands in for the device), all_gather the pieces and
must reproduce the unsharded result bit for bit — columns are independent (unittest/parallel-rnea.cpp:54)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, load_model, random_inputs


def test_column_range_partition():
    from pinocchio_b200.sharding import column_range
    for B in (0, 1, 7, 64, 1000, 65536):
        for W in (1, 2, 3, 4, 8):
            rs = [column_range(B, W, r) for r in range(W)]
            assert rs[0][0] == 0 and rs[-1][1] == B
            assert all(rs[k][1] == rs[k + 1][0] for k in range(W - 1))
            assert max(c1 - c0 for c0, c1 in rs) == -(-B // W)
    with pytest.raises(ValueError):
        column_range(10, 2, 2)


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import Oracle
    from pinocchio_b200.sharding import column_range, run_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = load_model("humanoid_random")
    orc = Oracle(model)
    q, v, a = random_inputs(model, B, 5)
    tau_local = run_sharded(lambda q_, v_, a_: orc.rnea(q_, v_, a_), (q, v, a), world, rank)
    c0, c1 = column_range(B, world, rank)
    assert tau_local.shape == (model.nv, c1 - c0)
    per = -(-B // world)
    pad = np.zeros((model.nv, per))
    pad[:, :c1 - c0] = tau_local
    pieces = [torch.zeros(model.nv, per, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(pieces, torch.from_numpy(pad))  # test-only gather; the product path has no collective
    if rank == 0:
        full = np.concatenate([p.numpy() for p in pieces], axis=1)[:, :B]
        np.save(os.path.join(out_dir, "tau.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_unsharded(tmp_path, oracle_cls):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    B = 101  # odd: the second rank gets the short shard
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    model = load_model("humanoid_random")
    q, v, a = random_inputs(model, B, 5)
    ref = oracle_cls(model).rnea(q, v, a)
    got = np.load(os.path.join(str(tmp_path), "tau.npy"))
    assert np.array_equal(got, ref)
