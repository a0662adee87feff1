"""Host-side multi-process logic (sharding, cost reduction, best-solution gather) with world_size 2 on
the gloo backend.  No kernels are launched: the data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_robotics_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        g = torch.Generator().manual_seed(0)
        cost = torch.randn(37, 16, generator=g)                 # [b, h] cost of the whole job
        terms = torch.randn(37 * 16, 4, generator=g)
        local = D.shard_trajectories(cost)
        lo, hi = D.shard_bounds(37, rank, world_size)
        assert local.shape[0] == hi - lo
        tot = D.total_cost(local)
        tterms = D.total_cost_terms(D.shard_trajectories(terms.reshape(37, 16, 4)))
        q = torch.randn(101, 7, generator=g)
        err = torch.rand(101, generator=g)
        valid = torch.rand(101, generator=g) > 0.4
        l2, h2 = D.shard_bounds(101, rank, world_size)
        bq, berr, bvalid = D.gather_best_solutions(q[l2:h2], err[l2:h2], valid[l2:h2], k=5)
        # per-iteration valid counts of a 20-iteration solve: rank 0 (51 seeds) is all-valid from iteration 7, rank 1 (50) from 11
        cnt = torch.full((20,), 3, dtype=torch.int32)
        cnt[(7 if rank == 0 else 11):] = h2 - l2
        first = D.first_all_valid_iteration(cnt, h2 - l2)
        never = D.first_all_valid_iteration(torch.zeros(20, dtype=torch.int32), h2 - l2)
        flag_all = D.all_valid(True, "cpu")
        flag_one = D.all_valid(rank != 1, "cpu")
        ret[rank] = dict(tot=float(tot), terms=tterms, bq=bq, berr=berr, bvalid=bvalid, ref_tot=float(cost.double().sum()),
                         ref_terms=terms.double().sum(0), q=q, err=err, valid=valid, flag_all=flag_all, flag_one=flag_one, first=first, never=never)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 1000003):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo():
    world_size, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world_size, port, ret), nprocs=world_size, join=True)
        r0, r1 = ret[0], ret[1]
        for r in (r0, r1):
            assert abs(r["tot"] - r["ref_tot"]) < 1e-9
            assert torch.allclose(r["terms"], r["ref_terms"], atol=1e-9)
            assert r["flag_all"] is True and r["flag_one"] is False
            assert r["first"] == 11 and r["never"] == -1
        assert torch.equal(r0["bq"], r1["bq"]) and torch.equal(r0["berr"], r1["berr"])
        key = torch.where(r0["valid"], r0["err"], torch.full_like(r0["err"], float("inf")))
        top = torch.argsort(key, stable=True)[:5]
        assert torch.equal(r0["bq"], r0["q"][top]) and torch.equal(r0["berr"], r0["err"][top])
        assert bool(r0["bvalid"].all())


def test_single_process_is_identity():
    assert D.first_all_valid_iteration(torch.tensor([0, 2, 5, 5]), 5) == 2
    c = torch.arange(12.0).reshape(3, 4)
    assert float(D.total_cost(c)) == float(c.sum())
    q = torch.randn(9, 3)
    bq, berr, bvalid = D.gather_best_solutions(q, torch.arange(9.0), torch.ones(9, dtype=torch.bool), k=4)
    assert torch.equal(bq, q[:4]) and torch.equal(berr, torch.arange(4.0))
