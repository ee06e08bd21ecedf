"""CPU tier: the N>1 path on the gloo backend, world_size 2.  The sharding
and the single all-reduce of quimb_b200.dist are exercised with a numpy
stand-in for the device contraction (the product has no CPU arithmetic)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _np_contract(arrays, inputs, output, optimize):
    from oracle import contract_np as cn
    return torch.from_numpy(np.ascontiguousarray(
        cn.array_contract([np.asarray(a) for a in arrays], inputs, output, "greedy")))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quimb_b200 import dist as qd
    rng = np.random.default_rng(0)               # same data on every rank
    a = rng.standard_normal((4, 3, 5)); b = rng.standard_normal((5, 3, 6)); c = rng.standard_normal((6, 4))
    inputs = [("i", "s", "k"), ("k", "s", "l"), ("l", "t")]
    out, mine = qd.contract_sliced([a, b, c], inputs, ("i", "t"), sliced_inds=("s", "k"),
                                   contract_fn=_np_contract)
    ref = np.einsum("isk,ksl,lt->it", a, b, c)
    ok = np.allclose(out.numpy(), ref, atol=1e-12)
    # units are disjoint and cover everything
    allu = [None] * world
    dist.all_gather_object(allu, mine)
    flat = sorted(u for m in allu for u in m)
    cover = flat == list(range(15))
    ret[rank] = (bool(ok), bool(cover), len(mine))
    dist.destroy_process_group()


def test_slice_parallel_contraction_gloo_world2():
    mp.set_start_method("spawn", force=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [mp.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] and ret[1][1]
    assert ret[0][2] + ret[1][2] == 15 and abs(ret[0][2] - ret[1][2]) <= 1


def test_shard_units_round_robin():
    from quimb_b200 import dist as qd
    assert qd.shard_units(10, 0, 4) == [0, 4, 8]
    assert qd.shard_units(10, 3, 4) == [3, 7]
    assert sorted(sum((qd.shard_units(7, r, 3) for r in range(3)), [])) == list(range(7))
    assert qd.shard_units(2, 5, 8) == []
