"""CPU tier, world_size 2 over gloo: the N>1 host logic.  Each rank matches its byte-balanced
shard (with the oracle standing in for the GPU, which this box does not have) and the
gathered result must equal the whole-batch result; the match-count reduction must agree."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers as H
from cases import load_case


def _load_sharding():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(H.PKG, "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, name, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = _load_sharding()
    t, strings, bits = load_case(name)
    chars, off = H.make_batch(strings)
    bounds = S.shard_by_bytes(off, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    c, o = S.local_view(chars, off, lo, hi)
    local = H.oracle_bits(t, c, o)
    full = S.gather_bits(local, bounds, rank, world, dist).numpy()
    total = S.total_matches(local, dist)
    ok = bool(np.array_equal(full, bits)) and total == int(bits.sum())
    q.put((rank, ok, lo, hi))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ex02_fwd", "nfa_config2"])
def test_two_ranks_gloo(name):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for (_, ok, _, _) in res)
    spans = sorted((lo, hi) for (_, _, lo, hi) in res)
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0]


def test_shard_by_bytes_is_a_balanced_partition():
    S = _load_sharding()
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 5000, size=10000)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    for world in (1, 2, 3, 4, 8):
        b = S.shard_by_bytes(off, world)
        assert b[0] == 0 and b[-1] == len(lens) and all(x <= y for x, y in zip(b, b[1:]))
        sizes = [int(off[b[r + 1]] - off[b[r]]) for r in range(world)]
        assert max(sizes) - min(sizes) <= 2 * 5000
    # degenerate: fewer strings than ranks, empty strings
    off = np.array([0, 0, 3], dtype=np.uint64)
    b = S.shard_by_bytes(off, 4)
    assert b[0] == 0 and b[-1] == 2 and all(x <= y for x, y in zip(b, b[1:]))
