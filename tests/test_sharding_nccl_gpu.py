"""GPU tier, world_size >= 2 over NCCL (skipped on a box with one GPU): ONE batch cut by bytes
(sharding.shard_by_bytes), every rank matches its slice on its own GPU through the C ABI, the result bits
are gathered and the match counts summed over NCCL (sharding.gather_bits / total_matches) -- the gathered
vector must equal the oracle's bits for the whole batch on every rank.  SURVEY.md section 8(e); the same
host logic runs over gloo with the oracle standing in for the GPU in tests/test_sharding_gloo.py."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import helpers as H
from cases import load_case

pytestmark = pytest.mark.gpu


def _load_sharding():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(H.PKG, "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _batch(name):
    """golden strings + a seeded random batch with very unequal lengths (so that byte balance != count balance)"""
    t, strings, _ = load_case(name)
    rng = np.random.default_rng(12)
    alpha = np.frombuffer(b"ab" if name.startswith("nfa") else b"aabbc", dtype=np.uint8)
    extra = [bytes(rng.choice(alpha, size=int(L))) for L in rng.integers(0, 60, size=6000)]
    extra += [bytes(rng.choice(alpha, size=int(L))) for L in rng.integers(2000, 9000, size=60)]
    chars, off = H.make_batch(strings + extra)
    return t, chars, off


def _worker(rank, world, port, name, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    S = _load_sharding()
    t, chars, off = _batch(name)
    want = H.oracle_bits(t, chars, off)
    bounds = S.shard_by_bytes(off, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    c, o = S.local_view(chars, off, lo, hi)
    m = H.rxm.Matcher(t, rank)
    local = m.match_host(np.ascontiguousarray(c), o)
    m.close()
    dev = torch.device("cuda", rank)
    full = S.gather_bits(local, bounds, rank, world, dist, device=dev).cpu().numpy()
    total = S.total_matches(local, dist, device=dev)
    ok = bool(np.array_equal(full, want)) and total == int(want.sum())
    q.put((rank, ok, lo, hi, int(off[hi] - off[lo])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL gather of the result bits)")
@pytest.mark.parametrize("name", ["nfa_config2", "ex05_fwd", "ex02_rev"])
def test_one_batch_sharded_over_the_gpus_gathers_the_oracle_bits(name):
    world = min(torch.cuda.device_count(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    spans = sorted((lo, hi) for (_, _, lo, hi, _) in res)
    assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [r[4] for r in res]
    assert max(sizes) - min(sizes) <= 2 * 9000  # balanced by bytes
