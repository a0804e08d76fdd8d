"""Synthetic batches for the configurations BASELINE.json names.

Harness code (bench.py, tests): builds `chars` + `offsets` tensors with torch, on
the GPU when one is present (the 1 M-string batches are ~2 GB) or on the CPU for
small test sizes.  Nothing here is on the match path.
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# table text -> small Python NFA (for GENERATING inputs only)
# --------------------------------------------------------------------------------------


def parse_table_text(text: str):
    hdr, edges = {}, []
    for line in text.splitlines():
        p = line.split()
        if not p or p[0] in ("rxm-tables", "end"):
            continue
        if p[0] in ("kind", "reversed", "states", "start", "finish", "cells", "edges"):
            hdr[p[0]] = p[1]
            continue
        frm, kind, sym, to = int(p[0]), p[1], p[2], int(p[3])
        byte = None
        if kind == "L":
            byte = int(sym[1:]) if sym.startswith("#") else ord(sym)
        edges.append((frm, kind, byte, to))
    return hdr, edges


def textbook_dfa(text: str):
    """Subset construction with textbook NFA semantics over the letters the table names.
    Returns (letters uint8[K], trans int64[K][S], accept bool[S], start, dist int64[S])
    where state 0 is dead and dist = fewest further letters to reach acceptance."""
    hdr, edges = parse_table_text(text)
    assert hdr["kind"] == "nfa"
    n = int(hdr["states"])
    start, finish = int(hdr["start"]), int(hdr["finish"])
    eps = [[] for _ in range(n)]
    lit = [[] for _ in range(n)]
    letters = sorted({b for (_, k, b, _) in edges if k == "L"})
    for frm, kind, byte, to in edges:
        if kind == "E":
            eps[frm].append(to)
        elif kind == "L":
            lit[frm].append((byte, to))
        elif kind == "A":
            lit[frm].append((None, to))

    def closure(s):
        st, seen = list(s), set(s)
        while st:
            q = st.pop()
            for t in eps[q]:
                if t not in seen:
                    seen.add(t)
                    st.append(t)
        return frozenset(seen)

    ids = {frozenset(): 0}
    sets = [frozenset()]
    s0 = closure({start})
    ids[s0] = 1
    sets.append(s0)
    rows = []
    i = 0
    while i < len(sets):
        row = []
        for b in letters:
            nx = set()
            for q in sets[i]:
                for (lb, to) in lit[q]:
                    if lb is None or lb == b:
                        nx.add(to)
            nx = closure(nx)
            if nx not in ids:
                ids[nx] = len(sets)
                sets.append(nx)
            row.append(ids[nx])
        rows.append(row)
        i += 1
    S, K = len(sets), len(letters)
    trans = np.array(rows, dtype=np.int64).T.reshape(K, S)
    accept = np.array([finish in s for s in sets], dtype=bool)
    INF = 1 << 40
    dist = np.where(accept, 0, INF).astype(np.int64)
    for _ in range(S):
        nd = dist.copy()
        for k in range(K):
            nd = np.minimum(nd, dist[trans[k]] + 1)
        if np.array_equal(nd, dist):
            break
        dist = nd
    return np.array(letters, dtype=np.uint8), trans, accept, 1, dist


# --------------------------------------------------------------------------------------
# generators
# --------------------------------------------------------------------------------------


def _lengths(n, lo, hi, gen, device):
    lens = torch.randint(lo, hi + 1, (n,), generator=gen, device=device, dtype=torch.int64)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=offsets[1:])
    return lens, offsets


def uniform_strings(n, lo, hi, alphabet: bytes, seed, device):
    """n strings, length uniform in [lo, hi], letters i.i.d. uniform over `alphabet`
    (BASELINE.json configs[1] read literally; these die within a few letters on config 2)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    lens, offsets = _lengths(n, lo, hi, gen, device)
    total = int(offsets[-1])
    alpha = torch.tensor(list(alphabet), dtype=torch.uint8, device=device)
    chars = alpha[torch.randint(0, len(alphabet), (total,), generator=gen, device=device)]
    return chars, offsets


def alive_strings(table_text: str, n, lo, hi, seed, device, flip_frac=0.5):
    """Random walks on the automaton that end in acceptance (so no prefix ever kills
    the active set and the matcher must read every byte), then the LAST letter of a
    `flip_frac` share of the strings is replaced by another letter (SURVEY 8d,
    config 2 "alive" set).  Letters are uniform among the choices that can still
    reach acceptance in the letters that remain."""
    letters, trans, accept, start, dist = textbook_dfa(table_text)
    K = len(letters)
    gen = torch.Generator(device=device).manual_seed(seed)
    lens, offsets = _lengths(n, lo, hi, gen, device)
    total = int(offsets[-1])
    chars = torch.empty(total, dtype=torch.uint8, device=device)
    T = torch.from_numpy(trans).to(device)          # [K][S]
    D = torch.from_numpy(dist).to(device)           # [S]
    LET = torch.from_numpy(letters).to(device)
    state = torch.full((n,), start, dtype=torch.int64, device=device)
    base = offsets[:-1]
    for p in range(hi):
        active = lens > p
        if not bool(active.any()):
            break
        rem = lens - p - 1
        nxt = T[:, state]                           # [K][n]
        ok = D[nxt] <= rem.unsqueeze(0)             # can still be accepted
        alive = nxt != 0
        score = torch.rand((K, n), generator=gen, device=device)
        score = torch.where(ok, score + 2.0, torch.where(alive, score + 1.0, score))
        choice = score.argmax(0)
        idx = (base + p)[active]
        chars[idx] = LET[choice[active]]
        state = torch.where(active, nxt.gather(0, choice.unsqueeze(0)).squeeze(0), state)
    if flip_frac > 0 and K > 1:
        flip = (torch.rand(n, generator=gen, device=device) < flip_frac) & (lens > 0)
        last = (offsets[1:] - 1)[flip]
        cur = chars[last]
        pos = (cur.unsqueeze(1) == LET.unsqueeze(0)).int().argmax(1)
        chars[last] = LET[(pos + 1) % K]
    return chars, offsets


def example5_strings(n, lo, hi, seed, device, corrupt_frac=0.5, kmax=64):
    """BASELINE.json configs[2]: x c x c x^m with x = a^k, k in [0, kmax], m chosen so
    that the length lands in [lo, hi]; `corrupt_frac` of the strings get one letter of
    the last block replaced by `b`."""
    gen = torch.Generator(device=device).manual_seed(seed)
    k = torch.randint(0, kmax + 1, (n,), generator=gen, device=device, dtype=torch.int64)
    target = torch.randint(lo, hi + 1, (n,), generator=gen, device=device, dtype=torch.int64)
    m = torch.where(k > 0, torch.clamp((target - 2 * k - 2) // torch.clamp(k, min=1), min=0),
                    torch.zeros_like(k))
    lens = 2 * k + 2 + k * m
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=offsets[1:])
    total = int(offsets[-1])
    chars = torch.full((total,), ord("a"), dtype=torch.uint8, device=device)
    base = offsets[:-1]
    chars[base + k] = ord("c")
    chars[base + 2 * k + 1] = ord("c")
    bad = (torch.rand(n, generator=gen, device=device) < corrupt_frac) & (k > 0) & (m > 0)
    # one letter inside the last block
    within = (torch.rand(n, generator=gen, device=device) * k.clamp(min=1)).long().clamp(max=kmax)
    pos = (offsets[1:] - 1 - within)[bad]
    chars[pos] = ord("b")
    return chars, offsets


def pumped_string(nn, pump):
    """matchers/example_runner.cpp:15-29 (== matcher.py:26-38)."""
    pump_count = len(pump) // 2 + 1
    del_count = len(pump) - pump_count
    res = pump[0]
    while len(res) + len(pump[0]) < (nn - del_count) // pump_count:
        res += pump[0]
    return res if len(pump) == 1 else (res + pump[1]) * del_count + res


def attack_strings(pump, suffix, prefix, sizes):
    """Attack strings of matchers/example_runner.cpp:15-29 / matcher.py:26-38
    (non-cumulative), with and without the failing suffix.  -> list[bytes]"""
    def pumped(nn):
        return pumped_string(nn, pump)
    out = []
    for s in sizes:
        p = pumped(s)
        out.append((prefix + p + suffix).encode())
        out.append((prefix + p).encode())
    return out


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous string-index range of `rank` (shard by string index)."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    hi = min(n_total, lo + per)
    return lo, hi


# --------------------------------------------------------------------------------------
# configs 4 and 5: attack strings / mixed example batches (host-side generation, numpy)
# --------------------------------------------------------------------------------------

README_EXAMPLES = {  # README.md:78-89 (test/example_N/regexp.txt + pump.txt)
    1: ("({a*}:1&1)*", ["a"], "b", ""),
    2: ("{(a|bb)*}:1aaba(&1|bb*aa)*", ["bbaa", "aaba", "bbaa"], "c", ""),
    3: ("{{a*}:1(&1)*}:2b&2a*", ["a", "b", "a"], "aab", ""),
    4: ("({a*}:1&1a*)*", ["a"], "b", ""),
    5: ("{a*}:1c{&1}:2c(&1|&2)*", ["aa"], "b", "aacaac"),
    6: ("({a*}:1b|&1)*", ["a"], "c", "aaab"),
    7: ("({a*}:1)*b&1", ["a", "b", "a"], "b", ""),
    8: ("(({a*}:1|b)(&1|b))*", ["a", "b", "a"], "c", "bb"),
    9: ("(({aa*b}:1(&1)*)|b(b|a*)*)*", ["bbaaa"], "c", ""),
    10: ("({a*}:1b|b&1)*c&1", ["aababba"], "cab", ""),
}


def _pumped(nn, pump):
    pump_count = len(pump) // 2 + 1
    del_count = len(pump) - pump_count
    reps = max(1, ((nn - del_count) // pump_count - 1) // len(pump[0]))
    res = pump[0] * reps
    while len(res) + len(pump[0]) < (nn - del_count) // pump_count:
        res += pump[0]
    return res if len(pump) == 1 else (res + pump[1]) * del_count + res


def attack_batch(pump, suffix, prefix, n, lo, hi, seed, log_uniform=True):
    """n attack strings prefix + pumped(size) [+ suffix] (matchers/example_runner.cpp:15-29,
    matcher.py:26-38), size log-uniform (or uniform) in [lo, hi]; half carry the failing suffix.
    -> (chars uint8[], offsets uint64[n+1]) numpy."""
    rng = np.random.default_rng(seed)
    if log_uniform:
        sizes = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)
    else:
        sizes = rng.integers(lo, hi + 1, size=n)
    with_suffix = rng.random(n) < 0.5
    parts = []
    for sz, ws in zip(sizes, with_suffix):
        body = _pumped(int(sz), pump)
        parts.append((prefix + body + (suffix if ws else "")).encode())
    lens = np.fromiter((len(p) for p in parts), dtype=np.uint64, count=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    return chars, offsets


def mixed_example_batch(example, n, seed, lo=16, hi=512):
    """Config 5 per-example batch: one third pumped attack strings, one third near misses of
    them (one edit), one third random strings over the expression's letters."""
    regex, pump, suffix, prefix = README_EXAMPLES[example]
    rng = np.random.default_rng(seed)
    letters = sorted({c for c in regex if c.isalpha()})
    parts = []
    for j in range(n):
        kind = j % 3
        sz = int(rng.integers(lo, hi + 1))
        if kind == 2:
            parts.append(bytes(rng.choice(np.frombuffer("".join(letters).encode(), dtype=np.uint8), size=sz)))
            continue
        s = prefix + _pumped(sz, pump) + (suffix if rng.random() < 0.5 else "")
        if kind == 1 and len(s) > 1:
            k = int(rng.integers(0, len(s)))
            s = s[:k] + letters[int(rng.integers(0, len(letters)))] + s[k + 1:]
        parts.append(s.encode())
    lens = np.fromiter((len(p) for p in parts), dtype=np.uint64, count=n)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer(b"".join(parts), dtype=np.uint8).copy()
    return chars, offsets
