"""GPU tier: the CUDA path, called through the C ABI, against the golden vectors
(bits produced by running the reference) and against the oracle on seeded batches.
Bit-exact: this is byte/integer work."""
import ctypes as C

import os

import numpy as np
import pytest

import helpers as H
from cases import BY_NAME, CASE_NAMES, load_case

rxm = H.rxm
pytestmark = pytest.mark.gpu


def _engine_cases():
    out = []
    for name in CASE_NAMES:
        if BY_NAME[name]["kind"] == "nfa":
            out.append((name, None))
        else:
            out += [(name, "k2"), (name, "k3"), (name, "k4")]
    return out


def _matcher(t, engine):
    """engine: None (planner's choice) or "k2"/"k3"/"k4" (rxm_tables_upload_opts)."""
    m = rxm.Matcher(t, 0, engine=engine)
    if engine:
        assert rxm.ENGINE_NAMES[m.plan().engine] == {"k2": "K2_THREAD", "k3": "K3_WARP", "k4": "K4_THREAD"}[engine]
    return m


@pytest.mark.parametrize("name,engine", _engine_cases())
def test_golden_host_buffers(name, engine):
    t, strings, bits = load_case(name)
    chars, off = H.make_batch(strings)
    m = _matcher(t, engine)
    got = m.match_host(chars, off)
    assert m.launch_count() >= 1
    assert m.overflow_count() == 0
    assert np.array_equal(got, bits), [strings[i] for i in np.nonzero(got != bits)[0][:5]]
    m.close()


@pytest.mark.parametrize("name", ["nfa_config2", "nfa_abb", "ex05_fwd", "ex02_rev"])
def test_golden_device_pointers(name):
    import torch
    t, strings, bits = load_case(name)
    chars, off = H.make_batch(strings)
    d_chars = torch.from_numpy(chars).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_out = torch.full((len(strings),), 7, dtype=torch.uint8, device="cuda")
    m = rxm.Matcher(t, 0)
    stream = torch.cuda.current_stream().cuda_stream
    m.match_ptrs(d_chars.data_ptr(), d_off.data_ptr(), len(strings), d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), bits)
    assert m.overflow_count() == 0
    m.close()


def _random_batch(rng, n, lo, hi, alphabet):
    lens = rng.integers(lo, hi + 1, size=n)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    chars = rng.choice(np.frombuffer(alphabet, dtype=np.uint8), size=int(off[-1])).astype(np.uint8)
    return chars, off


@pytest.mark.parametrize("name", ["nfa_config2", "nfa_quirk", "nfa_abb", "nfa_dots", "nfa_third", "nfa_mid"])
def test_nfa_random_batches_vs_oracle(name):
    t, _, _ = load_case(name)
    rng = np.random.default_rng(123)
    m = rxm.Matcher(t, 0)
    for (n, lo, hi) in ((5000, 0, 40), (3000, 1, 300), (257, 1000, 5000)):
        chars, off = _random_batch(rng, n, lo, hi, b"ab")
        assert np.array_equal(m.match_host(chars, off), H.oracle_bits(t, chars, off))
    m.close()


@pytest.mark.parametrize("engine", ["k2", "k3", "k4"])
@pytest.mark.parametrize("name", ["ex01_fwd", "ex02_fwd", "ex02_rev", "ex05_fwd", "ex05_rev",
                                  "ex08_rev", "ex09_fwd", "ex14_rev", "ex15_rev", "ex17_rev"])
def test_mfa_random_batches_vs_oracle(name, engine):
    t, _, _ = load_case(name)
    rng = np.random.default_rng(7)
    m = _matcher(t, engine)
    for (n, lo, hi, alpha) in ((4000, 0, 30, b"ab"), (2000, 1, 60, b"aaab"), (1000, 1, 40, b"abc"),
                               (300, 100, 400, b"aaaaab")):
        chars, off = _random_batch(rng, n, lo, hi, alpha)
        got = m.match_host(chars, off)
        want = H.oracle_bits(t, chars, off)
        assert np.array_equal(got, want), int((got != want).sum())
    m.close()


@pytest.mark.parametrize("engine", ["k2", "k3", "k4"])
def test_mfa_long_blocks_and_attack_strings(engine):
    """Long backreference blocks (idle-step skipping, warp-wide block compare) and the
    reference's attack strings (pump.txt recipes) up to 16 K chars, forward and reversed."""
    W = H.load_workloads()
    cases = [("ex05_fwd", ["aa"], "b", "aacaac"), ("ex05_rev", ["aa"], "b", "aacaac"),
             ("ex02_fwd", ["bbaa", "aaba", "bbaa"], "c", ""), ("ex02_rev", ["bbaa", "aaba", "bbaa"], "c", ""),
             ("ex01_fwd", ["a"], "b", ""), ("ex09_fwd", ["bbaaa"], "c", "")]
    for name, pump, suffix, prefix in cases:
        t, _, _ = load_case(name)
        strings = W.attack_strings(pump, suffix, prefix, [50, 300, 1000, 4000, 16000])
        for k in (100, 1000, 5000):  # x c x c x x x with a long x
            x = b"a" * k
            strings += [x + b"c" + x + b"c" + x + x + x, x + b"c" + x + b"c" + x + x[:-1] + b"b" + x]
        chars, off = H.make_batch(strings)
        want = H.oracle_bits(t, chars, off)
        m = _matcher(t, engine)
        got = m.match_host(chars, off)
        assert np.array_equal(got, want), (name, [len(strings[i]) for i in np.nonzero(got != want)[0]])
        m.close()


def test_empty_batch_and_empty_strings():
    t, _, _ = load_case("ex01_fwd")
    m = rxm.Matcher(t, 0)
    out = m.match_host(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64))
    assert out.shape == (0,)
    chars, off = H.make_batch([b"", b"aa", b"", b"b", b""])
    assert list(m.match_host(chars, off)) == [1, 1, 1, 0, 1]
    m.close()


def test_plan_reports_engine():
    t, _, _ = load_case("nfa_config2")
    m = rxm.Matcher(t, 0)
    p = m.plan()
    assert rxm.ENGINE_NAMES[p.engine] == "K1_DFA"
    assert p.dfa_states == 13 and p.dfa_classes == 3 and p.exact_step_differs == 0
    assert p.sm_count > 0
    m.close()
    t, _, _ = load_case("ex05_fwd")
    m = rxm.Matcher(t, 0)
    assert rxm.ENGINE_NAMES[m.plan().engine] == "K4_THREAD"
    m.close()
    # 2^8 active sets: still a table, in its two-lookup form (byte class, then [class][set] u16)
    t, _, _ = load_case("nfa_mid")
    m = rxm.Matcher(t, 0)
    p = m.plan()
    assert rxm.ENGINE_NAMES[p.engine] == "K1_DFA" and p.dfa_states == 386 and p.dfa_stride == 4  # Q[set][16]: four bytes per lookup
    m.close()


@pytest.mark.skipif(not os.path.exists(H.RXM_COMPILE), reason="bin/rxm_compile not built")
@pytest.mark.parametrize("regex,lo,hi", [("(a|b)*a(a|b)(a|b)(a|b)b(a|b)*", 33, 64), ("(a|b)*a(a|b)(a|b)(a|b)(a|b)b(a|b)*", 65, 128),
                                         ("(a|b|c)*a(a|b|c)(a|b|c)b(a|b|c)*", 33, 64),
                                         ("(a|b|c)*a(a|b|c)(a|b|c)(a|b|c)b(a|b|c)*", 65, 128)])
def test_k1_tables_of_33_to_128_sets(regex, lo, hi):
    """K1 with 16 / 32 KB of static tables next to its 192 KB of rows (SP = 64: T and a stride table; SP = 128: T
    alone): the kernels' dynamic shared-memory limit must leave room for them (a fixed 216 KB was refused by
    cudaFuncSetAttribute -- found by tests/fuzz/fuzz_tables_gpu.py); every stride of the automaton against the oracle."""
    t = rxm.Tables(H.compile_tables_text(regex))
    rng = np.random.default_rng(21)
    alpha = np.frombuffer(b"abc" if "c" in regex else b"ab", dtype=np.uint8)
    strings = [bytes(rng.choice(alpha, size=int(n))) for n in rng.integers(0, 700, size=3000)]
    strings += [bytes(rng.choice(alpha, size=int(n))) for n in rng.integers(3000, 5000, size=64)]
    chars, off = H.make_batch(strings)
    want = H.oracle_bits(t, chars, off)
    assert 0 < int(want.sum()) < len(want)
    seen = set()
    for flags in (0, rxm.OPT_K1_NO_OCT, rxm.OPT_K1_NO_QUAD):
        m = rxm.Matcher(t, 0, flags=flags)
        p = m.plan()
        assert rxm.ENGINE_NAMES[p.engine] == "K1_DFA" and lo <= p.dfa_states <= hi, p.dfa_states
        seen.add(p.dfa_stride)
        assert np.array_equal(m.match_host(chars, off), want), (flags, p.dfa_stride)
        m.close()
    assert 1 in seen


def test_host_offsets_that_run_backwards_are_rejected_and_the_handle_stays_usable():
    """rxm_match_batch with host buffers checks the offsets (while the copy of the strings runs): a string that starts
    behind its end is RXM_ERR_INVALID, no kernel sees it, and the next call on the handle is answered as usual."""
    t, strings, bits = load_case("nfa_abb")
    chars, off = H.make_batch(strings)
    bad = off.copy()
    k = len(bad) // 2
    bad[k] = bad[k + 1] + 5
    m = rxm.Matcher(t, 0)
    with pytest.raises(rxm.RxmError) as e:
        m.match_host(chars, bad)
    assert e.value.status == rxm.RXM_ERR_INVALID
    bad = off.copy()
    bad[0] = off[-1] + 1
    with pytest.raises(rxm.RxmError) as e:
        m.match_host(chars, bad)
    assert e.value.status == rxm.RXM_ERR_INVALID
    assert np.array_equal(m.match_host(chars, off), bits)
    m.close()


def test_mixed_host_device_pointers_are_rejected():
    import torch
    t, strings, _ = load_case("nfa_abb")
    chars, off = H.make_batch(strings)
    d_chars = torch.from_numpy(chars).cuda()
    out = np.empty(len(strings), dtype=np.uint8)
    m = rxm.Matcher(t, 0)
    with pytest.raises(rxm.RxmError) as e:
        m.match_ptrs(d_chars.data_ptr(), off.ctypes.data, len(strings), out.ctypes.data)
    assert e.value.status == rxm.RXM_ERR_INVALID
    m.close()


@pytest.mark.parametrize("name", ["nfa_config2", "nfa_abb", "nfa_dots", "nfa_mid"])
def test_k1_every_length_and_alignment(name):
    """Forward (config2) and right-to-left (abb, dots: reversed Glushkov) scans: every
    length 0..300 at every start alignment mod 16, plus a few long strings, so that the
    head/tail vectors of the staged scan are exercised in every position."""
    t, _, _ = load_case(name)
    rng = np.random.default_rng(99)
    strings = []
    for L in range(0, 301):
        strings.append(bytes(rng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=L)))
    for L in (511, 512, 513, 1000, 4095, 4096, 4097, 20000):
        strings.append(bytes(rng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=L)))
    # strings that stay alive: walk inside the language
    strings += [b"aaba" + b"a" * k for k in range(0, 70)]
    strings += [b"a" * k + b"abb" for k in range(0, 70)]
    chars, off = H.make_batch(strings)
    want = H.oracle_bits(t, chars, off)
    m = rxm.Matcher(t, 0)
    assert np.array_equal(m.match_host(chars, off), want)
    # shifted copy: prepend 1..15 junk bytes so every string changes alignment
    import torch
    for shift in (1, 7, 15):
        d_chars = torch.cat([torch.zeros(shift, dtype=torch.uint8),
                             torch.from_numpy(chars)]).cuda()
        d_off = torch.from_numpy(off.astype(np.int64)).cuda()
        d_out = torch.empty(len(strings), dtype=torch.uint8, device="cuda")
        m.match_ptrs(d_chars.data_ptr() + shift, d_off.data_ptr(), len(strings), d_out.data_ptr(),
                     torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want), shift
    m.close()


@pytest.mark.parametrize("name", ["nfa_config2", "nfa_quirk", "nfa_abb", "nfa_dots", "nfa_third", "nfa_alt"])
def test_k1_quad_stride_and_bytes_outside_the_window(name):
    """The eight- and four-bytes-per-lookup interiors (dfa_stride 8 / 4) against the oracle and against
    the one-byte scan (RXM_OPT_K1_NO_QUAD), on long strings that stay alive and carry bytes outside
    the stride's letter window -- singly, in every position of a 16-byte vector, and in runs."""
    t, _, _ = load_case(name)
    rng = np.random.default_rng(5)
    ab = np.frombuffer(b"ab", dtype=np.uint8)
    strings = []
    for L in (64, 200, 1000, 3000):
        for _ in range(40):
            strings.append(bytes(rng.choice(ab, size=L)))
    base = (b"aaba" + b"a" * 200 + b"bb" * 30 + b"aa" + b"abb") * 3
    strings.append(base)
    for junk in (b"c", b"d", b"`", b"z", b"\x00", b"\xff", b"A", b"1"):
        for pos in range(20, 20 + 33):
            s = bytearray(base)
            s[pos] = junk[0]
            strings.append(bytes(s))
        s = bytearray(base)
        s[100:140] = junk * 40
        strings.append(bytes(s))
    for _ in range(200):  # mostly-window strings with a few outsiders
        s = bytearray(rng.choice(ab, size=int(rng.integers(100, 600))).tobytes())
        for p in rng.integers(0, len(s), size=3):
            s[p] = int(rng.choice(np.frombuffer(b"abcd.`z", dtype=np.uint8)))
        strings.append(bytes(s))
    chars, off = H.make_batch(strings)
    want = H.oracle_bits(t, chars, off)
    m = rxm.Matcher(t, 0)
    stride = m.plan().dfa_stride
    got = m.match_host(chars, off)
    m.close()
    m4 = rxm.Matcher(t, 0, flags=rxm.OPT_K1_NO_OCT)
    stride4 = m4.plan().dfa_stride
    got4 = m4.match_host(chars, off)
    m4.close()
    m1 = rxm.Matcher(t, 0, flags=rxm.OPT_K1_NO_QUAD)
    assert m1.plan().dfa_stride == 1
    got1 = m1.match_host(chars, off)
    m1.close()
    assert np.array_equal(got1, want)
    assert np.array_equal(got4, want), (stride4, int((got4 != want).sum()))
    assert np.array_equal(got, want), (stride, int((got != want).sum()))
    assert stride4 in (1, 4) and stride in (stride4, 8)
    if name in ("nfa_config2", "nfa_abb"):  # literals a, b only: eight bytes per lookup, four on request
        assert stride == 8 and stride4 == 4


@pytest.mark.parametrize("mode", ["masks", "walk"])
@pytest.mark.parametrize("name", [n for n in CASE_NAMES if BY_NAME[n]["kind"] == "nfa"])
def test_bitset_engine_on_every_memory_free_fixture(name, mode):
    """K1B (the active set as a 128-bit mask, the reference's exact step per letter) in both of its
    forms -- bit-parallel follow masks and the edge walk: forced on every memory-free fixture,
    golden bits + random batches against the oracle; forward, right-to-left, Thompson, the
    visited quirk."""
    t, strings, bits = load_case(name)
    m = rxm.Matcher(t, 0, engine="bitset", flags=rxm.OPT_K1B_WALK if mode == "walk" else 0)
    assert rxm.ENGINE_NAMES[m.plan().engine] == "K1_BITSET"
    assert (m.plan().dfa_classes > 0) == (mode == "masks")
    chars, off = H.make_batch(strings)
    assert np.array_equal(m.match_host(chars, off), bits)
    rng = np.random.default_rng(17)
    for (n, lo, hi, alpha) in ((3000, 0, 40, b"ab"), (500, 1, 300, b"ab"), (300, 1, 60, b"abcd.")):
        chars, off = _random_batch(rng, n, lo, hi, alpha)
        assert np.array_equal(m.match_host(chars, off), H.oracle_bits(t, chars, off))
    assert m.overflow_count() == 0
    m.close()


def test_planner_hands_large_determinisations_to_the_bitset_engine():
    """nfa_huge has ~98 000 reachable active sets: no table fits, K1B instead of RXM_ERR_UNSUPPORTED;
    long strings that stay alive, raw-text route included."""
    t, strings, bits = load_case("nfa_huge")
    m = rxm.Matcher(t, 0)
    assert rxm.ENGINE_NAMES[m.plan().engine] == "K1_BITSET"
    chars, off = H.make_batch(strings)
    assert np.array_equal(m.match_host(chars, off), bits)
    rng = np.random.default_rng(4)
    ab = np.frombuffer(b"ab", dtype=np.uint8)
    long_strings = [bytes(rng.choice(ab, size=int(L))) for L in rng.integers(1000, 6000, size=64)]
    chars, off = H.make_batch(long_strings)
    want = H.oracle_bits(t, chars, off)
    assert np.array_equal(m.match_host(chars, off), want)
    assert np.array_equal(m.match_text_host(b"\n".join(long_strings)), want)
    # a batch large enough for the tile-sorted hand-out order (>= 4 tiles of 4096 strings)
    chars, off = _random_batch(rng, 20000, 0, 200, b"ab")
    assert np.array_equal(m.match_host(chars, off), H.oracle_bits(t, chars, off))
    m.close()


@pytest.mark.skipif(not os.path.exists(H.RXM_COMPILE), reason="bin/rxm_compile not built")
@pytest.mark.parametrize("regex,sets,stride", [("(a|b)*a" + "(a|b)" * 4 + "b(a|b)*", 98, 4),
                                               ("(a|b)*a" + "(a|b)" * 5 + "b(a|b)*", 194, 4),
                                               ("(b|a)*a" + "(a|b)" * 8 + "(a|b)*abb", 1218, 4)])
@pytest.mark.parametrize("right_to_left", [0, 1])
def test_k1_two_lookup_tables_with_strides(regex, sets, stride, right_to_left):
    """K1's two-lookup form (more than 64 sets over a two-letter window, more than 128 otherwise) with its stride
    table (four bytes per lookup): both strides against the oracle, forward and with the tables read right-to-left;
    bytes outside the window, long strings, a tile-sorted batch."""
    text = H.compile_tables_text(regex)
    if right_to_left:
        text = text.replace("reversed 0", "reversed 1")
    t = rxm.Tables(text)
    rng = np.random.default_rng(6)
    ab, abz = np.frombuffer(b"ab", dtype=np.uint8), np.frombuffer(b"abbaz", dtype=np.uint8)
    strings = [bytes(rng.choice(ab, size=int(n))) for n in rng.integers(0, 300, size=18000)]
    strings += [bytes(rng.choice(abz, size=int(n))) for n in rng.integers(0, 200, size=1500)]
    strings += [bytes(rng.choice(ab, size=int(n))) for n in rng.integers(3000, 6000, size=64)]
    chars, off = H.make_batch(strings)
    want = H.oracle_bits(t, chars, off)
    assert 0 < int(want.sum()) < len(want)
    seen = []
    for flags in (0, rxm.OPT_K1_NO_OCT, rxm.OPT_K1_NO_QUAD):
        m = rxm.Matcher(t, 0, flags=flags)
        p = m.plan()
        assert rxm.ENGINE_NAMES[p.engine] == "K1_DFA" and p.dfa_states == sets
        seen.append(p.dfa_stride)
        assert np.array_equal(m.match_host(chars, off), want), (flags, p.dfa_stride)
        m.close()
    assert seen == [stride, stride, 1]


def test_k1_holds_the_largest_table_that_fits_shared_memory():
    """nfa_blowup: 24 577 active sets, a 172 KB two-lookup table next to the ring (round 1's planner stopped at 4096
    sets and K1B ran it at 38 GB/s): golden bits, long strings that stay alive, the raw-text route, a tile-sorted
    batch -- and the bit-set engine on the same inputs."""
    t, strings, bits = load_case("nfa_blowup")
    m = rxm.Matcher(t, 0)
    p = m.plan()
    assert rxm.ENGINE_NAMES[p.engine] == "K1_DFA" and p.dfa_states == 24578 and p.dfa_classes == 3 and p.dfa_stride == 1
    b = rxm.Matcher(t, 0, engine="bitset")
    assert rxm.ENGINE_NAMES[b.plan().engine] == "K1_BITSET"
    chars, off = H.make_batch(strings)
    assert np.array_equal(m.match_host(chars, off), bits)
    rng = np.random.default_rng(4)
    ab = np.frombuffer(b"ab", dtype=np.uint8)
    long_strings = [bytes(rng.choice(ab, size=int(L))) for L in rng.integers(1000, 6000, size=64)]
    long_strings += [b"b" * 3000, b"a" * 3000, b"b" * 2000 + b"a" + b"b" * 12, b"b" * 2000 + b"a" + b"b" * 13, b"ab" * 7 + b"b" * 4000]
    chars, off = H.make_batch(long_strings)
    want = H.oracle_bits(t, chars, off)
    assert 0 < int(want.sum()) < len(want)
    assert np.array_equal(m.match_host(chars, off), want)
    assert np.array_equal(b.match_host(chars, off), want)
    assert np.array_equal(m.match_text_host(b"\n".join(long_strings)), want)
    chars, off = _random_batch(rng, 20000, 0, 200, b"ab")
    want = H.oracle_bits(t, chars, off)
    assert np.array_equal(m.match_host(chars, off), want)
    assert np.array_equal(b.match_host(chars, off), want)
    chars, off = _random_batch(rng, 3000, 0, 60, b"abz.")  # bytes of class 0
    assert np.array_equal(m.match_host(chars, off), H.oracle_bits(t, chars, off))
    m.close()
    b.close()


def test_k1_large_batch_properties():
    """BASELINE-sized property checks (no oracle at this size): the result vector is a
    pure function of each string -- a permuted batch gives the permuted bits, and a batch
    split in two gives the same bits as the whole."""
    import torch
    t, _, _ = load_case("nfa_config2")
    W = H.load_workloads()
    n = 200_000
    chars, off = W.alive_strings(t.text, n, 64, 4096, 5, "cuda")
    m = rxm.Matcher(t, 0)
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s)
    torch.cuda.synchronize()
    frac = float(out.float().mean())
    assert 0.45 < frac < 0.55
    # split
    h = n // 2
    out2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    m.match_ptrs(chars.data_ptr(), off.data_ptr(), h, out2.data_ptr(), s)
    off_b = (off[h:] - 0).contiguous()
    m.match_ptrs(chars.data_ptr(), off_b.data_ptr(), n - h, out2.data_ptr() + h, s)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    # oracle on a sample
    k = 2000
    off_h = off[:k + 1].cpu().numpy().astype(np.uint64)
    want = H.oracle_bits(t, chars[:int(off_h[-1])].cpu().numpy(), off_h)
    assert np.array_equal(out[:k].cpu().numpy(), want)
    m.close()


# ---- rxm_match_text: tokenisation on the device (match.cpp:22-24) -----------------------------
def _ref_tokens(text: bytes):
    """`cin >> text` until the token "exit": bytes.split() splits on the same six bytes."""
    toks = text.split()
    if b"exit" in toks:
        toks = toks[:toks.index(b"exit")]
    return toks


def _check_text(m, t, text: bytes):
    toks = _ref_tokens(text)
    got = m.match_text_host(text)
    assert len(got) == len(toks), (len(got), len(toks))
    assert m.saw_exit.value == (1 if b"exit" in text.split() else 0)
    if toks:
        chars, off = H.make_batch(toks)
        assert np.array_equal(got, H.oracle_bits(t, chars, off))


@pytest.mark.parametrize("name", ["nfa_config2", "ex05_fwd", "ex02_rev"])
def test_match_text_equals_cin_tokenisation(name):
    t, strings, bits = load_case(name)
    m = rxm.Matcher(t, 0)
    rng = np.random.default_rng(11)
    seps = [b" ", b"\n", b"\t", b"\r\n", b"\v", b"\f", b"  \n\n ", b"\t \r"]
    # the golden strings, with every kind of separator, no trailing separator
    all_strings = strings
    strings = [s for s in strings if s and s != b"exit" and not any(c in s for c in b" \t\n\v\f\r")]
    keep = set(strings)
    text = b"".join(s + seps[int(rng.integers(len(seps)))] for s in strings[:-1]) + strings[-1]
    got = m.match_text_host(text)
    assert np.array_equal(got, bits[[i for i, s in enumerate(all_strings) if s in keep]])
    _check_text(m, t, text)
    # leading / trailing whitespace, exit in the middle, exit-like tokens that are not the sentinel
    _check_text(m, t, b"  \n" + text + b" \n")
    _check_text(m, t, b"aaba exits aexit ex it exit aaba aaba")
    _check_text(m, t, b"exit aaba")
    _check_text(m, t, b"aaba\nexit")
    _check_text(m, t, b"aaba\nexit\n")
    for blank in (b"", b" ", b"\n\n\n", b" \t\r\n\v\f"):
        assert m.match_text_host(blank).shape == (0,)
    # bytes that are NOT whitespace for `cin >>`: NUL, 0x1c-0x1f, 0x85, 0xa0
    _check_text(m, t, b"aa\x00ba ab\x1cab \x85 \xa0aaba a\x0eb")
    m.close()


def test_match_text_tokens_across_piece_boundaries():
    """Tokens that straddle the tokeniser's 64-byte thread pieces and 16 KB block pieces, runs of
    one-letter tokens (the densest case), and a token that ends on the very last byte."""
    t, _, _ = load_case("nfa_config2")
    m = rxm.Matcher(t, 0)
    rng = np.random.default_rng(3)
    ab = np.frombuffer(b"ab", dtype=np.uint8)
    parts = []
    for L in list(range(1, 130)) + [16383, 16384, 16385, 40000, 63, 64, 65] + [1] * 300:
        parts.append(bytes(rng.choice(ab, size=L)))
    for sep in (b" ", b"\n", b"  "):
        text = sep.join(parts)
        _check_text(m, t, text)
        _check_text(m, t, text + sep)
    # one-letter tokens only: 2 bytes per token
    _check_text(m, t, b"a b " * 20000)
    # a long walk inside the language, then many of them
    alive = [b"aaba" + b"a" * int(k) for k in rng.integers(0, 3000, size=400)]
    _check_text(m, t, b"\n".join(alive))
    m.close()


def test_match_text_device_pointers_alignment_and_capacity():
    import torch
    t, _, _ = load_case("nfa_config2")
    m = rxm.Matcher(t, 0)
    toks = [b"aaba", b"ab", b"bbaaba", b"aabaa", b"b" * 100, b"aababbaa"] * 50
    text = b"\n".join(toks) + b"\n"
    chars, off = H.make_batch(toks)
    want = H.oracle_bits(t, chars, off)
    for shift in (0, 1, 5, 15):
        d = torch.cat([torch.full((shift,), ord("a"), dtype=torch.uint8),
                       torch.frombuffer(bytearray(text), dtype=torch.uint8)]).cuda()
        out = torch.full((len(toks) + 3,), 9, dtype=torch.uint8, device="cuda")
        n = m.match_text_ptrs(d.data_ptr() + shift, len(text), out.data_ptr(), out.numel(),
                              torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert n == len(toks)
        assert np.array_equal(out[:n].cpu().numpy(), want), shift
        assert out[n:].eq(9).all()
    # too little room: nothing is matched, the count comes back
    with pytest.raises(rxm.RxmError) as e:
        m.match_text_host(text, cap=10)
    assert e.value.status == rxm.RXM_ERR_INVALID and e.value.n_tokens == len(toks)
    # mixed host / device pointers
    d = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    host_out = np.empty(len(toks), dtype=np.uint8)
    with pytest.raises(rxm.RxmError) as e:
        m.match_text_ptrs(d.data_ptr(), len(text), host_out.ctypes.data, len(toks))
    assert e.value.status == rxm.RXM_ERR_INVALID
    m.close()


@pytest.mark.parametrize("engine", ["k4", "k3", "k2"])
def test_mfa_large_batch_properties(engine):
    """BASELINE-shaped MFA batch (config 3: example 5 on x c x c x^m strings of 64-4096; 200 k strings
    for K3, 20 k for K2): the bits are a pure function of each string -- a permuted batch gives the
    permuted bits, a split batch the same bits, the raw-text route the same bits -- and a sample
    equals the oracle."""
    import torch
    t, _, _ = load_case("ex05_fwd")
    W = H.load_workloads()
    n = 20_000 if engine == "k2" else 200_000
    chars, off = W.example5_strings(n, 64, 4096, 9, "cuda")
    m = _matcher(t, engine)
    s = torch.cuda.current_stream().cuda_stream
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s)
    torch.cuda.synchronize()
    assert m.overflow_count() == 0
    frac = float(out.float().mean())
    assert 0.3 < frac < 0.7
    # split
    h = n // 3
    out2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    m.match_ptrs(chars.data_ptr(), off.data_ptr(), h, out2.data_ptr(), s)
    m.match_ptrs(chars.data_ptr(), off[h:].contiguous().data_ptr(), n - h, out2.data_ptr() + h, s)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    # permutation of a slice (gathered on the host: 5000 strings)
    k = 5000
    off_h = off[:k + 1].cpu().numpy().astype(np.uint64)
    chars_h = chars[:int(off_h[-1])].cpu().numpy()
    strings = [bytes(chars_h[int(off_h[i]):int(off_h[i + 1])]) for i in range(k)]
    perm = np.random.default_rng(1).permutation(k)
    pc, po = H.make_batch([strings[i] for i in perm])
    got_p = m.match_host(pc, po)
    base = out[:k].cpu().numpy()
    assert np.array_equal(got_p, base[perm])
    # oracle on the slice, and the raw-text route on the same strings
    assert np.array_equal(base, H.oracle_bits(t, chars_h, off_h))
    assert np.array_equal(m.match_text_host(b"\n".join(strings)), base)
    m.close()


@pytest.mark.parametrize("name", ["nfa_config2", "ex05_fwd", "ex02_rev", "nfa_blowup", "nfa_huge"])
def test_sharded_1_2_4_8_ways_gives_the_identical_bit_vector(name):
    """the batch cut into 1, 2, 4 and 8 byte-balanced contiguous shards
    (re2-modification_b200/sharding.py -- what each rank of an N-GPU job owns), every shard matched
    on its own with rebased offsets, concatenated: the same bits as the whole batch."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(H.PKG, "sharding.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)
    t, strings, bits = load_case(name)
    rng = np.random.default_rng(21)
    extra = [bytes(rng.choice(np.frombuffer(b"aabbc", dtype=np.uint8), size=int(L))) for L in rng.integers(0, 400, size=3000)]
    chars, off = H.make_batch(strings + extra)
    m = rxm.Matcher(t, 0)
    whole = m.match_host(chars, off)
    assert np.array_equal(whole[:len(bits)], bits)
    for world in (2, 4, 8):
        bounds = S.shard_by_bytes(off, world)
        parts = []
        for r in range(world):
            c, o = S.local_view(chars, off, bounds[r], bounds[r + 1])
            parts.append(m.match_host(c, o))
        assert np.array_equal(np.concatenate(parts), whole), world
    m.close()


def test_random_expression_corpus_on_device():
    """The 280 random expressions of tests/golden/fuzz (bits from the reference's own code) through
    every device engine that can take them: the planner's choice, K2 and K3 for the MFAs, the
    bit-set engine in both forms for the memory-free ones, and the raw-text route."""
    from cases import load_fuzz_corpus
    corpus = load_fuzz_corpus()
    engines_seen = set()
    for regex, flags, kind, t, strings, bits in corpus:
        chars, off = H.make_batch(strings)
        tag = (regex, flags)
        variants = [{}]
        if kind == "mfa" and t.c.n_cells > 4:
            variants += [{"engine": "k2"}]  # more than 4 cells: the planner's choice is K2 as well
        elif kind == "mfa":
            variants += [{"engine": "k2"}, {"engine": "k3"}, {"engine": "k4"}]
        else:
            variants += [{"engine": "bitset"}, {"engine": "bitset", "flags": rxm.OPT_K1B_WALK}]
        for env in variants:
            m = rxm.Matcher(t, 0, **env)
            engines_seen.add((rxm.ENGINE_NAMES[m.plan().engine], tuple(sorted(env.items()))))
            got = m.match_host(chars, off)
            assert np.array_equal(got, bits), (tag, env, [strings[i] for i in np.nonzero(got != bits)[0][:3]])
            if not env:
                assert np.array_equal(m.match_text_host(b" ".join(strings)), bits), ("text", tag)
            assert m.overflow_count() == 0
            m.close()
    names = {e for e, _ in engines_seen}
    assert {"K1_DFA", "K1_BITSET", "K2_THREAD", "K3_WARP", "K4_THREAD"} <= names


def test_match_text_random_byte_soup():
    """Tokeniser fuzz: texts of random bytes (letters, every whitespace kind in runs of random
    length, NUL, high bytes, the word `exit` sprinkled in) of 0 .. 300 KB against bytes.split() +
    the oracle, at every alignment of the device pointer mod 16."""
    import torch
    t, _, _ = load_case("nfa_config2")
    m = rxm.Matcher(t, 0)
    rng = np.random.default_rng(77)
    pool = [b"a", b"b", b"a", b"b", b"ab", b"aaba", b" ", b"\n", b"\t", b"\r", b"\v", b"\f", b"  ", b"\n\n\n",
            b"\x00", b"\xa0", b"\x85", b"\x1f", b"z", b"exi", b"xit"]
    for rounds, (lo, hi, p_exit) in enumerate([(0, 40, 0.0), (0, 40, 0.2), (100, 3000, 0.0), (100, 3000, 0.01),
                                                (20000, 300000, 0.0), (20000, 300000, 0.00002)] * 3):
        k = int(rng.integers(lo, hi + 1))
        idx = rng.integers(0, len(pool), size=k)
        parts = [pool[i] for i in idx]
        if p_exit:
            for j in np.nonzero(rng.random(k) < p_exit)[0]:
                parts[j] = b" exit "
        text = b"".join(parts)
        toks = _ref_tokens(text)
        got = m.match_text_host(text)
        assert len(got) == len(toks), (rounds, len(got), len(toks))
        assert m.saw_exit.value == (1 if b"exit" in text.split() else 0)
        if toks:
            chars, off = H.make_batch(toks)
            assert np.array_equal(got, H.oracle_bits(t, chars, off)), rounds
        # device pointers at a random alignment
        shift = int(rng.integers(0, 16))
        d = torch.cat([torch.full((shift,), 32, dtype=torch.uint8), torch.frombuffer(bytearray(text) or bytearray(b" "), dtype=torch.uint8)]).cuda()
        out = torch.empty(max(1, len(toks)), dtype=torch.uint8, device="cuda")
        n = m.match_text_ptrs(d.data_ptr() + shift, len(text), out.data_ptr(), out.numel())
        torch.cuda.synchronize()
        assert n == len(toks)
        assert np.array_equal(out[:n].cpu().numpy(), got)
    m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ex02_fwd", "ex02_rev", "ex05_fwd"])
def test_set_concurrency_changes_the_grid_not_the_bits(name):
    """rxm_set_concurrency: handles sharing the device launch smaller persistent grids (down to a
    single block); results are identical, and two handles on streams of their own agree too."""
    import torch
    t, strings, bits = load_case(name)
    big = strings * 12  # several waves of strings for a one-block grid
    chars, off = H.make_batch(big)
    want = np.tile(bits, 12)
    with pytest.raises(H.rxm.RxmError):
        H.rxm.Matcher(t, 0).set_concurrency(0)
    for share in (1, 2, 7, 100000):
        m = H.rxm.Matcher(t, 0)
        m.set_concurrency(share)
        got = m.match_host(chars, off)
        assert np.array_equal(got, want), share
        m.close()
    dev = torch.device("cuda:0")
    dc, do = torch.from_numpy(chars).to(dev), torch.from_numpy(off.astype(np.int64)).to(dev)
    ms = [H.rxm.Matcher(t, 0) for _ in range(2)]
    outs = [torch.empty(len(big), dtype=torch.uint8, device=dev) for _ in ms]
    sts = [torch.cuda.Stream(device=dev) for _ in ms]
    torch.cuda.synchronize()
    for m, o, st in zip(ms, outs, sts):
        m.set_concurrency(2)
        m.match_ptrs(dc.data_ptr(), do.data_ptr(), len(big), o.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        assert np.array_equal(o.cpu().numpy(), want)


# ---- BASELINE configs 4 and 5 as bench.py builds them ---------------------------


@pytest.mark.parametrize("engine", [None, "k3", "k2"])
@pytest.mark.parametrize("name", ["ex02_fwd", "ex02_rev"])
def test_config4_attack_batch_every_bit(name, engine):
    """BASELINE config 4: example 2's attack strings (bbaa)^k aaba (bbaa)^k [c] (test/example_2/pump.txt,
    generator matchers/example_runner.cpp:15-29) of 435 .. 65 536 letters, forward and -reverse tables,
    1024 strings (256 for the thread-per-string K2): EVERY bit against the C restatement, through the
    planner's choice and the other MFA engines; device pointers (the long-string hand-out of rxm_api.cu)."""
    import torch
    W = H.load_workloads()
    t, _, _ = load_case(name)
    n = 256 if engine == "k2" else 1024
    chars, off = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", n, 435, 65536, 4004)
    assert int(np.diff(off).max()) > 60000
    want = H.oracle_bits(t, chars, off)
    m = _matcher(t, engine)
    d_chars = torch.from_numpy(chars).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    d_out = torch.full((n,), 7, dtype=torch.uint8, device="cuda")
    m.match_ptrs(d_chars.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    assert m.overflow_count() == 0
    assert np.array_equal(got, want), [int(off[i + 1] - off[i]) for i in np.nonzero(got != want)[0][:8]]
    if name == "ex02_fwd":
        assert 0 < int(want.sum()) < n
    assert np.array_equal(m.match_host(chars, off), want)  # host buffers: the staged route
    m.close()


@pytest.mark.skipif(not H.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,flags", [("ex02_fwd", []), ("ex02_rev", ["-reverse"])])
def test_config4_longest_strings_against_the_reference_itself(name, flags):
    """The longest attack strings the reference's own code (oracle/_ref/diploma_ref_bump) finishes in the
    time limit -- it is O(len^2) per string, mfa.cpp:136,203 copy the input per call: 16 strings up to
    16 K letters -- against the device."""
    W = H.load_workloads()
    t, _, _ = load_case(name)
    chars, off = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", 16, 4000, 16384, 4005, log_uniform=False)
    ref = H.reference_bits(W.README_EXAMPLES[2][0], flags, chars, off, timeout=900)
    m = rxm.Matcher(t, 0)
    got = m.match_host(chars, off)
    m.close()
    assert np.array_equal(got, ref)
    assert np.array_equal(H.oracle_bits(t, chars, off), ref)


def test_config5_mixed_example_batches_on_their_own_streams():
    """BASELINE config 5 as bench.py runs it: the ten README examples (README.md:78-89), 20 000 mixed
    strings each (pumped / near-miss / random, 16-512 letters), ten matchers on ten streams at once;
    every bit against the C restatement."""
    import torch
    W = H.load_workloads()
    dev = torch.device("cuda:0")
    jobs = []
    for ex in range(1, 11):
        t, _, _ = load_case(f"ex{ex:02d}_fwd")
        chars, off = W.mixed_example_batch(ex, 20000, 1000 * ex)
        jobs.append({"t": t, "chars": chars, "off": off, "m": rxm.Matcher(t, 0),
                     "d_chars": torch.from_numpy(chars).to(dev), "d_off": torch.from_numpy(off.astype(np.int64)).to(dev),
                     "d_out": torch.full((20000,), 7, dtype=torch.uint8, device=dev), "st": torch.cuda.Stream(device=dev)})
    torch.cuda.synchronize()
    for j in jobs:
        j["m"].match_ptrs(j["d_chars"].data_ptr(), j["d_off"].data_ptr(), 20000, j["d_out"].data_ptr(), j["st"].cuda_stream)
    torch.cuda.synchronize()
    engines = set()
    for ex, j in enumerate(jobs, 1):
        want = H.oracle_bits(j["t"], j["chars"], j["off"])
        got = j["d_out"].cpu().numpy()
        assert j["m"].overflow_count() == 0
        assert np.array_equal(got, want), (ex, int((got != want).sum()))
        engines.add(rxm.ENGINE_NAMES[j["m"].plan().engine])
        j["m"].close()
    assert engines == {"K4_THREAD"}


@pytest.mark.parametrize("name", ["ex02_rev", "ex08_rev", "ex15_rev", "ex05_rev"])
def test_k4_hands_outgrown_strings_to_k3_on_the_device(name):
    """Automata with more nodes than a K4 thread has slots: a batch large enough for the tile-sorted
    hand-out (20 000 strings), the strings K4 cannot hold are run by K3's list mode in the same call;
    all bits against the C restatement."""
    t, strings, bits = load_case(name)
    rng = np.random.default_rng(31)
    extra = [bytes(rng.choice(np.frombuffer(b"aaabbc", dtype=np.uint8), size=int(L))) for L in rng.integers(0, 120, size=20000)]
    chars, off = H.make_batch(strings + extra)
    m = rxm.Matcher(t, 0)
    # (the planner sends automata above 40 nodes -- example 8 reversed has 77 -- to K3 as a whole)
    assert rxm.ENGINE_NAMES[m.plan().engine] == ("K3_WARP" if t.c.n_states > 40 else "K4_THREAD")
    got = m.match_host(chars, off)
    assert m.overflow_count() == 0
    assert np.array_equal(got[:len(bits)], bits)
    assert np.array_equal(got, H.oracle_bits(t, chars, off))
    m.close()
