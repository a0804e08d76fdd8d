"""CPU tier for K4 (MFA, one thread per string): its per-string core (rxm_k4_core.cuh) compiled for the
host, and the kernel SOURCE (rxm_k4.cu, with its launch function and the hand-over of outgrown strings
to K3) on the SIMT emulator -- against the golden vectors (bits of the reference's own code), the
random-expression corpus, random tables and config-3 / config-4 / config-5 strings with the C
restatement (oracle/) as the checker."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from cases import BY_NAME, CASE_NAMES, load_case, load_fuzz_corpus
from conftest import hostsim_lib_path

MFA_CASES = [n for n in CASE_NAMES if BY_NAME[n]["kind"] == "mfa"]
_hs = None


def hostsim():
    global _hs
    if _hs is None:
        L = C.CDLL(hostsim_lib_path())
        L.hostsim_k4core_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                           C.c_void_p]
        L.hostsim_k4core_batch.restype = C.c_int
        L.hostsim_k4_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32,
                                       C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint32,
                                       C.c_uint64]
        L.hostsim_k4_batch.restype = C.c_int
        L.hostsim_span_equal.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.hostsim_span_equal.restype = C.c_int
        _hs = L
    return _hs


def _pad(chars, lead=16):
    pad = np.zeros(len(chars) + lead + 16, dtype=np.uint8)  # block compares read whole aligned 8-byte words
    pad[lead:lead + len(chars)] = chars
    return pad


def k4_core(t, chars, off, maxl=0):
    """-> (rc, bits (2 = the string met a limit), [steps run in full, repeated steps answered, strings at a limit])"""
    pad = _pad(chars)
    got = np.full(len(off) - 1, 7, dtype=np.uint8)
    info = (C.c_uint64 * 3)()
    rc = hostsim().hostsim_k4core_batch(C.cast(t.ptr, C.c_void_p), pad.ctypes.data + 16, off.ctypes.data, len(off) - 1,
                                        got.ctypes.data, maxl, info)
    return rc, got, list(info)


def k4_emulated(t, strings, maxl=0, order=None, seed=0, limit=200_000_000):
    """rxm_k4.cu on the SIMT emulator -> (rc, bits, overflow, strings handed on to K3, report)"""
    chars, off = H.make_batch(strings)
    pad = _pad(chars)
    got = np.full(len(strings), 7, dtype=np.uint8)
    msg, ovf, redo = C.create_string_buffer(600), C.c_ulonglong(0), C.c_ulonglong(0)
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.uint32)
    rc = hostsim().hostsim_k4_batch(C.cast(t.ptr, C.c_void_p), pad.ctypes.data + 16, off.ctypes.data, len(strings),
                                    got.ctypes.data, maxl, order.ctypes.data if order is not None else None, limit,
                                    C.byref(ovf), C.byref(redo), msg, 600, seed)
    return rc, got, ovf.value, redo.value, msg.value.decode()


def test_block_compare_every_alignment_and_length():
    """k4_span_equal: 8 bytes per iteration from aligned words; every pair of alignments, lengths 0..40,
    equal spans and spans that differ in exactly one byte (first, last, middle); only words that hold a
    byte of a span may be read -- the spans sit at the very ends of a page-aligned buffer."""
    hs = hostsim()
    rng = np.random.default_rng(1)
    buf = np.zeros(8192 + 64, dtype=np.uint8)
    base = (-buf.ctypes.data) % 4096  # buf[base : base + 4096] is one page
    page = buf[base:base + 4096]
    for L in list(range(0, 41)) + [63, 64, 65, 127, 200]:
        for oa in range(8):
            for ob in range(8):
                page[:] = rng.integers(0, 256, size=4096, dtype=np.uint8)
                a0, b0 = oa, 4096 - L - ob if ob else 4096 - L  # a at the start of the page, b flush with its end
                if ob:
                    b0 = 4096 - L - (8 - ob) if L + (8 - ob) <= 2048 else 2048 + ob
                b0 = max(b0, 2048)
                text = rng.integers(97, 100, size=L, dtype=np.uint8)
                page[a0:a0 + L] = text
                page[b0:b0 + L] = text
                pa, pb = page.ctypes.data + a0, page.ctypes.data + b0
                assert hs.hostsim_span_equal(pa, pb, L) == 1, (L, oa, ob)
                assert hs.hostsim_span_equal(pb, pa, L) == 1, (L, oa, ob)
                for k in sorted({0, L - 1, L // 2} - {-1}):
                    if L == 0:
                        break
                    page[b0 + k] ^= 0x10
                    assert hs.hostsim_span_equal(pa, pb, L) == 0, (L, oa, ob, k)
                    assert hs.hostsim_span_equal(pb, pa, L) == 0, (L, oa, ob, k)
                    page[b0 + k] ^= 0x10
                # bytes just outside the spans never matter
                if a0 > 0:
                    page[a0 - 1] ^= 0xff
                page[a0 + L] ^= 0xff
                if b0 + L < 4096:
                    page[b0 + L] ^= 0xff
                page[b0 - 1] ^= 0xff
                page[b0:b0 + L] = text
                page[a0:a0 + L] = text
                assert hs.hostsim_span_equal(pa, pb, L) == 1, (L, oa, ob, "outside")


@pytest.mark.parametrize("name", MFA_CASES)
def test_k4_core_on_host_matches_golden(name):
    t, strings, bits = load_case(name)
    chars, off = H.make_batch(strings)
    rc, got, info = k4_core(t, chars, off)  # the planner's pool: 8 .. 16 slots
    assert rc == 0
    ok = got != 2  # strings that outgrow the pool are REPORTED (the kernel hands them to K3), never answered wrongly
    assert np.array_equal(got[ok], bits[ok]), [strings[i] for i in np.nonzero((got != bits) & ok)[0][:5]]
    assert int((~ok).sum()) == info[2]
    if name.endswith("_fwd"):
        assert info[2] == 0  # every forward automaton of the reference's examples fits
    rc, got, info = k4_core(t, chars, off, 32)
    assert rc == 0 and (info[2] == 0 or name == "ex08_rev")
    ok = got != 2
    assert np.array_equal(got[ok], bits[ok])


def test_k4_core_random_expression_corpus_and_replay_share():
    n_run = replayed = limited = 0
    for regex, flags, kind, t, strings, bits in load_fuzz_corpus():
        if kind != "mfa" or t.c.n_cells > 4:
            continue
        chars, off = H.make_batch(strings)
        rc, got, info = k4_core(t, chars, off)
        assert rc == 0, (regex, flags)
        ok = got != 2
        assert np.array_equal(got[ok], bits[ok]), (regex, flags)
        limited += info[2]
        n_run += 1
        replayed += info[1]
    assert n_run > 100 and replayed > 300 and limited < 300


def test_k4_core_small_sets_report_what_they_cannot_hold():
    """maxl below the automaton's need: a string is either answered like the reference or reported (2) --
    never answered wrongly; the kernel hands the reported ones to K3."""
    reported = 0
    for name in ("ex02_rev", "ex05_rev", "ex08_rev", "ex14_rev", "ex15_rev", "ex09_fwd"):
        t, strings, bits = load_case(name)
        chars, off = H.make_batch(strings)
        for maxl in (1, 2, 4):
            rc, got, info = k4_core(t, chars, off, maxl)
            assert rc == 0
            ok = got != 2
            assert np.array_equal(got[ok], bits[ok]), (name, maxl)
            assert int((~ok).sum()) == info[2]
            reported += info[2]
    assert reported > 100


def test_k4_core_config3_config4_config5_strings():
    """The bench workloads at reduced size, the C restatement as the checker; on config 3 more than nine
    steps in ten are answered by their block compares alone."""
    W = H.load_workloads()
    t, _, _ = load_case("ex05_fwd")
    chars, off = W.example5_strings(2000, 64, 4096, 13, "cpu")
    chars, off = chars.numpy(), off.numpy().astype(np.uint64)
    rc, got, info = k4_core(t, chars, off)
    assert rc == 0 and np.array_equal(got, H.oracle_bits(t, chars, off))
    assert info[1] > 0.9 * (info[0] + info[1]), info
    assert 0.3 < got.mean() < 0.7
    for name in ("ex02_fwd", "ex02_rev"):
        t, _, _ = load_case(name)
        chars, off = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", 96, 435, 65536, 5)
        rc, got, info = k4_core(t, chars, off)
        assert rc == 0
        ok = got != 2
        assert np.array_equal(got[ok], H.oracle_bits(t, chars, off)[ok]), name
        assert name == "ex02_rev" or info[2] == 0
    for ex in range(1, 11):
        t, _, _ = load_case(f"ex{ex:02d}_fwd")
        chars, off = W.mixed_example_batch(ex, 1500, 77 + ex)
        rc, got, info = k4_core(t, chars, off)
        assert rc == 0 and info[2] == 0
        assert np.array_equal(got, H.oracle_bits(t, chars, off)), ex


def test_k4_core_random_tables():
    import importlib.util, random
    spec = importlib.util.spec_from_file_location("fuzz_tables", os.path.join(H.ROOT, "tests", "fuzz", "fuzz_tables_gpu.py"))
    F = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(F)
    rng, nprng = random.Random(11), np.random.default_rng(11)
    n_run = n_emul = 0
    for k in range(400):
        text = F.random_table(rng, True)
        t = H.rxm.Tables(text)
        if t.c.n_cells > 4:
            continue
        strings = [b"", b"a", b"b", b"ab", b"ba", b"aab", b"abab", b"cc"]
        strings += [bytes(nprng.choice(np.frombuffer(b"aaabbc1", dtype=np.uint8), size=int(L))) for L in nprng.integers(0, 40, size=120)]
        strings += [bytes(nprng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=int(nprng.integers(1, 6)))) * int(nprng.integers(2, 60))
                    for _ in range(8)]
        chars, off = H.make_batch(strings)
        want = H.oracle_bits(t, chars, off)
        rc, got, info = k4_core(t, chars, off)
        if rc == H.rxm.RXM_ERR_UNSUPPORTED:
            continue  # edge programs too large: the planner keeps the table on K2
        assert rc == 0, text
        ok = got != 2  # a (node, cells) pair the host analysis did not reach is reported, as in K3
        assert np.array_equal(got[ok], want[ok]), text
        n_run += 1
        if k % 8 == 0:
            rc, gote, ovf, redo, msg = k4_emulated(t, strings, maxl=(0, 2)[n_emul % 2], seed=k)
            assert rc == 0, (msg, text)
            if ovf == 0:
                assert np.array_equal(gote, want), ("K4 emulated", text)
            n_emul += 1
    assert n_run > 150 and n_emul > 20


@pytest.mark.parametrize("name", MFA_CASES)
def test_k4_kernel_source_on_the_simt_emulator_matches_golden(name):
    """rxm_k4.cu -- kernel and launch function -- compiled for the host: 128-thread blocks, tickets taken
    64 at a time per warp, per-thread sets in (emulated) shared memory; automata with more nodes than a
    thread has slots go through the hand-over to K3 (rxm_k3.cu's list mode) as well."""
    t, strings, bits = load_case(name)
    rc, got, ovf, redo, msg = k4_emulated(t, strings, seed=5)
    assert rc == 0 and ovf == 0, msg
    assert np.array_equal(got, bits), [strings[i] for i in np.nonzero(got != bits)[0][:5]]


@pytest.mark.parametrize("name", ["ex02_rev", "ex05_rev", "ex08_rev", "ex15_rev"])
def test_k4_emulated_hands_outgrown_strings_to_k3(name):
    t, strings, bits = load_case(name)
    rc, got, ovf, redo, msg = k4_emulated(t, strings, maxl=2)
    assert rc == 0 and ovf == 0, msg
    assert redo > 0
    assert np.array_equal(got, bits)


@pytest.mark.parametrize("name", ["ex02_fwd", "ex05_fwd", "ex08_rev"])
def test_k4_emulated_hand_out_order_and_ragged_batches(name):
    t, strings, bits = load_case(name)
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(strings)).astype(np.uint32)
    rc, got, ovf, redo, msg = k4_emulated(t, strings, order=perm, seed=2)
    assert rc == 0 and ovf == 0, msg
    assert np.array_equal(got, bits)
    for n in (0, 1, 2, 31, 33, 65, 129):
        rc, got, ovf, redo, msg = k4_emulated(t, strings[:n])
        assert rc == 0, msg
        assert np.array_equal(got, bits[:n])
    sub = [b"", strings[1], b"", b"", strings[2], b""]
    chars, off = H.make_batch(sub)
    rc, got, ovf, redo, msg = k4_emulated(t, sub)
    assert rc == 0, msg
    assert np.array_equal(got, H.oracle_bits(t, chars, off))


def test_k4_emulated_on_config3_strings():
    W = H.load_workloads()
    t, _, _ = load_case("ex05_fwd")
    chars, off = W.example5_strings(300, 64, 4096, 11, "cpu")
    chars, off = chars.numpy(), off.numpy().astype(np.uint64)
    strings = [bytes(chars[int(off[i]):int(off[i + 1])]) for i in range(len(off) - 1)]
    want = H.oracle_bits(t, chars, off)
    rc, got, ovf, redo, msg = k4_emulated(t, strings, seed=1)
    assert rc == 0 and ovf == 0 and redo == 0, msg
    assert np.array_equal(got, want)
    assert 0 < int(want.sum()) < len(strings)


def test_warp_cooperative_periodicity_check_every_alignment_and_distance():
    """rxm_k4.cu: k4_coop_verify -- phase A's check `s[j] == s[j - delta] for j in [vp, vcap)` done by a whole
    warp, 16-byte vectors, the stream delta bytes back picked with a loop-invariant word offset and funnel
    shifts -- on the SIMT emulator against numpy: every start alignment mod 16, distances 1..70 and a few
    thousand, differences planted before, at, and after the window's ends, windows shorter than a vector
    and longer than a round."""
    rng = np.random.default_rng(31)
    chunks, beg, ln, dl, vps, vcs, want = [], [], [], [], [], [], []
    pos = 0
    lead = 16
    for case in range(1500):
        delta = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 70, 1000, 4097])) if case % 3 else int(rng.integers(1, 71))
        reps = int(rng.integers(1, 60)) if delta < 100 else int(rng.integers(1, 3))
        n = delta * (reps + 1) + int(rng.integers(0, delta + 1))
        unit = rng.integers(97, 100, size=delta).astype(np.uint8)
        s = np.resize(unit, n).copy()
        for _ in range(int(rng.integers(0, 3))):  # planted differences
            s[int(rng.integers(0, n))] = 122
        vp = int(rng.integers(delta, n + 1))
        vcap = int(rng.integers(vp, n + 1))
        j = np.arange(vp, vcap)
        bad = j[s[j] != s[j - delta]]
        want.append(int(bad[0]) if len(bad) else vcap)
        shift = int(rng.integers(0, 16))  # every alignment of the string's first byte
        chunks.append(np.full(shift, 35, dtype=np.uint8))
        pos += shift
        beg.append(pos)
        chunks.append(s)
        pos += n
        ln.append(n); dl.append(delta); vps.append(vp); vcs.append(vcap)
    chars = np.concatenate([np.full(lead, 36, dtype=np.uint8)] + chunks + [np.full(32, 37, dtype=np.uint8)])
    beg = np.asarray(beg, dtype=np.uint64)
    arrs = [np.asarray(a, dtype=np.uint32) for a in (ln, dl, vps, vcs)]
    res = np.zeros(len(want), dtype=np.uint32)
    L = hostsim()
    L.hostsim_coop_verify.argtypes = [C.c_void_p] * 6 + [C.c_uint32, C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint64]
    L.hostsim_coop_verify.restype = C.c_int
    for seed in (0, 9):
        msg = C.create_string_buffer(600)
        rc = L.hostsim_coop_verify(chars.ctypes.data + lead, beg.ctypes.data, arrs[0].ctypes.data, arrs[1].ctypes.data,
                                   arrs[2].ctypes.data, arrs[3].ctypes.data, len(want), res.ctypes.data, msg, 600, seed)
        assert rc == 0, msg.value.decode()
        diff = np.nonzero(res != np.asarray(want, dtype=np.uint32))[0]
        assert len(diff) == 0, [(int(k), int(res[k]), want[k], dl[k], vps[k], vcs[k], ln[k]) for k in diff[:5]]


@pytest.mark.skipif(not os.path.exists(H.RXM_COMPILE), reason="bin/rxm_compile not built")
@pytest.mark.parametrize("regex", ["{(a|b|c|d|e|f|g|h|i|j|k|l)*}:1m&1(n|&1)*", "({(ab|cd|ef|gh|ij)*}:1x|y&1)*z&1"])
def test_k4_leaf_lists_by_letter_class_with_more_literals_than_classes(regex):
    """rxm_plan.cpp: K4's leaf lists are split by the input letter's class, at most kProgMaxClasses - 1 literals get a
    class of their own and the rest share class 0 with the bytes no edge carries -- an expression with 13 distinct
    literals on the host core and on the emulated kernel against the C restatement; every list is a subsequence of
    the key's undivided leaf list."""
    t = H.rxm.Tables(H.compile_tables_text(regex))
    rng = np.random.default_rng(3)
    letters = sorted({c for c in regex.encode() if chr(c).isalpha()})
    alpha = np.frombuffer(bytes(letters) + b"q1", dtype=np.uint8)
    strings = [bytes(rng.choice(alpha, size=int(n))) for n in rng.integers(0, 40, size=1500)]
    for _ in range(300):  # strings inside the language: a block, its separator, the block again
        blk = bytes(rng.choice(alpha[:10], size=int(rng.integers(0, 12))))
        strings.append(blk + bytes([letters[-2] if b"m" in regex.encode() else ord("x")]) + blk + blk * int(rng.integers(0, 3)))
    chars, off = H.make_batch(strings)
    off = off.astype(np.uint64)
    want = H.oracle_bits(t, chars, off)
    assert 0 < int(want.sum()) < len(want)
    rc, got, info = k4_core(t, chars, off)
    assert rc == 0 and np.array_equal(got, want), int((got != want).sum())
    rc, got, ovf, redo, msg = k4_emulated(t, strings, seed=2)
    assert rc == 0, msg
    assert np.array_equal(np.where(got == 2, want, got), want) and ovf == 0
