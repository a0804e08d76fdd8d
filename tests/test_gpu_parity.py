"""GPU tier: the CUDA path, called through the C ABI, against the golden vectors
(bits produced by running the reference) and against the oracle on seeded batches.
Bit-exact: this is byte/integer work."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from cases import BY_NAME, CASE_NAMES, load_case

rxm = H.rxm
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASE_NAMES)
def test_golden_host_buffers(name):
    t, strings, bits = load_case(name)
    chars, off = H.make_batch(strings)
    m = rxm.Matcher(t, 0)
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


@pytest.mark.parametrize("name", ["nfa_config2", "nfa_quirk", "nfa_abb", "nfa_dots", "nfa_third"])
def test_nfa_random_batches_vs_oracle(name):
    t, _, _ = load_case(name)
    rng = np.random.default_rng(123)
    m = rxm.Matcher(t, 0)
    for (n, lo, hi) in ((5000, 0, 40), (3000, 1, 300), (257, 1000, 5000)):
        chars, off = _random_batch(rng, n, lo, hi, b"ab")
        assert np.array_equal(m.match_host(chars, off), H.oracle_bits(t, chars, off))
    m.close()


@pytest.mark.parametrize("name", ["ex01_fwd", "ex02_fwd", "ex02_rev", "ex05_fwd", "ex05_rev",
                                  "ex08_rev", "ex09_fwd", "ex14_rev", "ex15_rev", "ex17_rev"])
def test_mfa_random_batches_vs_oracle(name):
    t, _, _ = load_case(name)
    rng = np.random.default_rng(7)
    m = rxm.Matcher(t, 0)
    for (n, lo, hi, alpha) in ((4000, 0, 30, b"ab"), (2000, 1, 60, b"aaab"), (1000, 1, 40, b"abc"),
                               (300, 100, 400, b"aaaaab")):
        chars, off = _random_batch(rng, n, lo, hi, alpha)
        got = m.match_host(chars, off)
        want = H.oracle_bits(t, chars, off)
        assert np.array_equal(got, want), int((got != want).sum())
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
    assert rxm.ENGINE_NAMES[m.plan().engine] in ("K2_THREAD", "K3_WARP")
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
