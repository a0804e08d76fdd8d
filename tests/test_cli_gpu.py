"""GPU tier: the `-match` driver whose simulation runs on the device
(re2-modification_b200/bin/diploma_rxm) against the reference's own driver protocol:
same stdout (banners + one 0/1 line per token) as oracle/_ref/diploma_ref_bump where that
binary exists, and the golden bits in any case.  Both binaries are prebuilt (they embed the
reference front end); nothing here reads /root/reference."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import helpers as H
from cases import BY_NAME, load_case

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(H.PKG, "bin", "diploma_rxm")


def _run(binary, flags, regex, tokens, cwd):
    text = regex + "\n" + "\n".join(tokens) + "\nexit\n"
    r = subprocess.run([binary, "-match", *flags], input=text.encode(), capture_output=True,
                       cwd=cwd, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-400:]
    return r.stdout.decode()


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="diploma_rxm not built")
@pytest.mark.parametrize("name", ["nfa_config2", "ex01_fwd", "ex02_rev", "ex05_fwd", "ex10_rev"])
def test_match_driver_stdin_protocol(name):
    m = BY_NAME[name]
    _, strings, bits = load_case(name)
    keep = [i for i, s in enumerate(strings) if s and s != b"exit"]  # `cin >>` cannot carry ""
    tokens = [strings[i].decode() for i in keep]
    with tempfile.TemporaryDirectory() as td:
        out = _run(DRIVER, m["flags"], m["regex"], tokens, td)
        got = [ln for ln in out.splitlines() if ln in ("0", "1")]
        assert [int(x) for x in got] == [int(bits[i]) for i in keep]
        if os.path.exists(H.REF_BUMP):
            ref = _run(H.REF_BUMP, m["flags"], m["regex"], tokens, td)
            assert out == ref  # banners included


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="diploma_rxm not built")
def test_match_driver_batch_route():
    m = BY_NAME["ex05_fwd"]
    _, strings, bits = load_case("ex05_fwd")
    chars, off = H.make_batch(strings)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.rxmb"), os.path.join(td, "out.bits")
        H.write_batch_file(fin, chars, off)
        r = subprocess.run([DRIVER, "-match", "-regex", m["regex"], "-batch", fin, fout],
                           capture_output=True, cwd=td, timeout=300)
        assert r.returncode == 0, r.stderr.decode()[-400:]
        assert np.array_equal(np.fromfile(fout, dtype=np.uint8), bits)


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="diploma_rxm not built")
def test_match_driver_pieces_eof_and_tokens_after_exit():
    """Raw-text route: several tokens per line, input cut into 64 KB pieces at whitespace
    (a token longer than a piece included), end of input without `exit`, and tokens after
    `exit` ignored (match.cpp:24)."""
    m = BY_NAME["nfa_config2"]
    t, _, _ = load_case("nfa_config2")
    rng = np.random.default_rng(2)
    ab = np.frombuffer(b"ab", dtype=np.uint8)
    toks = [bytes(rng.choice(ab, size=int(L))) for L in rng.integers(1, 3000, size=400)]
    toks.insert(100, b"aaba" + b"a" * 200_000)  # longer than a piece
    chars, off = H.make_batch(toks)
    want = [int(x) for x in H.oracle_bits(t, chars, off)]
    body = b""
    for i, tk in enumerate(toks):
        body += tk + (b" " if i % 3 else b"\n")
    with tempfile.TemporaryDirectory() as td:
        for tail, flags in ((b"", []), (b"exit\nab aaba\n", []), (b"", ["-chunk", "65536"]),
                            (b"exit aaba", ["-chunk", "65536"])):
            r = subprocess.run([DRIVER, "-match", *flags], input=m["regex"].encode() + b"\n" + body + tail,
                               capture_output=True, cwd=td, timeout=300)
            assert r.returncode == 0, r.stderr.decode()[-400:]
            got = [int(ln) for ln in r.stdout.decode().splitlines() if ln in ("0", "1")]
            assert got == want, (tail, flags, len(got), len(want))


def _cumulative_lengths(pump, suffix, prefix, max_len):
    """Lengths of example_runner.cpp:118-145's strings (prefix grows by pumped + suffix; pump_size
    doubles each round and once more after every tenth)."""
    W = H.load_workloads()
    out, pump_size, count = [], 500, 0
    cur = prefix
    length = len(prefix) + pump_size + len(suffix)
    while length < max_len:
        cur = cur + W.pumped_string(pump_size, pump) + suffix
        length = len(cur)
        out.append(length)
        pump_size += pump_size
        count += 1
        if count % 10 == 0:
            pump_size *= 2
    return out


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="diploma_rxm not built")
def test_match_driver_benchmark_route():
    """`-match N` (example_runner.cpp:84-151): three result files of `len seconds` lines over the
    cumulative attack strings, here for example 5 (pump aa / suffix b / prefix aacaac)."""
    m = BY_NAME["ex05_fwd"]
    with tempfile.TemporaryDirectory() as td:
        ex = os.path.join(td, "test", "example_5")
        os.makedirs(ex)
        open(os.path.join(ex, "regexp.txt"), "w").write(m["regex"] + "\n")
        open(os.path.join(ex, "pump.txt"), "w").write("aa\nb\naacaac\n")
        r = subprocess.run([DRIVER, "-match", "5", "-maxlen", "200000"], capture_output=True, cwd=td, timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-400:]
        assert r.stdout.decode().splitlines()[0] == m["regex"]
        want = _cumulative_lengths(["aa"], "b", "aacaac", 200000)
        assert len(want) >= 5
        for name in ("diploma_results.txt", "diploma_bnf_results.txt", "diploma_reverse_results.txt"):
            rows = [ln.split() for ln in open(os.path.join(ex, name)).read().splitlines()]
            assert [int(a) for a, _ in rows] == want[:len(rows)], name
            assert len(rows) >= 3 and all(0 < float(b) < 1 for _, b in rows)


RUNNER_REF = os.path.join(H.REF_DIR, "diploma_ref_runner")


def _runner_golden(n):
    """tests/golden/runner/exNN.txt (tests/golden/make_runner_golden.py: the reference's own bits)."""
    meta, lines = {}, []
    for ln in open(os.path.join(H.GOLDEN, "runner", f"ex{n:02d}.txt")).read().splitlines():
        if ln.startswith("#"):
            continue
        key, _, rest = ln.partition(" ")
        if key in ("REGEX", "PUMP", "SUFFIX", "PREFIX"):
            meta[key] = rest
        elif key.startswith("RUNNER"):
            lines.append(ln)
    return meta, lines


@pytest.mark.skipif(not os.path.exists(DRIVER), reason="diploma_rxm not built")
@pytest.mark.parametrize("n", list(range(1, 11)))
def test_match_driver_benchmark_route_bits_equal_the_reference(n):
    """f4 (matchers/example_runner.cpp:84-151): `diploma_rxm -match N -bits FILE` -- the three automata
    (plain / -bnf / -reverse, :109-111) on the cumulative attack strings (:123) -- gives, string by string,
    the lengths AND the match bits the reference's own code gives for the same loop
    (oracle/ref_runner.cpp: the reference's pumped_string / split linked from example_runner.cpp, its
    parse / compile / MFA::match; committed as tests/golden/runner/), and the same again live where
    oracle/_ref/diploma_ref_runner exists.  README examples 1-10; the `len seconds` result files keep the
    reference's format."""
    meta, want = _runner_golden(n)
    assert len(want) >= 8
    with tempfile.TemporaryDirectory() as td:
        ex = os.path.join(td, "test", f"example_{n}")
        os.makedirs(ex)
        open(os.path.join(ex, "regexp.txt"), "w").write(meta["REGEX"] + "\n")
        open(os.path.join(ex, "pump.txt"), "w").write(meta["PUMP"] + "\n" + meta["SUFFIX"] + "\n" + meta["PREFIX"] + "\n")
        bits = os.path.join(td, "bits.txt")
        r = subprocess.run([DRIVER, "-match", str(n), "-maxlen", "16000", "-bits", bits], capture_output=True, cwd=td,
                           timeout=600)
        assert r.returncode == 0, r.stderr.decode()[-400:]
        got = open(bits).read().splitlines()
        assert got == want
        lens = [int(ln.split()[1]) for ln in want if ln.startswith("RUNNER ")]
        # the loop's last string is the first one NOT below -maxlen (example_runner.cpp:120 tests the length
        # of the string before): it is matched and timed, the bits files stop at -maxlen
        every = _cumulative_lengths(meta["PUMP"].split(","), meta["SUFFIX"], meta["PREFIX"], 16000)
        assert [x for x in every if x <= 16000] == lens
        for name in ("diploma_results.txt", "diploma_bnf_results.txt", "diploma_reverse_results.txt"):
            rows = [ln.split() for ln in open(os.path.join(ex, name)).read().splitlines()]
            assert [int(a) for a, _ in rows] == every[:len(rows)] and len(rows) >= 3, name
        if os.path.exists(RUNNER_REF):
            live = subprocess.run([RUNNER_REF, str(n), "-maxlen", "16000"], capture_output=True, cwd=td, timeout=900)
            assert live.returncode == 0
            assert [ln for ln in live.stdout.decode(errors="replace").splitlines() if ln.startswith("RUNNER")] == got
