"""CPU tier, only where the reference build exists (oracle/_ref + bin/rxm_compile, i.e. in the
build container; skipped on a box without them): the pieces that were generated FROM the reference
still agree with it -- tables from the front end equal the committed fixtures, and the oracle
equals the reference's own code on fresh random strings."""
import os

import numpy as np
import pytest

import helpers as H
from cases import BY_NAME, CASE_NAMES, load_case

pytestmark = pytest.mark.skipif(not H.have_reference(), reason="oracle/_ref or bin/rxm_compile not built")


@pytest.mark.parametrize("name", CASE_NAMES)
def test_front_end_reproduces_committed_tables(name):
    m = BY_NAME[name]
    assert H.compile_tables_text(m["regex"], m["flags"]) == load_case(name)[0].text


@pytest.mark.parametrize("name", ["ex01_fwd", "ex02_rev", "ex05_rev", "ex08_rev", "ex14_rev",
                                  "ex15_rev", "nfa_config2", "nfa_quirk", "nfa_dots"])
def test_oracle_equals_reference_on_fresh_strings(name):
    m = BY_NAME[name]
    t, _, _ = load_case(name)
    rng = np.random.default_rng(abs(hash(name)) % 10000)
    strings = []
    for _ in range(400):
        L = int(rng.integers(1, 60))
        alpha = b"ab" if rng.random() < 0.6 else b"aaab"
        strings.append(bytes(rng.choice(np.frombuffer(alpha, dtype=np.uint8), size=L)))
    chars, off = H.make_batch(strings)
    ref = H.reference_bits(m["regex"], m["flags"], chars, off, binary=H.REF_BUMP)
    assert np.array_equal(H.oracle_bits(t, chars, off), ref)
