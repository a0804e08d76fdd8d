"""Test helper (uses the oracle loader of tests/helpers.py): run a random-expression corpus --
the committed one or a larger one made by tests/golden/make_fuzz_corpus.py -- through every device
engine and list the cases that differ from the reference's bits.  python tests/fuzz/corpus_gpu.py [corpus.jsonl]"""
import os, sys
TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, TESTS)
import numpy as np
import helpers as H
from cases import load_fuzz_corpus
rxm = H.rxm
bad = 0
for regex, flags, kind, t, strings, bits in load_fuzz_corpus(sys.argv[1] if len(sys.argv) > 1 else None):
    chars, off = H.make_batch(strings)
    variants = [{}] + ([{"engine": "k2"}] if kind == "mfa" and t.c.n_cells > 4 else
                       [{"engine": "k2"}, {"engine": "k3"}, {"engine": "k4"}] if kind == "mfa" else
                       [{"engine": "bitset"}, {"engine": "bitset", "flags": 2}])
    for env in variants:
        try:
            m = rxm.Matcher(t, 0, **env)
        except rxm.RxmError as e:
            print("UPLOAD", regex, flags, env, e, "states", t.c.n_states, "edges", t.c.n_edges, "cells", t.c.n_cells); bad += 1
            continue
        try:
            got = m.match_host(chars, off)
            if not np.array_equal(got, bits):
                print("BITS", regex, flags, env, rxm.ENGINE_NAMES[m.plan().engine], int((got != bits).sum())); bad += 1
        except rxm.RxmError as e:
            print("MATCH", regex, flags, env, rxm.ENGINE_NAMES[m.plan().engine], e.status, "states", t.c.n_states, "edges", t.c.n_edges, "cells", t.c.n_cells); bad += 1
        m.close()
print("failures", bad)
