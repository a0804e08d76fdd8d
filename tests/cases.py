"""Access to the committed golden fixtures (tests/golden/cases, manifest.json)."""
from __future__ import annotations

import json
import os

import numpy as np

import helpers as H

with open(os.path.join(H.GOLDEN, "manifest.json")) as f:
    MANIFEST = json.load(f)
CASE_NAMES = [m["name"] for m in MANIFEST]
BY_NAME = {m["name"]: m for m in MANIFEST}


def load_case(name):
    """-> (Tables, list[bytes], bits uint8[n])"""
    t = H.rxm.Tables.load(os.path.join(H.GOLDEN, "cases", name + ".rxt"))
    strings, bits = [], []
    with open(os.path.join(H.GOLDEN, "cases", name + ".golden")) as f:
        for line in f:
            if line.startswith("#"):
                continue
            line = line.rstrip("\n")
            b, s = line[0], line[2:]
            bits.append(int(b))
            strings.append(b"" if s == "<empty>" else s.encode())
    return t, strings, np.array(bits, dtype=np.uint8)


def load_fuzz_corpus(path=None):
    """Random-expression fixtures (tests/golden/make_fuzz_corpus.py; bits from the reference's own code).
    -> list of (regex, flags, kind, Tables, list[bytes], bits uint8[n])"""
    out = []
    paths = [path] if path else [os.path.join(H.GOLDEN, "fuzz", "corpus.jsonl"), os.path.join(H.GOLDEN, "fuzz", "corpus_9cells.jsonl")]
    for pth in paths:
        for line in open(pth):
            c = json.loads(line)
            out.append((c["regex"], c["flags"], c["kind"], H.rxm.Tables(c["tables"]),
                        [s.encode() for s in c["strings"]],
                        np.frombuffer(c["bits"].encode(), dtype=np.uint8) - ord("0")))
    return out
