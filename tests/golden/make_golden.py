#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

The reference ships no match-bit vectors (its test/example_*/*_results.txt are
timings), so the golden bits come from the reference's own compiled code:
oracle/_ref/diploma_ref_bump (oracle/Makefile; reference objects + never-reuse
allocator = the canonical tie-break).  The stock-glibc build is run on the same
batches and the strings where it disagrees are recorded in manifest.json.

Inputs per expression: the reference's own fixtures -- test/example_N/regexp.txt
line 1, pump.txt via the attack-string recipe of
matchers/example_runner.cpp:15-29,123 (cumulative) and matcher.py:26-38
(non-cumulative), test/example_N/input_strings.txt -- plus seeded random and
near-miss strings.  Tables come from the product front end (bin/rxm_compile).

Needs /root/reference (build time only).  Run:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

REF = os.environ.get("RXM_REF", "/root/reference")
CASES_DIR = os.path.join(HERE, "cases")

NFA_EXPRESSIONS = [
    # (name, regex) -- memory-free expressions covering the three branches of
    # Regexp::compile (regex.cpp:315-341): forward Glushkov, reversed Glushkov, Thompson
    ("config2", "(a|bb)*aaba(a|bb*aa)*", ["aaba", "bbaaba", "aabaa", "aababbaa", "abbaabaabbbbaa"]),  # BASELINE.json configs[1]; Thompson branch
    ("quirk", "(b|a)(b)*baa(b|a)(a)*(a)*", ["bbbaaa"]),  # `visited` filters a letter edge
    ("astarb", "a*b", ["b", "aaab"]),
    ("abb", "(a|b)*abb", ["abb", "babb"]),
    ("dot", "ab*c.d", ["acxd", "abbbcad", "ac.d", "acd", "abcdd"]),
    ("alt", "a(b|c)*d", ["ad", "abcbd", "abcb"]),
    ("dots", ".*ab.*", ["ab", "xxabyy", "ba"]),
    ("abba", "(ab|ba)*", ["abba", "abab", "aba"]),
    ("third", "(a|b)*a(a|b)(a|b)", ["abb", "aaa", "bab"]),
    ("lit", "abcabc", ["abcabc", "abcab", "abcabcc"]),
    # an `a`, six letters, a `b`: 2^8 active sets -> more than 128, K1's two-lookup table form (K1_CLASSED)
    ("mid", "(a|b)*a" + "(a|b)" * 6 + "b(a|b)*",
     ["a" + "a" * 6 + "b", "b" * 30, "ab" * 20, "a" * 7 + "b", "a" * 6 + "b", "ba" + "b" * 6 + "ba"]),
    # an `a`, twelve letters, a `b`: 24 577 active sets -> the largest table K1 holds in shared memory (172 KB,
    # two-lookup form); Thompson branch, 77 nodes.  (Round 1's planner stopped at 4096 sets: K1B ran it.)
    ("blowup", "(a|b)*a" + "(a|b)" * 12 + "b(a|b)*",
     ["a" + "a" * 12 + "b", "b" * 30, "ab" * 20, "a" * 13 + "b", "a" * 12 + "b", "ba" + "b" * 12 + "ba"]),
    # fourteen letters in between: ~98 000 active sets -> no table fits, the planner hands it to the bit-set
    # engine (K1B); Thompson branch, 89 nodes
    ("huge", "(a|b)*a" + "(a|b)" * 14 + "b(a|b)*",
     ["a" + "a" * 14 + "b", "b" * 30, "ab" * 20, "a" * 15 + "b", "a" * 14 + "b", "ba" + "b" * 14 + "ba"]),
]


def pumped_string(n: int, pump: list[str]) -> str:
    """matchers/example_runner.cpp:15-29 (== matcher.py:26-38)."""
    pump_count = len(pump) // 2 + 1
    del_count = len(pump) - pump_count
    res = pump[0]
    while len(res) + len(pump[0]) < (n - del_count) // pump_count:
        res += pump[0]
    if len(pump) == 1:
        return res
    return (res + pump[1]) * del_count + res


def read_pump(i: int):
    lines = open(f"{REF}/test/example_{i}/pump.txt").read().split("\n")
    lines += ["", "", ""]
    pump = lines[0].split(",")
    return pump, lines[1], lines[2]


def strings_for_example(i: int, regex: str, rng: random.Random) -> list[bytes]:
    pump, suffix, prefix = read_pump(i)
    out: list[str] = [""]
    # (1) non-cumulative attack strings (matcher.py) with and without the failing suffix
    for n in (2, 5, 9, 16, 33, 64, 100, 180, 300):
        p = pumped_string(n, pump)
        out.append(prefix + p + suffix)
        out.append(prefix + p)
        out.append(p)
    # (2) cumulative variant (example_runner.cpp:123 mutates `prefix`)
    cum = prefix
    for n in (6, 12, 24, 48):
        cum = cum + pumped_string(n, pump) + suffix
        out.append(cum)
    # (3) the reference's own input_strings.txt (bounded so the fixture stays small)
    path = f"{REF}/test/example_{i}/input_strings.txt"
    if os.path.exists(path):
        lines = [ln.strip() for ln in open(path) if ln.strip()]
        lines.sort(key=len)
        short = [ln for ln in lines if len(ln) <= 700]
        out += short[:40]
        out += [ln for ln in lines if len(ln) > 700][:2]
    # (4) seeded random strings over the expression's letters (+ one foreign letter)
    letters = sorted({c for c in regex if c.isalpha()}) or ["a"]
    foreign = next(c for c in "zyxwvu" if c not in letters)
    for _ in range(120):
        L = rng.randint(1, 24)
        out.append("".join(rng.choice(letters) for _ in range(L)))
    for _ in range(60):
        L = rng.randint(1, 40)
        out.append("".join(letters[0] if rng.random() < 0.75 else rng.choice(letters)
                           for _ in range(L)))
    for _ in range(20):
        L = rng.randint(1, 12)
        out.append("".join(rng.choice(letters + [foreign]) for _ in range(L)))
    # (5) near misses: one edit applied to a pumped string
    for _ in range(60):
        p = prefix + pumped_string(rng.choice((6, 12, 20, 40)), pump)
        if rng.random() < 0.3:
            p += suffix
        k = rng.randrange(len(p) + 1)
        r = rng.random()
        if r < 0.4 and k < len(p):
            p = p[:k] + rng.choice(letters) + p[k + 1:]
        elif r < 0.7:
            p = p[:k] + rng.choice(letters) + p[k:]
        elif k < len(p):
            p = p[:k] + p[k + 1:]
        out.append(p)
    # a digit in the input meets a read edge labelled with the same digit (mfa.cpp:171)
    out += ["ab1", "a1", "1", "aa1aa", "a2a"]
    return [s.encode() for s in out]


def strings_for_nfa(regex: str, rng: random.Random) -> list[bytes]:
    letters = sorted({c for c in regex if c.isalpha()}) or ["a"]
    foreign = next(c for c in "zyxwvu" if c not in letters)
    out = [""]
    # every string up to length 6 over the first two letters (covers the quirk case exhaustively)
    two = letters[:2] if len(letters) >= 2 else letters + [foreign]
    for L in range(1, 7):
        for x in range(len(two) ** L):
            s, y = "", x
            for _ in range(L):
                s += two[y % len(two)]
                y //= len(two)
            out.append(s)
    for _ in range(150):
        L = rng.randint(1, 30)
        out.append("".join(rng.choice(letters) for _ in range(L)))
    for _ in range(30):
        L = rng.randint(1, 12)
        out.append("".join(rng.choice(letters + [foreign, "."]) for _ in range(L)))
    for _ in range(10):
        L = rng.randint(200, 900)
        out.append("".join(rng.choice(letters) for _ in range(L)))
    return [s.encode() for s in out]


def write_case(name: str, regex: str, flags: list[str], strings: list[bytes]) -> dict:
    text = H.compile_tables_text(regex, flags)
    t = H.rxm.Tables(text)
    chars, off = H.make_batch(strings)
    bits = H.reference_bits(regex, flags, chars, off, binary=H.REF_BUMP)
    stock = H.reference_bits(regex, flags, chars, off, binary=H.REF_STOCK)
    mine = H.oracle_bits(t, chars, off)
    if not np.array_equal(mine, bits):
        bad = np.nonzero(mine != bits)[0]
        raise SystemExit(f"{name}: C restatement differs from the reference on {len(bad)} strings, "
                         f"first: {strings[bad[0]]!r}")
    with open(os.path.join(CASES_DIR, name + ".rxt"), "w") as f:
        f.write(text)
    with open(os.path.join(CASES_DIR, name + ".golden"), "w") as f:
        f.write(f"# regex {regex}\n# flags {' '.join(flags)}\n")
        for s, b in zip(strings, bits):
            f.write(f"{int(b)} {s.decode() if s else '<empty>'}\n")
    dis = [int(i) for i in np.nonzero(stock != bits)[0]]
    c = t.c
    return {
        "name": name, "regex": regex, "flags": flags,
        "kind": "mfa" if c.kind == 1 else "nfa", "reversed": int(c.reversed),
        "n_states": int(c.n_states), "n_edges": int(c.n_edges), "n_cells": int(c.n_cells),
        "strings": len(strings), "ones": int(bits.sum()),
        "stock_glibc_disagrees": dis,
    }


def main():
    only = set(sys.argv[1:])  # optional: case names to (re)generate; the others keep their files
    if not H.have_reference():
        raise SystemExit("needs oracle/_ref (make -C oracle) and bin/rxm_compile "
                         "(make -C re2-modification_b200 front)")
    os.makedirs(CASES_DIR, exist_ok=True)
    manifest = []
    old = {}
    if only and os.path.exists(os.path.join(HERE, "manifest.json")):
        old = {m["name"]: m for m in json.load(open(os.path.join(HERE, "manifest.json")))}
    for i in range(1, 18):
        regex = open(f"{REF}/test/example_{i}/regexp.txt").readline().strip()
        for flags in ([], ["-reverse"]):
            rng = random.Random(1000 * i + len(flags))
            name = f"ex{i:02d}" + ("_rev" if flags else "_fwd")
            if only and name not in only:
                if name in old:
                    manifest.append(old[name])
                continue
            m = write_case(name, regex, flags, strings_for_example(i, regex, rng))
            manifest.append(m)
            print(m)
    for name, regex, extra in NFA_EXPRESSIONS:
        if only and "nfa_" + name not in only:
            if "nfa_" + name in old:
                manifest.append(old["nfa_" + name])
            continue
        strings = strings_for_nfa(regex, random.Random(len(name) * 7919)) + [x.encode() for x in extra]
        m = write_case("nfa_" + name, regex, [], strings)
        manifest.append(m)
        print(m)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
