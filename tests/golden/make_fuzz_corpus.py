#!/usr/bin/env python3
"""Random-EXPRESSION fixtures: golden bits from the reference's own code for expressions it never
shipped.  Expressions are drawn from the README grammar (literals a-c and `.`, alternation, star,
memory cells {r}:k, references &k), compiled by the product front end (bin/rxm_compile == the
reference's parse/compile + the flattening stage), forward and -reverse; strings are random over
{a,b,c}.  The bits come from oracle/_ref/diploma_ref_bump.  Expressions the reference cannot
handle (front end crashes / loops, epsilon cycles, tables outside the device limits) are skipped
and counted.  The C restatement must agree with the reference on every kept case, or the script
stops.  Output: tests/golden/fuzz/corpus.jsonl (one case per line).

Needs /root/reference builds (oracle/_ref, bin/rxm_compile).  Run:  python tests/golden/make_fuzz_corpus.py [n_cases] [seed] [output.jsonl] [max_cell]
"""
from __future__ import annotations

import json
import os
import random
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402


MAX_CELL = 3  # cell names drawn from 1..MAX_CELL (argv[4]; 9 exercises the engines' many-cell paths)


def gen(rng: random.Random, depth: int, cells: list[int], under_star: bool):
    """-> AST: ("lit", c) | ("any",) | ("ref", k) | ("cat", x, y) | ("alt", x, y) | ("star", x) | ("mem", x, k)"""
    r = rng.random()
    if depth <= 0 or r < 0.30:
        x = rng.random()
        if x < 0.72 or not cells and x < 0.9:
            return ("lit", rng.choice("aabbc"))
        if x < 0.78:
            return ("any",)
        return ("ref", rng.choice(cells) if cells and rng.random() < 0.85 else rng.randint(1, MAX_CELL))
    if r < 0.55:
        return ("cat", gen(rng, depth - 1, cells, under_star), gen(rng, depth - 1, cells, under_star))
    if r < 0.70:
        return ("alt", gen(rng, depth - 1, cells, under_star), gen(rng, depth - 1, cells, under_star))
    if r < 0.85 and not under_star:  # no star directly under a star: the reference loops on epsilon cycles
        if rng.random() < 0.5:
            return ("star", ("lit", rng.choice("abc")))
        return ("star", gen(rng, depth - 1, cells, True))
    k = rng.randint(1, MAX_CELL)
    inner = gen(rng, depth - 1, cells, under_star)
    if k not in cells:
        cells.append(k)
    return ("mem", inner, k)


def render(a) -> str:
    t = a[0]
    if t == "lit":
        return a[1]
    if t == "any":
        return "."
    if t == "ref":
        return "&" + str(a[1])
    if t == "cat":
        return render(a[1]) + render(a[2])
    if t == "alt":
        return "(" + render(a[1]) + "|" + render(a[2]) + ")"
    if t == "star":
        return (a[1][1] + "*") if a[1][0] == "lit" else "(" + render(a[1]) + ")*"
    return "{" + render(a[1]) + "}:" + str(a[2])


def sample(a, env: dict, rng: random.Random) -> str:
    """A string of the expression's language under the usual backreference reading (an unset cell
    reads as empty) -- the reference mostly, not always, accepts these."""
    t = a[0]
    if t == "lit":
        return a[1]
    if t == "any":
        return rng.choice("abc")
    if t == "ref":
        return env.get(a[1], "")
    if t == "cat":
        x = sample(a[1], env, rng)
        return x + sample(a[2], env, rng)
    if t == "alt":
        return sample(a[1 + rng.randint(0, 1)], env, rng)
    if t == "star":
        return "".join(sample(a[1], env, rng) for _ in range(rng.choice([0, 1, 1, 2, 3, 5])))
    v = sample(a[1], env, rng)
    env[a[2]] = v
    return v


def strings_for(rng: random.Random, ast) -> list[bytes]:
    out = [b"a", b"b", b"aa", b"ab", b"ba", b"abc", b"aab", b"aabaab", b"abab", b"cc"]
    for _ in range(60):  # in (or near) the language
        s = sample(ast, {}, rng)
        if s and len(s) <= 120 and rng.random() < 0.3:  # a near miss: one letter changed / dropped / added
            i = rng.randrange(len(s))
            s = rng.choice([s[:i] + rng.choice("abc") + s[i + 1:], s[:i] + s[i + 1:], s[:i] + rng.choice("abc") + s[i:]])
        if 0 < len(s) <= 120:
            out.append(s.encode())
    for _ in range(60):
        L = rng.randint(1, 24)
        out.append("".join(rng.choice("aaabbc") for _ in range(L)).encode())
    for _ in range(10):
        L = rng.randint(30, 90)
        out.append("".join(rng.choice("aab") for _ in range(L)).encode())
    for _ in range(10):  # periodic: backreference blocks repeat
        u = "".join(rng.choice("ab") for _ in range(rng.randint(1, 5)))
        out.append((u * rng.randint(2, 12) + rng.choice(["", "a", "b", "c"])).encode())
    return out


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    if not H.have_reference():
        raise SystemExit("needs oracle/_ref (make -C oracle) and bin/rxm_compile (make -C re2-modification_b200 front)")
    global MAX_CELL
    if len(sys.argv) > 4:
        MAX_CELL = int(sys.argv[4])
    rng = random.Random(seed)
    kept, skipped, seen = [], {"compile": 0, "tables": 0, "reference": 0, "dup": 0}, set()
    attempts = 0
    while len(kept) < n_cases and attempts < 40 * n_cases:
        attempts += 1
        ast = gen(rng, rng.randint(3, 5), [], False)
        regex = render(ast)
        if len(regex) < 6:
            continue
        flags = ["-reverse"] if rng.random() < 0.4 else []
        if (regex, tuple(flags)) in seen or len(regex) > 48:
            skipped["dup"] += 1
            continue
        seen.add((regex, tuple(flags)))
        try:
            r = subprocess.run([H.RXM_COMPILE, "-match", *flags, "-regex", regex], capture_output=True, timeout=10)
        except subprocess.TimeoutExpired:
            skipped["compile"] += 1
            continue
        if r.returncode != 0 or not r.stdout.startswith(b"rxm-tables"):
            skipped["compile"] += 1
            continue
        text = r.stdout.decode()
        try:
            t = H.rxm.Tables(text)
        except H.rxm.RxmError:
            skipped["tables"] += 1
            continue
        strings = strings_for(rng, ast)
        chars, off = H.make_batch(strings)
        try:
            ref = H.reference_bits(regex, flags, chars, off, binary=H.REF_BUMP, timeout=60)
        except (subprocess.CalledProcessError, subprocess.TimeoutExpired, FileNotFoundError):
            skipped["reference"] += 1
            continue
        if len(ref) != len(strings):
            skipped["reference"] += 1
            continue
        mine = H.oracle_bits(t, chars, off)
        if not np.array_equal(mine, ref):
            bad = np.nonzero(mine != ref)[0]
            raise SystemExit(f"C restatement differs from the reference: {regex!r} {flags}: "
                             f"{len(bad)} strings, first {strings[bad[0]]!r}")
        kept.append({"regex": regex, "flags": flags, "kind": "mfa" if t.c.kind == 1 else "nfa",
                     "tables": text, "strings": [s.decode() for s in strings],
                     "bits": "".join(str(int(b)) for b in ref)})
        if len(kept) % 10 == 0:
            print(len(kept), "kept;", skipped, flush=True)
    out_path = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else os.path.join(HERE, "fuzz", "corpus.jsonl")
    with open(out_path, "w") as f:
        for c in kept:
            f.write(json.dumps(c) + "\n")
    print("kept", len(kept), "skipped", skipped, "mfa", sum(c["kind"] == "mfa" for c in kept),
          "ones", sum(c["bits"].count("1") for c in kept))


if __name__ == "__main__":
    main()
