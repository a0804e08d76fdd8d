"""GPU fuzz with LONG strings and LARGE batches on the random-expression corpus: for every
expression of tests/golden/fuzz/corpus.jsonl, strings are sampled from its language with heavily
pumped stars (up to several thousand letters, long backreference blocks), mixed with near misses,
in batches large enough for the tile-sorted hand-out order (>= 16384 strings); every device engine
against the C restatement (oracle/, the checker).  python tests/fuzz/fuzz_long_gpu.py [n_expr] [n_strings] [seed]"""
import os, random, sys, time
TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, TESTS)
import numpy as np
import helpers as H
from cases import load_fuzz_corpus
rxm = H.rxm


def parse(s):
    """The corpus generator's syntax -> AST (see tests/golden/make_fuzz_corpus.py)."""
    pos = 0

    def atom():
        nonlocal pos
        c = s[pos]
        if c == "(":
            pos += 1
            x = seq()
            if s[pos] == "|":
                pos += 1
                y = seq()
                x = ("alt", x, y)
            assert s[pos] == ")"
            pos += 1
            if pos < len(s) and s[pos] == "*":
                pos += 1
                return ("star", x)
            return x
        if c == "{":
            pos += 1
            x = seq()
            assert s[pos:pos + 2] == "}:"
            k = int(s[pos + 2])
            pos += 3
            return ("mem", x, k)
        if c == "&":
            pos += 2
            return ("ref", int(s[pos - 1]))
        pos += 1
        a = ("any",) if c == "." else ("lit", c)
        if pos < len(s) and s[pos] == "*":
            pos += 1
            return ("star", a)
        return a

    def seq():
        nonlocal pos
        x = atom()
        while pos < len(s) and s[pos] not in "|)}":
            x = ("cat", x, atom())
        return x

    out = seq()
    assert pos == len(s), (s, pos)
    return out


def sample(a, env, rng, pump):
    t = a[0]
    if t == "lit":
        return a[1]
    if t == "any":
        return rng.choice("abc")
    if t == "ref":
        return env.get(a[1], "")
    if t == "cat":
        x = sample(a[1], env, rng, pump)
        return x + sample(a[2], env, rng, pump)
    if t == "alt":
        return sample(a[1 + rng.randint(0, 1)], env, rng, pump)
    if t == "star":
        k = rng.choice([0, 1, 2, 3, pump, rng.randint(0, pump)])
        return "".join(sample(a[1], env, rng, max(1, pump // 4)) for _ in range(k))
    v = sample(a[1], env, rng, pump)
    env[a[2]] = v
    return v


def main():
    n_expr = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    n_str = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    rng = random.Random(seed)
    corpus = load_fuzz_corpus()
    rng.shuffle(corpus)
    corpus.sort(key=lambda c: -c[0].count("*"))  # expressions with stars first: they give the long strings
    bad = 0
    t_all = time.time()
    for regex, flags, kind, t, _, _ in corpus[:n_expr]:
        ast = parse(regex)
        strings = []
        while len(strings) < n_str:
            s = sample(ast, {}, rng, rng.choice([300, 1500]) if rng.random() < 0.05 else rng.choice([2, 6, 20, 60]))
            if len(s) > 6000:
                s = s[:6000]
            if s and rng.random() < 0.25:
                i = rng.randrange(len(s))
                s = rng.choice([s[:i] + rng.choice("abc") + s[i + 1:], s[:i] + s[i + 1:], s[:i] + rng.choice("abc") + s[i:]])
            if s:
                strings.append(s.encode())
        chars, off = H.make_batch(strings)
        want = H.oracle_bits(t, chars, off)
        variants = [{}] + ([{"engine": "k2"}] if kind == "mfa" else [{"engine": "bitset"}, {"flags": 1}])
        for env in variants:
            m = rxm.Matcher(t, 0, **env)
            got = m.match_host(chars, off)
            if not np.array_equal(got, want):
                bad += 1
                i = int(np.nonzero(got != want)[0][0])
                print("BITS", regex, flags, env, rxm.ENGINE_NAMES[m.plan().engine], int((got != want).sum()), "first", strings[i][:80], len(strings[i]))
            if not env:
                gt = m.match_text_host(b"\n".join(strings))
                if not np.array_equal(gt, want):
                    bad += 1
                    print("TEXT", regex, flags, int((gt != want).sum()) if len(gt) == len(want) else ("len", len(gt)))
            if m.overflow_count():
                bad += 1
                print("OVERFLOW", regex, flags, env, m.overflow_count())
            m.close()
        print("ok", regex, flags, kind, "mean len %.0f" % (len(chars) / n_str), "ones %.2f" % want.mean(), flush=True)
    print("failures", bad, "seconds %.0f" % (time.time() - t_all))


if __name__ == "__main__":
    main()
