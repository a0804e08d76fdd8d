"""GPU fuzz on RANDOM TABLES (not produced by the reference's builders): the C ABI takes any valid
rxm_tables, so the kernels must agree with the C restatement (oracle/, the checker) on structures
the builders never emit -- letter targets with incoming epsilon edges (K1B then walks the edges),
parallel edges, `never` edges, read edges whose cell nothing opens, open and close of several cells
on one edge.  Epsilon edges only go to higher state numbers (no epsilon cycle: the reference itself
overflows its stack there).  A table an engine declines with RXM_ERR_UNSUPPORTED is counted, not an
error.  python tests/fuzz/fuzz_tables_gpu.py [n_tables] [seed]"""
import os, random, sys
TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, TESTS)
import numpy as np
import helpers as H
rxm = H.rxm


def random_table(rng, mfa):
    n = rng.randint(2, 9)
    cells = rng.randint(1, 3) if mfa else 0
    lines, n_edges = [], 0
    for q in range(n):
        for _ in range(rng.choice([0, 1, 1, 2, 2, 3, 4])):
            r = rng.random()
            if r < 0.22 and q + 1 < n:
                kind, sym, to = "E", "-", rng.randint(q + 1, n - 1)
            elif r < 0.30:
                kind, sym, to = "A", "-", rng.randrange(n)
            elif r < 0.34:
                kind, sym, to = "N", "-", rng.randrange(n)
            elif mfa and r < 0.55:
                kind, sym, to = "L", str(rng.randint(1, cells)), rng.randrange(n)
            else:
                kind, sym, to = "L", rng.choice("aabbc"), rng.randrange(n)
            acts = ""
            if mfa:
                for k in range(1, cells + 1):
                    x = rng.random()
                    if x < 0.25:
                        acts += f" o{k}"
                    elif x < 0.45:
                        acts += f" c{k}"
            lines.append(f"{q} {kind} {sym} {to}{acts}")
            n_edges += 1
    text = (f"rxm-tables 1\nkind {'mfa' if mfa else 'nfa'}\nreversed {rng.randint(0, 1)}\nstates {n}\nstart 0\n"
            f"finish {rng.randrange(n)}\ncells {cells}\nedges {n_edges}\n" + "".join(l + "\n" for l in lines) + "end\n")
    return text


def main():
    n_tables = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
    rng = random.Random(seed)
    nprng = np.random.default_rng(seed)
    bad = declined = done = 0
    for it in range(n_tables):
        mfa = rng.random() < 0.6
        text = random_table(rng, mfa)
        try:
            t = rxm.Tables(text)
        except rxm.RxmError as e:
            print("PARSE", e, text)
            bad += 1
            continue
        strings = [b"", b"a", b"b", b"ab", b"ba", b"aab", b"abab", b"cc"]
        for _ in range(300):
            L = int(nprng.integers(0, 40))
            strings.append(bytes(nprng.choice(np.frombuffer(b"aaabbc1", dtype=np.uint8), size=L)))
        for _ in range(12):
            u = bytes(nprng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=int(nprng.integers(1, 6))))
            strings.append(u * int(nprng.integers(2, 60)))
        chars, off = H.make_batch(strings)
        want = H.oracle_bits(t, chars, off)
        variants = [{}] + ([{"engine": "k2"}, {"engine": "k3"}, {"engine": "k4"}] if mfa else
                           [{"engine": "bitset"}, {"engine": "bitset", "flags": 2}, {"flags": 1}])
        for env in variants:
            try:
                m = rxm.Matcher(t, 0, **env)
            except rxm.RxmError as e:
                if e.status == rxm.RXM_ERR_UNSUPPORTED:
                    declined += 1
                    continue
                print("UPLOAD", env, e, text)
                bad += 1
                continue
            try:
                got = m.match_host(chars, off)
                done += 1
                if not np.array_equal(got, want):
                    bad += 1
                    i = int(np.nonzero(got != want)[0][0])
                    print("BITS", env, rxm.ENGINE_NAMES[m.plan().engine], int((got != want).sum()), "first", strings[i], "want", want[i])
                    print(text)
            except rxm.RxmError as e:
                if e.status == rxm.RXM_ERR_OVERFLOW:
                    declined += 1  # a string hit a kernel limit and was reported -- never a wrong bit
                else:
                    print("MATCH", env, e, text)
                    bad += 1
            m.close()
    print("tables", n_tables, "engine runs", done, "declined", declined, "failures", bad)


if __name__ == "__main__":
    main()
