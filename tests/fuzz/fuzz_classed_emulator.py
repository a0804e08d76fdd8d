"""CPU fuzz of K1's two-lookup tables and their strides on the SIMT emulator (the kernel SOURCE, rxm_k1.cu, through
rxm::k1_launch): random memory-free tables over {a, b} (plus `.` edges now and then) whose determinisation has more
than 64 sets (two-lookup tables; a few direct ones where a `.` edge widens nothing), forward and right-to-left, every
stride the tables allow (4 / 1 bytes per lookup) against the C
restatement (oracle/, the checker).  python tests/fuzz/fuzz_classed_emulator.py [n_tables] [seed]"""
import os, random, sys
TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, TESTS)
import numpy as np
import helpers as H
import test_oracle_golden as T


def random_table(rng):
    n = rng.randint(9, 16)
    lines = []
    for q in range(n):
        for _ in range(rng.choice([1, 2, 2, 3, 3, 4])):
            r = rng.random()
            if r < 0.10 and q + 1 < n:
                lines.append(f"{q} E - {rng.randint(q + 1, n - 1)}")
            elif r < 0.14:
                lines.append(f"{q} A - {rng.randrange(n)}")
            else:
                lines.append(f"{q} L {rng.choice('ab')} {rng.randrange(n)}")
    return (f"rxm-tables 1\nkind nfa\nreversed {rng.randint(0, 1)}\nstates {n}\nstart 0\nfinish {rng.randrange(n)}\n"
            f"cells 0\nedges {len(lines)}\n" + "".join(l + "\n" for l in lines) + "end\n")


def main():
    n_tables = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rng, nprng = random.Random(seed), np.random.default_rng(seed)
    hs = T.hostsim()
    tried = used = bad = 0
    strides = {1: 0, 4: 0, 8: 0}
    while used < n_tables and tried < 200 * n_tables:
        tried += 1
        text = random_table(rng)
        t = H.rxm.Tables(text)
        strings = [b"", b"a", b"b"]
        strings += [bytes(nprng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=int(L))) for L in nprng.integers(0, 120, size=120)]
        strings += [bytes(nprng.choice(np.frombuffer(b"aabbz", dtype=np.uint8), size=int(L))) for L in nprng.integers(0, 90, size=30)]
        strings += [bytes(nprng.choice(np.frombuffer(b"ab", dtype=np.uint8), size=int(L))) for L in nprng.integers(300, 700, size=6)]
        rc, got, ovf, info, msg = T.k1_emulated(t, strings, seed=tried)
        if rc == H.rxm.RXM_ERR_UNSUPPORTED or (rc == 0 and info[0] <= 64):
            continue
        used += 1
        chars, off = H.make_batch(strings)
        want = H.oracle_bits(t, chars, off)
        for flags in (0, 2, 1):
            hs.hostsim_k1_no_quad(flags)
            try:
                rc, got, ovf, info, msg = T.k1_emulated(t, strings, seed=tried + flags)
            finally:
                hs.hostsim_k1_no_quad(0)
            strides[info[2]] += 1
            if rc != 0 or ovf != 0 or not np.array_equal(got, want):
                bad += 1
                i = int(np.nonzero(got != want)[0][0]) if rc == 0 and len(got) == len(want) and (got != want).any() else -1
                print("FAIL", rc, msg, "sets", info[0], "stride", info[2], "first", strings[i] if i >= 0 else None)
                print(text)
    print("tables tried", tried, "with more than 64 sets", used, "runs per stride", strides, "failures", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
