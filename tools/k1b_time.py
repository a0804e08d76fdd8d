"""Tuning helper (GPU; not part of the product or the tests): memory-free automata with large determinisations.
nfa_blowup (24 577 active sets: K1's largest two-lookup table) on the planner's choice and on the bit-set engine,
nfa_huge (~98 000 sets: no table) on the bit-set engine.  200 k random {a,b} strings of 64-4096 letters."""
import importlib.util, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("rxm", os.path.join(ROOT, "re2-modification_b200", "rxm.py"))
rxm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rxm)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = torch.Generator(device="cuda").manual_seed(1)
lens = torch.randint(64, 4097, (n,), device="cuda", generator=rng)
off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
chars = (torch.randint(0, 2, (int(off[-1]),), device="cuda", generator=rng) + 97).to(torch.uint8)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
ref = {}
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None  # e.g. nfa_mid (for an ncu capture of one kernel)
cases = [("nfa_blowup", None), ("nfa_blowup", "bitset"), ("nfa_huge", None), ("nfa_mid", None)]
if only and any(c.endswith(".rxt") for c in only):  # table files (tools/build/*.rxt) on the planner's engine
    cases = [(c, None) for c in only]
for case, engine in cases:
    if only and case not in only:
        continue
    t = rxm.Tables.load(case if case.endswith(".rxt") else os.path.join(ROOT, "tests", "golden", "cases", case + ".rxt"))
    m = rxm.Matcher(t, 0, engine=engine, flags=int(os.environ.get("RXM_K1_FLAGS", "0")))
    p = m.plan()
    for _ in range(2):
        m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    bits = out.clone()
    same = ""
    if case in ref:
        same = " same_bits=%s" % bool(torch.equal(ref[case], bits))
    ref.setdefault(case, bits)
    print("%s engine=%s sets=%d stride=%d: %.3f ms, %.2f M strings/s, %.1f GB/s, match frac %.3f%s" % (
        case, rxm.ENGINE_NAMES[p.engine], p.dfa_states, p.dfa_stride, ms, n / ms / 1e3, int(off[-1]) / ms / 1e6,
        float(out.float().mean()), same), flush=True)
    m.close()
