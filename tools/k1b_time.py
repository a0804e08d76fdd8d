"""Tuning helper: K1B (bit-set engine) throughput on the 77-node nfa_blowup automaton."""
import importlib.util, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("rxm", os.path.join(ROOT, "re2-modification_b200", "rxm.py"))
rxm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rxm)
t = rxm.Tables.load(os.path.join(ROOT, "tests", "golden", "cases", "nfa_blowup.rxt"))
m = rxm.Matcher(t, 0)
n = 200000
rng = torch.Generator(device="cuda").manual_seed(1)
lens = torch.randint(64, 4097, (n,), device="cuda", generator=rng)
off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); off[1:] = torch.cumsum(lens, 0)
chars = (torch.randint(0, 2, (int(off[-1]),), device="cuda", generator=rng) + 97).to(torch.uint8)
out = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s); torch.cuda.synchronize()
t0 = time.perf_counter(); m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("K1B blowup order=%s: %.1f ms, %.2f M strings/s, %.2f GB/s, match frac %.3f" % (os.environ.get("RXM_K3_ORDER", "sorted"), dt * 1e3, n / dt / 1e6, int(off[-1]) / dt / 1e9, float(out.float().mean())))
