"""Tuning helper (GPU): ONE MFA job timed with device-resident buffers (for ncu captures).
python tools/mfa_one.py <case e.g. ex09_fwd> <example number for the mixed batch | config3> <n strings> [engine] [steps]"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "re2-modification_b200")


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


rxm, W = _load("rxm"), _load("workloads")
case, what, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
engine = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "auto" else None
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
if what == "config3":
    ch, of = W.example5_strings(n, 64, 4096, 1000, dev)
else:
    c_np, o_np = W.mixed_example_batch(int(what), n, 1000 * int(what))
    ch, of = torch.from_numpy(c_np).to(dev), torch.from_numpy(o_np.astype(np.int64)).to(dev)
m = rxm.Matcher(rxm.Tables.load(os.path.join(ROOT, "tests", "golden", "cases", case + ".rxt")), 0, engine=engine)
out = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    m.match_ptrs(ch.data_ptr(), of.data_ptr(), n, out.data_ptr(), s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    m.match_ptrs(ch.data_ptr(), of.data_ptr(), n, out.data_ptr(), s)
e1.record()
torch.cuda.synchronize()
print(f"{case} {what} engine={rxm.ENGINE_NAMES[m.plan().engine]} n={n} bytes={int(of[-1])} "
      f"ms/step={e0.elapsed_time(e1) / steps:.3f} match_frac={float(out.float().mean()):.4f} overflow={m.overflow_count()}")
