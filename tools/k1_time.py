"""Tuning helper (not part of the product or the tests): time rxm_match_batch on the config-2
workload with device-resident buffers, no parity check.  Used with RXM_K1_VARIANT=... and with
probe builds of librxm.so (make EXTRA=-DRXM_K1_PROBE) to separate the staging path from the
lookups.  python tools/k1_time.py [n_strings] [steps]"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


rxm = _load("rxm", os.path.join(ROOT, "re2-modification_b200", "rxm.py"))
W = _load("workloads", os.path.join(ROOT, "re2-modification_b200", "workloads.py"))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
text = open(os.path.join(ROOT, "tests", "golden", "cases", "nfa_config2.rxt")).read()
t = rxm.Tables(text)
chars, off = W.alive_strings(text, n, 64, 4096, 5, "cuda")
out = torch.empty(n, dtype=torch.uint8, device="cuda")
flags = int(os.environ.get("RXM_K1_FLAGS", "0"))  # rxm.OPT_K1_NO_OCT = 8, OPT_K1_NO_QUAD = 1
m = rxm.Matcher(t, 0, flags=flags)
s = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
nbytes = int(off[-1].item())
w = torch.arange(1, n + 1, device="cuda", dtype=torch.int64) % 1000003
print(f"lib={os.path.basename(rxm.LIB_PATH)} variant={os.environ.get('RXM_K1_VARIANT', '0')} stride={m.plan().dfa_stride} ms/step={ms:.4f} "
      f"input_GB/s={nbytes / ms / 1e6:.0f} match_frac={float(out.float().mean()):.4f} checksum={int((out.to(torch.int64) * w).sum())}")
