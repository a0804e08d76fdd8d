"""Tuning helper (GPU; not part of the product or the tests): the MFA engines timed on the bench
workloads with device-resident buffers.  The engines (rxm_tables_upload_opts; "auto" = the planner's choice) are compared
inside one process:  python tools/mfa_time.py [config3|config4|config5] [engines, e.g. k4,k3] [n] [steps]"""
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
wl = sys.argv[1] if len(sys.argv) > 1 else "config3"
engines = (sys.argv[2] if len(sys.argv) > 2 else "k4,k3").split(",")
n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0")
case = lambda nm: rxm.Tables.load(os.path.join(ROOT, "tests", "golden", "cases", nm + ".rxt"))
jobs = []
if wl == "config3":
    ch, of = W.example5_strings(n or 1_000_000, 64, 4096, 1000, dev)
    jobs.append(("ex05_fwd", ch, of))
elif wl == "config4":
    c_np, o_np = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", n or 4096, 435, 65536, 1000)
    ch, of = torch.from_numpy(c_np).to(dev), torch.from_numpy(o_np.astype(np.int64)).to(dev)
    jobs += [("ex02_fwd", ch, of), ("ex02_rev", ch, of)]
else:
    for ex in range(1, 11):
        c_np, o_np = W.mixed_example_batch(ex, (n or 1_000_000) // 10, 1000 * ex)
        jobs.append((f"ex{ex:02d}_fwd", torch.from_numpy(c_np).to(dev), torch.from_numpy(o_np.astype(np.int64)).to(dev)))
s = torch.cuda.current_stream().cuda_stream
ref = {}
for eng in engines:
    total = 0.0
    for name, ch, of in jobs:
        nn = of.numel() - 1
        m = rxm.Matcher(case(name), 0, engine=None if eng == "auto" else eng)
        out = torch.empty(nn, dtype=torch.uint8, device=dev)
        for _ in range(2):
            m.match_ptrs(ch.data_ptr(), of.data_ptr(), nn, out.data_ptr(), s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            m.match_ptrs(ch.data_ptr(), of.data_ptr(), nn, out.data_ptr(), s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        total += ms
        same = ""
        if name in ref:
            same = " same_bits=%s" % bool(torch.equal(ref[name], out))
        else:
            ref[name] = out.clone()
        print(f"{wl} {name} engine={rxm.ENGINE_NAMES[m.plan().engine]} n={nn} bytes={int(of[-1])} ms/step={ms:.3f} "
              f"GB/s={int(of[-1]) / ms / 1e6:.1f} match_frac={float(out.float().mean()):.4f} overflow={m.overflow_count()}{same}",
              flush=True)
        m.close()
    print(f"{wl} engine={eng} total ms/step (jobs one after another) = {total:.3f}", flush=True)
