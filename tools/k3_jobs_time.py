"""Tuning helper (GPU): config 4's two jobs (example 2 forward / -reverse on attack strings) timed
alone and side by side, with and without rxm_set_concurrency.  python tools/k3_jobs_time.py"""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "re2-modification_b200")


def _load(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


rxm, W = _load("rxm"), _load("workloads")
dev = torch.device("cuda:0")
c_np, o_np = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", 4096, 435, 65536, 1000)
ch, of = torch.from_numpy(c_np).to(dev), torch.from_numpy(o_np.astype(np.int64)).to(dev)
n = of.numel() - 1
jobs = {}
for name in ("ex02_fwd", "ex02_rev"):
    t = rxm.Tables.load(os.path.join(ROOT, "tests", "golden", "cases", name + ".rxt"))
    jobs[name] = (rxm.Matcher(t, 0), torch.empty(n, dtype=torch.uint8, device=dev), torch.cuda.Stream(device=dev))


def run(names, share, reps=3):
    for nm in names:
        jobs[nm][0].set_concurrency(share)
    best = 1e9
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record(cur)
        evs = []
        for nm in names:
            m, out, st = jobs[nm]
            st.wait_stream(cur)
            m.match_ptrs(ch.data_ptr(), of.data_ptr(), n, out.data_ptr(), st.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(st)
            evs.append(ev)
        for ev in evs:
            cur.wait_event(ev)
        e1.record(cur)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for names, share in ((["ex02_fwd"], 1), (["ex02_rev"], 1), (["ex02_fwd"], 2), (["ex02_rev"], 2),
                     (["ex02_fwd", "ex02_rev"], 1), (["ex02_fwd", "ex02_rev"], 2)):
    print("+".join(names), "share", share, "%.1f ms" % run(names, share), flush=True)
