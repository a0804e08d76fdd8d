"""Tuning helper (CPU): how much work a full step of the MFA simulation is, per BASELINE workload -- K4's per-string
core (rxm_k4_core.cuh) on the host with its statistics switched on.  python tools/k4_stats.py"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from cases import load_case  # noqa: E402

os.makedirs(os.path.join(ROOT, "tools", "build"), exist_ok=True)
SO = os.path.join(ROOT, "tools", "build", "libk4stat.so")
subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-x", "c++", "-o", SO,
                os.path.join(ROOT, "tools", "k4_stats.cpp"), os.path.join(H.PKG, "csrc", "rxm_plan.cpp")], check=True)
L = C.CDLL(SO)
L.hostsim_k4core_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p]
NAMES = ["find", "ins", "reins", "cmp", "cmpbytes", "src", "active", "waiting", "items", "fire"]
W = H.load_workloads()


def run(label, name, chars, off):
    t, _, _ = load_case(name)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    chars = np.ascontiguousarray(chars)
    got = np.empty(len(off) - 1, dtype=np.uint8)
    info, o = (C.c_uint64 * 3)(), (C.c_uint64 * 10)()
    L.k4stat_get(o)
    assert L.hostsim_k4core_batch(C.cast(t.ptr, C.c_void_p), chars.ctypes.data, off.ctypes.data, len(off) - 1,
                                  got.ctypes.data, 0, info) == 0
    L.k4stat_get(o)
    d = dict(zip(NAMES, [o[i] for i in range(10)]))
    s = max(int(info[0]), 1)
    print(f"{label:28s} {name}: letters {int(off[-1])}, full steps {info[0]}, repeated steps answered {info[1]}, strings handed on "
          f"{info[2]} | per full step: configurations {d['src'] / s:.2f} (active {d['active'] / s:.2f}), items walked "
          f"{d['items'] / s:.2f}, fired {d['fire'] / s:.2f}, inserted {d['ins'] / s:.2f}, re-inserted {d['reins'] / s:.2f}, block "
          f"compares {d['cmp'] / s:.3f} of {d['cmpbytes'] / max(d['cmp'], 1):.0f} bytes")


c, o = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", 48, 435, 20000, 3)
run("config 4 (attack strings)", "ex02_fwd", c, o)
run("config 4 (attack strings)", "ex02_rev", c, o)
ch, of = W.example5_strings(2000, 64, 4096, 1000, "cpu")
run("config 3", "ex05_fwd", ch.numpy(), of.numpy())
for ex in range(1, 11):
    c, o = W.mixed_example_batch(ex, 3000, 1000 * ex)
    run("config 5 (README example %d)" % ex, f"ex{ex:02d}_fwd", c, o)
