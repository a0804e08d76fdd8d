#!/usr/bin/env python3
"""Headline benchmark of the batch matcher (BASELINE.json metric / configs[1]).

One "step" = one pass of the hot path (rxm_match_batch, include/rxm.h) over one
batch of synthetic strings.  Default workload at any N: per GPU, 1 000 000 random
{a,b} strings of 64-4096 chars against the config-2 expression
(a|bb)*aaba(a|bb*aa)* -- the "alive" set of SURVEY.md 8d (random walks on the
automaton, last letter flipped for half of them), because i.i.d. uniform strings
die within a few letters and would time the early exit, not the scan.  The
uniform set is timed too and reported under "uniform_iid".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--strings S] [--workload config2|config3]
  python bench.py --impl reference ...      the reference's own CPU code on host cores

Under torchrun each rank drives its own GPU on its own shard (strings are
independent: no data-path collective); times are max over ranks.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "re2-modification_b200")
CASES = os.path.join(ROOT, "tests", "golden", "cases")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

WORKLOADS = {
    # name: (table fixture, expression, flags, description)
    "config2": ("nfa_config2", "(a|bb)*aaba(a|bb*aa)*", [],
                "config2: (a|bb)*aaba(a|bb*aa)* on random {a,b} strings of 64-4096 chars "
                "(alive random-walk set, last letter flipped for 50%)"),
    "config3": ("ex05_fwd", "{a*}:1c{&1}:2c(&1|&2)*", [],
                "config3: example 5 {a*}:1c{&1}:2c(&1|&2)* on x c x c x^m strings of 64-4096 chars, "
                "50% corrupted in the last block"),
    "config4": ("ex02_fwd+ex02_rev", "{(a|bb)*}:1aaba(&1|bb*aa)*", [],
                "config4: example 2 {(a|bb)*}:1aaba(&1|bb*aa)* forward AND -reverse tables on attack strings "
                "(bbaa)^k aaba (bbaa)^k [c], lengths log-uniform 435-65536"),
    "config5": ("ex01..ex10 forward", "README examples 1-10", [],
                "config5: all 10 README examples, per GPU 100k strings each (pumped / near-miss / random, "
                "16-512 chars), one matcher call per example per step"),
}


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_workload(W, name, table_text, n, seed, device, uniform=False):
    if name == "config2":
        if uniform:
            return W.uniform_strings(n, 64, 4096, b"ab", seed, device)
        return W.alive_strings(table_text, n, 64, 4096, seed, device)
    if name == "config3":
        return W.example5_strings(n, 64, 4096, seed, device)
    raise SystemExit(f"unknown workload {name}")


# --------------------------------------------------------------------------------------
# reference CPU implementation (oracle/_ref, or the C port when it is absent)
# --------------------------------------------------------------------------------------

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_parallel(binary, regex, flags, chars, offsets, procs):
    """`procs` processes of the reference's -match code, each on a contiguous shard.
    Returns (wall seconds of the slowest match loop, bits)."""
    n = len(offsets) - 1
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "in.rxmb")
        with open(path, "wb") as f:
            f.write(b"RXMBATCH")
            f.write(np.uint64(n).tobytes())
            f.write(np.uint64(len(chars)).tobytes())
            f.write(np.ascontiguousarray(offsets, dtype=np.uint64).tobytes())
            f.write(np.ascontiguousarray(chars, dtype=np.uint8).tobytes())
        per = (n + procs - 1) // procs
        jobs = []
        for p in range(procs):
            lo, hi = min(n, p * per), min(n, (p + 1) * per)
            if lo >= hi:
                continue
            wd = os.path.join(td, f"w{p}")
            os.makedirs(wd)
            out = os.path.join(wd, "out.bits")
            cmd = [binary, "-match", *flags, "-regex", regex, "-batch", path, out,
                   "-range", str(lo), str(hi)]
            jobs.append((lo, hi, out, subprocess.Popen(cmd, cwd=wd, stdout=subprocess.DEVNULL,
                                                       stderr=subprocess.PIPE)))
        bits = np.zeros(n, dtype=np.uint8)
        slowest = 0.0
        for lo, hi, out, pr in jobs:
            _, err = pr.communicate()
            if pr.returncode != 0:
                raise RuntimeError(f"reference process failed: {err.decode()[-300:]}")
            for line in err.decode().splitlines():
                if line.startswith("ORACLE_TIME"):
                    slowest = max(slowest, float(line.split()[1]))
            bits[lo:hi] = np.fromfile(out, dtype=np.uint8)
        return slowest, bits


def run_port(tables, chars, offsets):
    """The C restatement (oracle/librxm_oracle.so), one thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    t0 = time.perf_counter()
    bits = H.oracle_bits(tables, chars, offsets)
    return time.perf_counter() - t0, bits


def cpu_sample(W, wl, table_text, per_core, cores, seed, dev="cpu"):
    """A bounded sample of the workload for the CPU legs; generated on the CPU unless the caller (the GPU
    arm's cpu_baseline leg, which owns a device anyway) says otherwise."""
    chars, offsets = make_workload(W, wl, table_text, per_core * cores, seed, dev)
    return chars.cpu().numpy(), offsets.cpu().numpy().astype(np.uint64)


def reference_binary():
    for b in ("diploma_ref_O2", "diploma_ref"):
        p = os.path.join(REF_DIR, b)
        if os.path.exists(p):
            return p, b
    return None, None


def cpu_baseline(W, rxm, wl, tables, regex, flags, gpu_bits=None):
    """Reference -match code on the host cores, on a bounded sample of the workload.
    gpu_bits(chars, offsets) -> the device's bits for the same sample (SURVEY 7 policy: gate on the
    canonical allocation-order build, also report the stock-glibc build)."""
    cores = host_cores()
    res = {}
    binary, bname = reference_binary()
    if binary:
        per_core = 1500 if wl == "config2" else 60
        chars, offsets = cpu_sample(W, wl, tables.text, per_core, cores, 4242, "cuda")
        n = len(offsets) - 1
        t_all, bits_stock = run_reference_parallel(binary, regex, flags, chars, offsets, cores)
        one = slice(0, per_core + 1)
        o1 = offsets[one]
        t_one, _ = run_reference_parallel(binary, regex, flags, chars[:int(o1[-1])], o1, 1)
        res = {
            "value": n / t_all, "unit": "strings/s", "cores": cores, "kind": "reference",
            "build": bname + " (unmodified reference sources, -O2; upstream CMake sets no -O flag)",
            "sample": f"{n} strings ({int(offsets[-1])} bytes) of the same workload, "
                      f"{cores} processes x {per_core} strings, string-parallel",
            "input_mb_s": float(offsets[-1]) / t_all / 1e6,
            "single_thread": {"value": per_core / t_one, "unit": "strings/s", "cores": 1,
                              "sample": f"{per_core} strings"},
        }
        o0 = os.path.join(REF_DIR, "diploma_ref")
        if os.path.exists(o0) and bname != "diploma_ref":
            k = max(1, per_core // 8)
            ok = offsets[:k + 1]
            t0, _ = run_reference_parallel(o0, regex, flags, chars[:int(ok[-1])], ok, 1)
            res["single_thread_O0_upstream_flags"] = {"value": k / t0, "unit": "strings/s",
                                                       "cores": 1, "sample": f"{k} strings"}
            # upstream's own flags (CMakeLists.txt:5-6 sets no -O), string-parallel over every core
            kp = max(1, per_core // 8) * cores
            okp = offsets[:kp + 1]
            t0p, _ = run_reference_parallel(o0, regex, flags, chars[:int(okp[-1])], okp, cores)
            res["string_parallel_O0_upstream_flags"] = {"value": kp / t0p, "unit": "strings/s", "cores": cores,
                                                         "sample": f"{kp} strings, {cores} processes"}
        bump = os.path.join(REF_DIR, "diploma_ref_bump_O2")
        if os.path.exists(bump):
            # the gate is the canonical (allocation-order) build; the stock-glibc build is reported
            _, bits_bump = run_reference_parallel(bump, regex, flags, chars, offsets, cores)
            dis = {"strings": int(n), "stock_vs_canonical": int((bits_stock != bits_bump).sum()),
                   "note": "stock = diploma_ref_O2 (glibc malloc, heap-order tie-breaks), canonical = "
                           "diploma_ref_bump_O2 (never-reuse allocator); the device is gated on canonical"}
            if gpu_bits is not None:
                got = gpu_bits(chars, offsets)
                dis["device_vs_canonical"] = int((got != bits_bump).sum())
                dis["device_vs_stock"] = int((got != bits_stock).sum())
            res["stock_glibc_disagreement"] = dis
    else:
        per = 20000 if wl == "config2" else 2000
        chars, offsets = cpu_sample(W, wl, tables.text, per, 1, 4242)
        t, _ = run_port(tables, chars, offsets)
        res = {"value": per / t, "unit": "strings/s", "cores": 1, "kind": "port",
               "sample": f"{per} strings ({int(offsets[-1])} bytes) of the same workload, "
                         "C restatement oracle/rxm_oracle.c, one thread",
               "input_mb_s": float(offsets[-1]) / t / 1e6}
    return res


def cpu_baseline_jobs(jobs, wl):
    """configs 4/5: the reference on a bounded sample of every job (string-parallel)."""
    cores = host_cores()
    binary, bname = reference_binary()
    if not binary:
        return {"value": None, "unit": "strings/s", "cores": 0, "kind": "port",
                "sample": "oracle/_ref not built on this box"}
    tot_n, tot_t, tot_b = 0, 0.0, 0
    for j in jobs:
        off = j["offsets"].cpu().numpy().astype(np.uint64)
        lens = np.diff(off)
        # the reference is O(len^2) per string (mfa.cpp:136,203 copy the input per call):
        # only strings up to 8K chars are sampled, one per core (config 4) / 40 per core (config 5)
        idx = np.nonzero(lens <= 8192)[0][: cores * (1 if wl == "config4" else 40)]
        if len(idx) == 0:
            continue
        ch = j["chars"].cpu().numpy()
        parts = [ch[int(off[i]):int(off[i + 1])] for i in idx]
        so = np.zeros(len(idx) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in parts], out=so[1:])
        sc = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
        t, _ = run_reference_parallel(binary, j["regex"], j["flags"], sc, so, cores)
        tot_n += len(idx)
        tot_t += t
        tot_b += int(so[-1])
    return {"value": tot_n / tot_t if tot_t else None, "unit": "strings/s", "cores": cores, "kind": "reference",
            "build": bname + " (-O2 build of the unmodified reference sources)",
            "sample": f"{tot_n} strings ({tot_b} bytes) of the same workload with length <= 8192, "
                      f"string-parallel over {cores} processes, jobs timed one after another",
            "input_mb_s": tot_b / tot_t / 1e6 if tot_t else None}


def reference_arm(args, W, rxm):
    """--impl reference: the reference's CPU implementation, all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    if wl not in ("config2", "config3"):
        raise SystemExit("bench.py --impl reference: config2 / config3 (the multi-job workloads report their "
                         "reference figure in the GPU arm's cpu_baseline)")
    case, regex, flags, desc = WORKLOADS[wl]
    # this arm maps neither the product library nor the GPU: the table text is only needed by the string
    # generator (plain Python), the sample is made on the CPU, the timed region is the reference's own
    # processes
    table_text = open(os.path.join(CASES, case + ".rxt")).read()
    cores = host_cores()
    binary, bname = reference_binary()
    per_core = 1500 if wl == "config2" else 60
    chars, offsets = cpu_sample(W, wl, table_text, per_core, cores if binary else 1, 4242, "cpu")
    tables = None
    if not binary:  # no oracle/_ref on this box: the C restatement needs the parsed table (librxm's parser)
        tables = rxm.Tables(table_text)
    n = len(offsets) - 1
    times = []
    for step in range(args.warmup + args.steps):
        if binary:
            t, _ = run_reference_parallel(binary, regex, flags, chars, offsets, cores)
        else:
            t, _ = run_port(tables, chars, offsets)
        if step >= args.warmup:
            times.append(t)
    total = sum(times)
    value = n * len(times) / total
    line = {
        "impl": "reference", "metric": "strings_per_sec", "value": value, "unit": "strings/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "strings_per_step": n, "bytes_per_step": int(offsets[-1])},
        "input_gb_s": float(offsets[-1]) * len(times) / total / 1e9,
        "cpu_baseline": {"value": value, "unit": "strings/s",
                         "cores": cores if binary else 1,
                         "kind": "reference" if binary else "port",
                         "build": (bname + " (-O2 build of the unmodified reference sources)") if binary
                                  else "oracle/rxm_oracle.c",
                         "sample": f"{n} strings per step, string-parallel over "
                                   f"{cores if binary else 1} processes"},
        "e2e": {"value": value, "unit": "strings/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(gpu_id)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path).read().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                power.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        # samples under load = the upper half by SM clock (idle samples bracket the region)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


RESULT_OUT = sys.stdout
ORIG_AFFINITY = None


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off (what `numactl` would do for a
    one-process-per-GPU job): pinned host buffers are then allocated next to the GPU's PCIe root,
    which is what the host-buffer (`e2e`) path pays for when several ranks copy at once.  Returns what
    was found -- {"node", "pci", "host_numa_nodes", "bound_cpus", "why"} -- so that a record says WHY a
    rank was not bound (round 1's records only said null); the original affinity is kept for the
    CPU-baseline leg."""
    global ORIG_AFFINITY
    info = {"node": None, "pci": None, "host_numa_nodes": None, "bound_cpus": None, "why": None}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["host_numa_nodes"] = len(nodes)
    except OSError as e:
        info["why"] = f"/sys/devices/system/node: {e.strerror}"
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]  # 0000:1b:00.0
        info["pci"] = bdf
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            info["why"] = "the kernel reports numa_node -1 for the GPU's PCI device (no affinity exposed: a VM or a single-node host)"
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        ORIG_AFFINITY = os.sched_getaffinity(0)
        cpus &= ORIG_AFFINITY
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["node"] = node
            info["bound_cpus"] = len(cpus)
        else:
            info["why"] = f"node {node} has no CPU this process may run on"
    except Exception as e:  # noqa: BLE001 -- recorded, not fatal
        info["why"] = f"{type(e).__name__}: {e}"
    return info


# --------------------------------------------------------------------------------------
# GPU arm: jobs (one matcher call each per step), timing, parity spot check
# --------------------------------------------------------------------------------------

def build_jobs(W, rxm, wl, strings, seed, dev, local_rank, host=None):
    """The jobs of one workload on this rank: dicts with name, regex, flags, tables, chars, offsets (device
    tensors), n, bytes, out, matcher.  `host` = precomputed list of (name, regex, flags, chars, offsets) numpy
    batches (the sharded mode cuts one global batch); otherwise the batches are generated here."""
    import torch
    case, regex, flags, desc = WORKLOADS[wl]
    jobs = []
    if host is not None:
        for name, rgx, fl, c_np, o_np in host:
            jobs.append({"name": name, "regex": rgx, "flags": fl,
                         "tables": rxm.Tables.load(os.path.join(CASES, name + ".rxt")),
                         "chars": torch.from_numpy(np.ascontiguousarray(c_np)).to(dev),
                         "offsets": torch.from_numpy(o_np.astype(np.int64)).to(dev)})
    elif wl in ("config2", "config3"):
        tables = rxm.Tables.load(os.path.join(CASES, case + ".rxt"))
        ch, of = make_workload(W, wl, tables.text, strings, seed, dev)
        jobs.append({"name": case, "regex": regex, "flags": flags, "tables": tables, "chars": ch, "offsets": of})
    else:
        for name, rgx, fl, c_np, o_np in host_batches(W, wl, strings, seed):
            jobs.append({"name": name, "regex": rgx, "flags": fl,
                         "tables": rxm.Tables.load(os.path.join(CASES, name + ".rxt")),
                         "chars": torch.from_numpy(c_np).to(dev),
                         "offsets": torch.from_numpy(o_np.astype(np.int64)).to(dev)})
    for j in jobs:
        j["n"] = int(j["offsets"].numel() - 1)
        j["bytes"] = int(j["offsets"][-1]) if j["n"] else 0
        j["out"] = torch.empty(max(j["n"], 1), dtype=torch.uint8, device=dev)[:j["n"]]
        j["matcher"] = rxm.Matcher(j["tables"], local_rank)
    return jobs


def host_batches(W, wl, strings, seed):
    """configs 4 and 5 as numpy batches: (fixture name, regex, flags, chars, offsets) per job."""
    regex = WORKLOADS[wl][1]
    out = []
    if wl == "config4":
        n4 = strings if strings != 1_000_000 else 4096
        c_np, o_np = W.attack_batch(["bbaa", "aaba", "bbaa"], "c", "", n4, 435, 65536, seed)
        for cname, fl in (("ex02_fwd", []), ("ex02_rev", ["-reverse"])):
            out.append((cname, regex, fl, c_np, o_np))
    else:  # config5
        per = strings // 10
        for ex in range(1, 11):
            c_np, o_np = W.mixed_example_batch(ex, per, 1000 * ex + seed)
            out.append((f"ex{ex:02d}_fwd", W.README_EXAMPLES[ex][0], [], c_np, o_np))
    return out


class JobRunner:
    """One step = every job's rxm_match_batch once; several jobs run on streams of their own, forked from and
    joined back into the timed stream (handles may run concurrently, include/rxm.h)."""

    def __init__(self, jobs, dev, share):
        import torch
        self.torch = torch
        self.jobs = jobs
        for j in jobs:
            j["matcher"].set_concurrency(share)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.job_streams = [torch.cuda.Stream(device=dev) for _ in jobs] if len(jobs) > 1 else []
        self.fork_ev = torch.cuda.Event()
        self.join_evs = [torch.cuda.Event() for _ in self.job_streams]

    def step_device(self):
        torch = self.torch
        if not self.job_streams:
            j = self.jobs[0]
            j["matcher"].match_ptrs(j["chars"].data_ptr(), j["offsets"].data_ptr(), j["n"], j["out"].data_ptr(), self.stream)
            return
        cur = torch.cuda.current_stream()
        self.fork_ev.record(cur)
        for j, st, ev_j in zip(self.jobs, self.job_streams, self.join_evs):
            st.wait_event(self.fork_ev)
            j["matcher"].match_ptrs(j["chars"].data_ptr(), j["offsets"].data_ptr(), j["n"], j["out"].data_ptr(), st.cuda_stream)
            ev_j.record(st)
        for ev_j in self.join_evs:
            cur.wait_event(ev_j)

    def launch_count(self):
        return sum(j["matcher"].launch_count() for j in self.jobs)

    def overflow_count(self):
        return sum(j["matcher"].overflow_count() for j in self.jobs)

    def close(self):
        for j in self.jobs:
            j["matcher"].close()

    def time_steps(self, steps, total_bytes, dev):
        """K steps with a CUDA event between them.  Inputs smaller than twice the 126 MB L2 get a 512 MB buffer
        written between the timed steps (outside every per-step event pair); larger inputs evict themselves.
        -> (per-step ms list, total ms over the K steps as the device saw them, flushed?)"""
        torch = self.torch
        flush = total_bytes < 2 * 126 * (1 << 20)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        if flush:
            flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
            ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
            for k in range(steps):
                flush_buf.fill_(k & 0xff)
                ev0[k].record()
                self.step_device()
                ev[k + 1].record()
            torch.cuda.synchronize()
            step_ms = [ev0[k].elapsed_time(ev[k + 1]) for k in range(steps)]
            return step_ms, sum(step_ms), True
        ev[0].record()
        for k in range(steps):
            self.step_device()
            ev[k + 1].record()
        torch.cuda.synchronize()
        step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        return step_ms, ev[0].elapsed_time(ev[steps]), False


def parity_spot_check(jobs, H):
    """Outside every timed region: the device's bits against the C restatement on a sample that spans the WHOLE
    length range of every job -- evenly spaced picks from the strings ordered by length, the longest included.
    Returns the number of strings checked; raises SystemExit on any difference."""
    checked = 0
    for j in jobs:
        if j["n"] == 0:
            continue
        off = j["offsets"].cpu().numpy().astype(np.uint64)
        lens = np.diff(off)
        long_strings = j["bytes"] / max(1, j["n"]) >= 5000
        k = min(j["n"], 96 if long_strings else 3000)
        order = np.argsort(lens, kind="stable")
        pick = np.unique(order[np.linspace(0, j["n"] - 1, k).astype(np.int64)])
        ch = j["chars"].cpu().numpy()
        parts = [ch[int(off[i]):int(off[i + 1])] for i in pick]
        so = np.zeros(len(pick) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in parts], out=so[1:])
        sc = np.concatenate(parts) if int(so[-1]) else np.zeros(0, dtype=np.uint8)
        want = H.oracle_bits(j["tables"], sc, so)
        got = j["out"].cpu().numpy()[pick]
        if not np.array_equal(got, want):
            bad = pick[np.nonzero(got != want)[0][:5]]
            raise SystemExit(f"bench.py: {j['name']}: {int((got != want).sum())} of {len(pick)} bits differ from the oracle "
                             f"(strings {bad.tolist()}, lengths {lens[bad].tolist()})")
        checked += len(pick)
    return checked


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def engine_of(rxm, j):
    """The kernel family that runs job j: the planner's choice for the automaton -- except that batches of long
    strings (mean length above 4096) of a K4 automaton are K3's, 32 lanes per string (picked per batch, rxm_api.cu)."""
    name = rxm.ENGINE_NAMES.get(j["matcher"].plan().engine, "?")
    if name == "K4_THREAD" and j["n"] and j["bytes"] / j["n"] > 4096:
        return "K3_WARP(32 lanes; long strings)"
    return name


def extra_workload(W, rxm, H, wl, strings, steps, dev, local_rank, seed):
    """A short measurement of another BASELINE config inside the default run: the same
    timing rules as the headline (warm-up, CUDA events, L2 flush for small inputs, parity spot check)."""
    import torch
    jobs = build_jobs(W, rxm, wl, strings, seed, dev, local_rank)
    share = len(jobs) if wl == "config4" else 1
    R = JobRunner(jobs, dev, share)
    n = sum(j["n"] for j in jobs)
    total_bytes = sum(j["bytes"] for j in jobs)
    for _ in range(3):
        R.step_device()
    torch.cuda.synchronize()
    l0 = R.launch_count()
    step_ms, total_ms, flushed = R.time_steps(steps, total_bytes, dev)
    launches = R.launch_count() - l0
    if R.overflow_count():
        raise SystemExit(f"bench.py: {wl}: strings hit a kernel limit")
    checked = parity_spot_check(jobs, H)
    k_ms = statistics.mean(step_ms)
    peak, _ = hbm_peak()
    algo = total_bytes + 9 * n
    res = {
        "workload": WORKLOADS[wl][3], "strings": n, "bytes": total_bytes, "jobs_per_step": len(jobs), "steps": steps,
        "ms_per_step": k_ms, "strings_per_sec": n / (k_ms / 1e3), "input_gb_s": total_bytes / (k_ms / 1e3) / 1e9,
        "roofline_frac": algo / (k_ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_step": algo,
        "engine": "+".join(sorted({engine_of(rxm, j) for j in jobs})),
        "gpu_launches": int(launches), "l2_flushed_between_steps": flushed, "parity_checked": checked,
        "match_fraction": float(sum(float(j["out"].float().sum().item()) for j in jobs) / max(1, n)),
    }
    # the same call end to end: pinned HOST buffers, H2D of the strings + offsets and D2H of the bits inside the step
    stream = torch.cuda.current_stream().cuda_stream
    for j in jobs:
        j["h_chars"] = torch.empty(j["bytes"], dtype=torch.uint8, pin_memory=True)
        j["h_chars"].copy_(j["chars"][:j["bytes"]])
        j["h_off"] = torch.empty(j["n"] + 1, dtype=torch.int64, pin_memory=True)
        j["h_off"].copy_(j["offsets"])
        j["h_out"] = torch.empty(j["n"], dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def step_host():
        for j in jobs:
            j["matcher"].match_ptrs(j["h_chars"].data_ptr(), j["h_off"].data_ptr(), j["n"], j["h_out"].data_ptr(), stream)

    step_host()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    same = all(bool(torch.equal(j["h_out"], j["out"].cpu())) for j in jobs)
    if not same:
        raise SystemExit(f"bench.py: {wl}: host route and device route disagree")
    res["e2e"] = {"ms_per_step": e2e_ms, "value": n / (e2e_ms / 1e3), "unit": "strings/s",
                  "h2d_bytes_per_step": int(total_bytes + 8 * (n + len(jobs))), "d2h_bytes_per_step": int(n),
                  "note": "rxm_match_batch with pinned host buffers: copy in, kernels, copy out, one job after another"}
    eff = os.path.join(ROOT, "profiles", "mfa_kernel_efficiency.json")  # from the committed ncu captures
    if os.path.exists(eff):
        res["ncu"] = json.load(open(eff)).get(wl)
    R.close()
    del jobs, R
    torch.cuda.empty_cache()
    return res


def large_table_workloads(rxm, H, dev, steps=3, n=200_000):
    """Memory-free automata with large determinisations, briefly, in the default run (not a BASELINE config; the
    planner's regimes beyond config 2's 13 sets): 386 sets (two-lookup table with its four-byte stride), 24 577 sets
    (the largest table K1 holds in shared memory), ~98 000 sets (no table: the bit-set engine K1B).  200 k random
    {a, b} strings of 64-4096 letters, device-resident, CUDA events on the launching stream; parity spot check against
    the C restatement outside the timed region."""
    import torch
    g = torch.Generator(device=dev).manual_seed(1)
    lens = torch.randint(64, 4097, (n,), device=dev, generator=g)
    off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    total = int(off[-1])
    chars = torch.empty(total + 64, dtype=torch.uint8, device=dev)
    chars[:total] = (torch.randint(0, 2, (total,), device=dev, generator=g) + 97).to(torch.uint8)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    peak, _ = hbm_peak()
    res = []
    for case in ("nfa_mid", "nfa_blowup", "nfa_huge"):
        t = rxm.Tables.load(os.path.join(ROOT, "tests", "golden", "cases", case + ".rxt"))
        m = rxm.Matcher(t, dev.index or 0)
        p = m.plan()
        for _ in range(2):
            m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            m.match_ptrs(chars.data_ptr(), off.data_ptr(), n, out.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        job = {"name": case, "n": n, "bytes": total, "offsets": off, "chars": chars, "out": out, "tables": t}
        checked = parity_spot_check([job], H)
        res.append({"automaton": case, "engine": rxm.ENGINE_NAMES.get(p.engine, "?"), "active_sets": int(p.dfa_states),
                    "bytes_per_lookup": int(p.dfa_stride), "strings": n, "bytes": total, "steps": steps, "ms_per_step": ms,
                    "strings_per_sec": n / (ms / 1e3), "input_gb_s": total / (ms / 1e3) / 1e9,
                    "roofline_frac": (total + 9 * n) / (ms / 1e3) / 1e9 / peak, "parity_checked": checked,
                    "match_fraction": float(out.float().mean().item())})
        m.close()
    del chars, off, out
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--strings", type=int, default=1_000_000, help="strings per GPU (weak) / in the whole batch (strong)")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank owns a batch of --strings; strong: ONE batch of --strings is cut by bytes "
                         "over the ranks (sharding.py), the result bits are gathered over NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-workloads", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL prints its version
    # line to stdout at the first collective) is sent to stderr instead
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rxm = _load("rxm", os.path.join(PKG, "rxm.py"))  # ctypes declarations only: librxm.so is mapped on first use
    W = _load("workloads", os.path.join(PKG, "workloads.py"))
    if args.impl == "reference":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""  # the reference arm does not touch a GPU
        return reference_arm(args, W, rxm)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H  # the oracle (checker), outside every timed region

    wl = args.workload
    case, regex, flags, desc = WORKLOADS[wl]
    strong = args.scaling == "strong"
    sharded = None
    if strong:
        # ONE batch (every rank builds the same one from the same seed -- what reading one shared input would
        # give), cut into byte-balanced contiguous string ranges, one per rank (sharding.py)
        S = _load("sharding", os.path.join(PKG, "sharding.py"))
        if wl in ("config2", "config3"):
            tbl = rxm.Tables.load(os.path.join(CASES, case + ".rxt"))
            ch, of = make_workload(W, wl, tbl.text, args.strings, 1000, dev)
            whole = [(case, regex, flags, ch.cpu().numpy(), of.cpu().numpy().astype(np.uint64))]
            del ch, of
        else:
            whole = host_batches(W, wl, args.strings, 0)
        sharded = []
        mine = []
        for name, rgx, fl, c_np, o_np in whole:
            bounds = S.shard_by_bytes(o_np, world)
            lc, lo = S.local_view(c_np, o_np, bounds[rank], bounds[rank + 1])
            mine.append((name, rgx, fl, lc, lo))
            sharded.append({"name": name, "bounds": bounds, "n": len(o_np) - 1})
        jobs = build_jobs(W, rxm, wl, args.strings, 0, dev, local_rank, host=mine)
    else:
        jobs = build_jobs(W, rxm, wl, args.strings, 1000 + rank, dev, local_rank)
    # config 4 is bounded by its longest string in each of the two automata: the handles share the
    # device (rxm_set_concurrency) so that the two launches run side by side; config 5's ten jobs are
    # throughput-bound and stay at one launch filling the device after another
    share = int(os.environ.get("RXM_BENCH_SHARE", len(jobs) if wl == "config4" else 1))
    R = JobRunner(jobs, dev, share)
    n = sum(j["n"] for j in jobs)
    total_bytes = sum(j["bytes"] for j in jobs)
    tables, chars, offsets, out, m = (jobs[0][k] for k in ("tables", "chars", "offsets", "out", "matcher"))
    stream = R.stream
    step_device = R.step_device

    gathered = []

    def step_sharded():
        """strong scaling: the match on this rank's slice, then the gather of the result bits and the sum of
        the match counts over NCCL (sharding.gather_bits / total_matches) -- the whole job's result on every rank"""
        step_device()
        gathered.clear()
        for j, sh in zip(jobs, sharded):
            gathered.append(S.gather_bits(j["out"], sh["bounds"], rank, world, dist, device=dev) if world > 1
                            else j["out"])
        tot = torch.stack([g.sum(dtype=torch.int64) for g in gathered]).sum()
        return tot

    # ---- kernel-resident timing (inputs already in HBM) ---------------------------------
    for _ in range(args.warmup):
        step_sharded() if strong else step_device()
    barrier()
    props = torch.cuda.get_device_properties(local_rank)
    gpu_id = getattr(props, "uuid", None)
    gpu_id = f"GPU-{gpu_id}" if gpu_id and not str(gpu_id).startswith("GPU-") else (gpu_id or local_rank)
    sampler = ClockSampler(gpu_id)
    # the timed region is padded so nvidia-smi (100 ms period) sees it
    t_pad = time.perf_counter()
    while time.perf_counter() - t_pad < 0.5:
        step_device()
        torch.cuda.synchronize()
    launches0 = R.launch_count()
    torch.cuda.cudart().cudaProfilerStart()  # `ncu --profile-from-start off`: exactly the timed steps' launches
    if strong:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            total_matches = step_sharded()
        e1.record()
        barrier()
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        step_ms = [e0.elapsed_time(e1) / args.steps] * args.steps
        flush = False
        total_matches = int(total_matches.item())
    else:
        step_ms, local_total, flush = R.time_steps(args.steps, total_bytes, dev)
        barrier()
        total_ms = max_over_ranks(local_total)
    torch.cuda.cudart().cudaProfilerStop()
    launches = R.launch_count() - launches0
    while time.perf_counter() - t_pad < 1.2:
        step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    n_all = sum_over_ranks(float(n))
    bytes_all = sum_over_ranks(float(total_bytes))
    value = n_all / (ms_per_step / 1e3)
    if R.overflow_count():
        raise SystemExit("bench.py: strings hit a kernel limit")

    # ---- parity spot check against the oracle (outside every timed region) ---------------
    k_chk = parity_spot_check(jobs, H)
    match_frac = float(sum(float(j["out"].float().sum().item()) for j in jobs) / max(1, n))
    extra = {}
    if strong:
        # the gathered vector (every rank holds it) against this rank's own bits and, on rank 0, against the
        # bits of the WHOLE batch matched on one GPU (what N = 1 gives)
        ok = True
        for j, sh, g in zip(jobs, sharded, gathered):
            lo_, hi_ = sh["bounds"][rank], sh["bounds"][rank + 1]
            ok = ok and bool(torch.equal(g[lo_:hi_], j["out"]))
        whole_equal = None
        if rank == 0 and world > 1:
            whole_equal = True
            for (name, rgx, fl, c_np, o_np), g in zip(whole, gathered):
                mt = rxm.Matcher(rxm.Tables.load(os.path.join(CASES, name + ".rxt")), local_rank)
                dc = torch.from_numpy(np.ascontiguousarray(c_np)).to(dev)
                do = torch.from_numpy(o_np.astype(np.int64)).to(dev)
                o1 = torch.empty(len(o_np) - 1, dtype=torch.uint8, device=dev)
                mt.match_ptrs(dc.data_ptr(), do.data_ptr(), len(o_np) - 1, o1.data_ptr(), stream)
                torch.cuda.synchronize()
                whole_equal = whole_equal and bool(torch.equal(o1, g))
                mt.close()
                del dc, do, o1
        if not ok or whole_equal is False:
            raise SystemExit("bench.py: gathered result vector differs from the single-GPU bits")
        extra["sharded"] = {
            "note": "ONE batch cut by bytes over the ranks (sharding.shard_by_bytes); the timed step is the match of "
                    "each rank's slice + all_gather of the result bits + sum of match counts over NCCL",
            "strings_total": int(sum(sh["n"] for sh in sharded)), "total_matches": total_matches,
            "bounds": {sh["name"]: [int(b) for b in sh["bounds"]] for sh in sharded} if world <= 8 else None,
            "gathered_equals_local": ok, "gathered_equals_single_gpu": whole_equal,
        }

    # ---- end to end through the C ABI with HOST buffers -----------------------------------
    for j in jobs:
        j["h_chars"] = torch.empty(j["bytes"], dtype=torch.uint8, pin_memory=True)
        j["h_chars"].copy_(j["chars"][:j["bytes"]])
        j["h_off"] = torch.empty(j["n"] + 1, dtype=torch.int64, pin_memory=True)
        j["h_off"].copy_(j["offsets"])
        j["h_out"] = torch.empty(j["n"], dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def step_host():
        for j in jobs:
            j["matcher"].match_ptrs(j["h_chars"].data_ptr(), j["h_off"].data_ptr(), j["n"],
                                    j["h_out"].data_ptr(), stream)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    for j in jobs:
        if not torch.equal(j["h_out"], j["out"].cpu()):
            raise SystemExit("bench.py: host-buffer path and device-pointer path disagree")
    e2e_value = n_all * e2e_steps / e2e_s

    # ---- host -> device copies alone: each rank by itself, then all ranks at once ----
    h2d = {}
    if jobs and jobs[0]["bytes"] > (64 << 20):
        hb, db = jobs[0]["h_chars"], jobs[0]["chars"][:jobs[0]["bytes"]]

        def copy_ms(reps=3):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                db.copy_(hb, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        copy_ms(1)
        alone = []
        for r in range(world):  # one rank at a time
            barrier()
            alone.append(copy_ms() if r == rank else 0.0)
            barrier()
        barrier()
        together = copy_ms()
        barrier()
        gb = jobs[0]["bytes"] / 1e9
        mine_alone = max(alone)
        h2d = {"bytes": jobs[0]["bytes"],
               "alone_gb_s": sum_over_ranks(gb / (mine_alone / 1e3)) / world,
               "together_gb_s_per_gpu": sum_over_ranks(gb / (together / 1e3)) / world,
               "together_gb_s_slowest_gpu": gb / (max_over_ranks(together) / 1e3),
               "note": "pinned host buffer -> device, cudaMemcpyAsync of the whole chars array; alone = one rank "
                       "copying while the others wait, together = all ranks at once (what the e2e step pays)"}

    # ---- uniform i.i.d. variant of config 2 (early exit) -----------------------------------
    if wl == "config2" and not strong:
        del jobs[0]["h_chars"]
        u_chars, u_off = make_workload(W, wl, tables.text, n, 2000 + rank, dev, uniform=True)
        for _ in range(3):
            m.match_ptrs(u_chars.data_ptr(), u_off.data_ptr(), n, out.data_ptr(), stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            m.match_ptrs(u_chars.data_ptr(), u_off.data_ptr(), n, out.data_ptr(), stream)
        e1.record()
        barrier()
        u_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        extra["uniform_iid"] = {
            "note": "i.i.d. uniform {a,b} strings: the active set dies within a few letters "
                    "(automata.cpp:186-188), so this times the early exit, not a scan",
            "ms_per_step": u_ms, "strings_per_sec": n_all / (u_ms / 1e3),
            "nominal_input_gb_s": sum_over_ranks(float(int(u_off[-1]))) / (u_ms / 1e3) / 1e9,
            "match_fraction": float(out.float().mean().item()),
        }
        del u_chars, u_off

    # ---- raw-text route (config 2): tokenise on the device, then match ------------------------
    if wl == "config2" and not strong:
        nt = min(n, 250_000)
        t_off = jobs[0]["offsets"][:nt + 1]
        t_bytes = int(t_off[-1])
        lens = (t_off[1:] - t_off[:-1])
        text = torch.full((t_bytes + nt,), 10, dtype=torch.uint8, device=dev)  # '\n' after every string
        idx = torch.arange(t_bytes, device=dev, dtype=torch.int64)
        idx += torch.repeat_interleave(torch.arange(nt, device=dev, dtype=torch.int64), lens)
        text[idx] = jobs[0]["chars"][:t_bytes]
        del idx
        t_out = torch.empty(nt, dtype=torch.uint8, device=dev)
        for _ in range(3):
            got_n = m.match_text_ptrs(text.data_ptr(), text.numel(), t_out.data_ptr(), nt, stream)
        torch.cuda.synchronize()
        t_ref = torch.empty(nt, dtype=torch.uint8, device=dev)
        m.match_ptrs(jobs[0]["chars"].data_ptr(), jobs[0]["offsets"].data_ptr(), nt, t_ref.data_ptr(), stream)
        torch.cuda.synchronize()
        if got_n != nt or not torch.equal(t_out, t_ref):
            raise SystemExit("bench.py: raw-text route disagrees with the offsets route")
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            m.match_text_ptrs(text.data_ptr(), text.numel(), t_out.data_ptr(), nt, stream)
        e1.record()
        barrier()
        t_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        h_text = torch.empty(text.numel(), dtype=torch.uint8, pin_memory=True)
        h_text.copy_(text)
        h_tout = torch.empty(nt, dtype=torch.uint8, pin_memory=True)
        torch.cuda.synchronize()
        m.match_text_ptrs(h_text.data_ptr(), h_text.numel(), h_tout.data_ptr(), nt, stream)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            m.match_text_ptrs(h_text.data_ptr(), h_text.numel(), h_tout.data_ptr(), nt, stream)
        th_s = (time.perf_counter() - t0) / e2e_steps
        if not torch.equal(h_tout, t_out.cpu()):
            raise SystemExit("bench.py: raw-text host route disagrees with the device route")
        extra["text_route"] = {
            "note": "rxm_match_text: newline-separated raw text, tokenised on the device "
                    "(match.cpp:22-24 semantics), then the same kernels; includes one stream "
                    "synchronisation per call (token count to the host)",
            "strings": nt, "text_bytes": int(text.numel()),
            "device_ms_per_step": t_ms, "device_strings_per_sec": nt / (t_ms / 1e3),
            "device_input_gb_s": text.numel() / (t_ms / 1e3) / 1e9,
            "host_ms_per_step": th_s * 1e3, "host_strings_per_sec": nt / th_s,
            "host_input_gb_s": text.numel() / th_s / 1e9,
        }
        del text, h_text

    engines = "+".join(sorted({engine_of(rxm, j) for j in jobs}))
    dfa_stride = int(jobs[0]["matcher"].plan().dfa_stride)
    n_jobs, n_streams = len(jobs), max(1, len(R.job_streams))

    def device_bits(chars_np, offsets_np):  # the device's answer for the CPU baseline's sample (same handle)
        return m.match_host(chars_np, offsets_np)

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        algo_bytes = total_bytes + 9 * n  # per launch on this rank: len + 8 B offset + 1 B result
        k_ms = statistics.mean(step_ms)
        achieved = algo_bytes / (k_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", f"traffic_{wl}.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": "strings_per_sec", "value": value, "unit": "strings/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc, "strings_per_gpu": n, "bytes_per_gpu": total_bytes,
                       "mean_len": total_bytes / max(1, n), "match_fraction": match_frac,
                       "engine": engines, "dfa_stride": dfa_stride, "jobs_per_step": n_jobs,
                       "handles_sharing_device": share, "job_streams": n_streams,
                       "numa": numa,
                       "l2": ("L2 flushed between timed steps (512 MB written; inputs are %.2f GB per GPU)" if flush else
                              "inputs (%.2f GB per GPU) are larger than the 126 MB L2") % (total_bytes / 1e9),
                       "sharding": ("ONE batch of %d strings cut by bytes over the ranks, result bits gathered over NCCL"
                                    % args.strings) if strong else
                                   "by string index, one rank per GPU, no data-path collective"},
            "input_gb_s": bytes_all / (ms_per_step / 1e3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "frac_of_nominal_8_tb_s": achieved / 8000.0},  # both denominators
            "e2e": {"value": e2e_value, "unit": "strings/s",
                    "h2d_bytes_per_step": total_bytes + 8 * (n + n_jobs), "d2h_bytes_per_step": n + 8 * n_jobs,
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "input_gb_s": bytes_all * e2e_steps / e2e_s / 1e9, "h2d_copy_alone_vs_together": h2d},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_checked": k_chk,
        }
        line.update(extra)
        if ORIG_AFFINITY is not None:
            os.sched_setaffinity(0, ORIG_AFFINITY)  # the CPU baseline uses every host core
        if world == 1 and not args.no_cpu_baseline:
            if wl in ("config2", "config3"):
                line["cpu_baseline"] = cpu_baseline(W, rxm, wl, tables, regex, flags, device_bits)
            else:
                line["cpu_baseline"] = cpu_baseline_jobs(jobs, wl)
    R.close()
    for j in jobs:
        j.clear()
    del jobs, chars, offsets, out
    torch.cuda.empty_cache()
    if rank == 0:
        # ---- the other BASELINE configs, briefly, in the default run (config 2 stays the headline) ----
        if world == 1 and wl == "config2" and not strong and not args.no_extra_workloads:
            wk = {}
            for xwl, xs, xn in (("config3", 3, 1_000_000), ("config4", 2, 4096), ("config5", 3, 1_000_000)):
                wk[xwl] = extra_workload(W, rxm, H, xwl, xn, xs, dev, local_rank, 1000)
            try:  # (an extra: a failure here must not take the headline line with it; a parity difference still does)
                wk["memory_free_large_tables"] = large_table_workloads(rxm, H, dev)
            except Exception as e:  # noqa: BLE001
                wk["memory_free_large_tables"] = {"error": repr(e)}
            line["workloads"] = wk
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
