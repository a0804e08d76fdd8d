"""Shared test helpers: batches, the oracle (C restatement), the reference binaries.

Everything under oracle/ is loaded ONLY from here (tests), from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "re2-modification_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_LIB = os.path.join(ROOT, "oracle", "librxm_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BUMP = os.path.join(REF_DIR, "diploma_ref_bump")
REF_STOCK = os.path.join(REF_DIR, "diploma_ref")
RXM_COMPILE = os.path.join(PKG, "bin", "rxm_compile")


def load_rxm():
    """Import re2-modification_b200/rxm.py (the directory name is not an identifier)."""
    spec = importlib.util.spec_from_file_location("rxm", os.path.join(PKG, "rxm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


rxm = load_rxm()


def load_workloads():
    spec = importlib.util.spec_from_file_location("workloads", os.path.join(PKG, "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(ORACLE_LIB)
        L.rxm_oracle_match.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.rxm_oracle_match.restype = C.c_int
        L.rxm_oracle_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_void_p]
        L.rxm_oracle_match_batch.restype = None
        _oracle = L
    return _oracle


def make_batch(strings):
    """list[bytes] -> (chars uint8[total], offsets uint64[n+1])."""
    lens = np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings))
    offsets = np.zeros(len(strings) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer(b"".join(strings), dtype=np.uint8).copy()
    return chars, offsets


def oracle_bits(tables, chars, offsets):
    n = len(offsets) - 1
    out = np.empty(n, dtype=np.uint8)
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    oracle().rxm_oracle_match_batch(C.cast(tables.ptr, C.c_void_p), chars.ctypes.data, offsets.ctypes.data, n,
                                    out.ctypes.data)
    return out


def write_batch_file(path, chars, offsets):
    with open(path, "wb") as f:
        f.write(b"RXMBATCH")
        f.write(np.uint64(len(offsets) - 1).tobytes())
        f.write(np.uint64(len(chars)).tobytes())
        f.write(np.ascontiguousarray(offsets, dtype=np.uint64).tobytes())
        f.write(np.ascontiguousarray(chars, dtype=np.uint8).tobytes())


def reference_bits(regex: str, flags, chars, offsets, binary=REF_BUMP, timeout=600):
    """Run the reference's own code (oracle/_ref) on a batch; returns uint8 bits."""
    with tempfile.TemporaryDirectory() as td:
        bin_in, bin_out = os.path.join(td, "in.rxmb"), os.path.join(td, "out.bits")
        write_batch_file(bin_in, chars, offsets)
        cmd = [binary, "-match", *flags, "-regex", regex, "-batch", bin_in, bin_out]
        subprocess.run(cmd, cwd=td, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, timeout=timeout)
        return np.fromfile(bin_out, dtype=np.uint8)


def compile_tables_text(regex: str, flags=()):
    """Front end (reference parse/compile + flattening stage) -> table text."""
    r = subprocess.run([RXM_COMPILE, "-match", *flags, "-regex", regex], check=True,
                       capture_output=True)
    return r.stdout.decode()


def have_reference():
    return os.path.exists(REF_BUMP) and os.path.exists(RXM_COMPILE)
