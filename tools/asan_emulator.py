"""Bounds check of the KERNEL SOURCES in place of `compute-sanitizer --tool memcheck` (closed on this pool):
tests/hostsim -- the kernels and launch functions of rxm_k1.cu, rxm_k1b.cu, rxm_k2.cu, rxm_k3.cu, rxm_k4.cu and
rxm_tok.cu compiled for the host under the SIMT emulator -- built with AddressSanitizer, every batch held in a
heap block of EXACTLY the bytes include/rxm.h lets the kernels read (the strings, from their first byte rounded
down to 32 bytes to their last rounded up to 32), so that a read or write outside the contract, in global or in
(emulated) shared memory, is an ASan report.  Run by tools/asan_emulator.sh (needs LD_PRELOAD of libasan)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from cases import BY_NAME, CASE_NAMES, load_case  # noqa: E402

L = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim_asan.so"))
libc = C.CDLL(None)
libc.posix_memalign.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t]
libc.free.argtypes = [C.c_void_p]


def exact(chars, shift):
    """the batch in a 32-aligned heap block: `shift` junk bytes (the odd alignment of a sub-buffer), the strings,
    and nothing but the round-up to 32 behind them -> (block, pointer to the first string)"""
    size = (shift + len(chars) + 31) & ~31
    if os.environ.get("RXM_ASAN_SELFTEST"):  # a block 32 bytes short of the contract: ASan must report the scan's last read
        size -= 32
    p = C.c_void_p()
    assert libc.posix_memalign(C.byref(p), 32, max(size, 32)) == 0
    C.memset(p, 0x5a, max(size, 32))
    if len(chars):
        C.memmove(p.value + shift, chars.ctypes.data, len(chars))
    return p, p.value + shift


def run(fn, t, strings, *extra, shift=0):
    chars, off = H.make_batch(strings)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    blk, ptr = exact(chars, shift)
    got = np.full(len(strings), 7, dtype=np.uint8)
    msg = C.create_string_buffer(600)
    rc = fn(C.cast(t.ptr, C.c_void_p), C.c_void_p(ptr), C.c_void_p(off.ctypes.data), C.c_uint64(len(strings)),
            C.c_void_p(got.ctypes.data), *extra, msg, C.c_uint32(600), C.c_uint64(3))
    libc.free(blk)
    return rc, got, msg.value.decode()


def main():
    rng = np.random.default_rng(5)
    runs = 0
    for name in CASE_NAMES:
        t, strings, bits = load_case(name)
        kind = BY_NAME[name]["kind"]
        alpha = np.frombuffer(b"ab" if kind == "nfa" else b"aabbc", dtype=np.uint8)
        extra = [bytes(rng.choice(alpha, size=int(n))) for n in rng.integers(0, 700, size=120)]
        extra += [b"a" * 1500 + b"c" + b"a" * 1500 + b"c" + b"a" * 3000]  # long repeated blocks (K4's phase A, block compares)
        batch = strings + extra
        want = H.oracle_bits(t, *[np.ascontiguousarray(x) for x in H.make_batch(batch)])
        for shift in (0, 5, 17):
            if kind == "nfa":
                # K1B, follow masks: the string in aligned 16-byte vectors, forward or downwards from its end
                ovf = C.c_ulonglong(0)
                rc, got, msg = run(L.hostsim_k1b_batch, t, batch, C.c_int(0), None, C.c_uint64(400_000_000), C.byref(ovf),
                                   shift=shift)
                assert rc == 0 and np.array_equal(got, want), (name, "k1b", rc, msg)
                runs += 1
                if name == "nfa_huge":  # no table fits: K1B only
                    continue
                info = (C.c_uint32 * 3)()
                ovf = C.c_ulonglong(0)
                rc, got, msg = run(L.hostsim_k1_batch, t, batch, C.c_uint64(400_000_000), C.byref(ovf), info, shift=shift)
                assert rc == 0 and np.array_equal(got, want), (name, "k1", rc, msg)
                runs += 1
            else:
                ovf, redo = C.c_ulonglong(0), C.c_ulonglong(0)
                rc, got, msg = run(L.hostsim_k4_batch, t, batch, C.c_uint32(0), None, C.c_uint64(400_000_000),
                                   C.byref(ovf), C.byref(redo), shift=shift)
                assert rc == 0, (name, "k4", rc, msg)
                ok = (got == want) | (got == 2)  # 2: handed on to K3 (outgrew a thread's slots)
                assert ok.all() or redo.value > 0, (name, "k4", int((~ok).sum()))
                runs += 1
                rc, got, msg = run(L.hostsim_k3_batch, t, batch, C.c_uint32(8), None, C.c_uint64(400_000_000), C.byref(ovf),
                                   shift=shift)
                assert rc == 0 and np.array_equal(got, want), (name, "k3", rc, msg)
                runs += 1
        print(f"{name}: ok", flush=True)
    print(f"asan emulator: {runs} kernel-source runs on exact-size heap blocks, no report")


if __name__ == "__main__":
    main()
