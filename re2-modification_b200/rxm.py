"""ctypes binding of the C ABI in include/rxm.h (librxm.so).

Python is only the harness language here (tests, bench.py, smoke): the product
is the C-ABI library and the C++ front end.  This module loads the in-tree
``librxm.so`` and fails loudly if it is missing -- there is no Python or CPU
implementation of the match path behind it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RXM_LIB") or os.path.join(_HERE, "librxm.so")  # RXM_LIB: tuning builds

RXM_OK = 0
RXM_ERR_INVALID = 1
RXM_ERR_UNSUPPORTED = 2
RXM_ERR_NO_DEVICE = 3
RXM_ERR_CUDA = 4
RXM_ERR_NOMEM = 5
RXM_ERR_PARSE = 6
RXM_ERR_OVERFLOW = 7

RXM_KIND_NFA, RXM_KIND_MFA = 0, 1
ENGINE_NAMES = {1: "K1_DFA", 2: "K1_BITSET", 3: "K2_THREAD", 4: "K3_WARP", 5: "K4_THREAD"}

# every symbol include/rxm.h declares (tests check that the library exports them all)
ABI_SYMBOLS = [
    "rxm_tables_format", "rxm_tables_parse", "rxm_tables_release", "rxm_tables_validate",
    "rxm_tables_upload", "rxm_tables_upload_opts", "rxm_plan_query", "rxm_free", "rxm_match_batch", "rxm_match_text",
    "rxm_launch_count", "rxm_overflow_count", "rxm_set_concurrency", "rxm_strerror", "rxm_last_cuda_error",
]


class RxmTables(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("kind", C.c_uint32), ("reversed", C.c_uint32),
        ("n_states", C.c_uint32), ("n_edges", C.c_uint32), ("start", C.c_uint32),
        ("finish", C.c_uint32), ("n_cells", C.c_uint32),
        ("edge_begin", C.POINTER(C.c_uint32)), ("edge_kind", C.POINTER(C.c_uint8)),
        ("edge_sym", C.POINTER(C.c_uint8)), ("edge_to", C.POINTER(C.c_uint16)),
        ("edge_open", C.POINTER(C.c_uint16)), ("edge_close", C.POINTER(C.c_uint16)),
    ]


class RxmUploadOpts(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("engine", C.c_uint32), ("flags", C.c_uint32), ("k3_tile", C.c_uint32),
                ("reserved", C.c_uint32 * 4)]


OPT_K1_NO_QUAD, OPT_K1B_WALK, OPT_INDEX_ORDER, OPT_K1_NO_OCT = 1, 2, 4, 8
ENGINE_IDS = {"k1": 1, "bitset": 2, "k2": 3, "k3": 4, "k4": 5}


class RxmPlanInfo(C.Structure):
    _fields_ = [
        ("engine", C.c_uint32), ("dfa_states", C.c_uint32), ("dfa_classes", C.c_uint32),
        ("exact_step_differs", C.c_uint32), ("n_states", C.c_uint32), ("n_edges", C.c_uint32),
        ("n_cells", C.c_uint32), ("reversed", C.c_uint32), ("sm_count", C.c_uint32),
        ("dfa_stride", C.c_uint32), ("reserved", C.c_uint32 * 6),
    ]


class RxmError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what}: status {status} ({strerror(status)})")


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The loaded librxm.so; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `make -C re2-modification_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.rxm_tables_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(RxmTables))]
        L.rxm_tables_parse.restype = C.c_int
        L.rxm_tables_format.argtypes = [C.POINTER(RxmTables), C.c_char_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]
        L.rxm_tables_format.restype = C.c_int
        L.rxm_tables_release.argtypes = [C.POINTER(RxmTables)]
        L.rxm_tables_release.restype = None
        L.rxm_tables_validate.argtypes = [C.POINTER(RxmTables)]
        L.rxm_tables_validate.restype = C.c_int
        L.rxm_tables_upload.argtypes = [C.POINTER(RxmTables), C.c_int, C.POINTER(C.c_void_p)]
        L.rxm_tables_upload.restype = C.c_int
        L.rxm_tables_upload_opts.argtypes = [C.POINTER(RxmTables), C.c_int, C.POINTER(RxmUploadOpts), C.POINTER(C.c_void_p)]
        L.rxm_tables_upload_opts.restype = C.c_int
        L.rxm_plan_query.argtypes = [C.c_void_p, C.POINTER(RxmPlanInfo)]
        L.rxm_plan_query.restype = C.c_int
        L.rxm_free.argtypes = [C.c_void_p]
        L.rxm_free.restype = C.c_int
        L.rxm_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                      C.c_void_p]
        L.rxm_match_batch.restype = C.c_int
        L.rxm_match_text.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.c_void_p]
        L.rxm_match_text.restype = C.c_int
        L.rxm_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rxm_launch_count.restype = C.c_int
        L.rxm_set_concurrency.argtypes = [C.c_void_p, C.c_uint32]
        L.rxm_set_concurrency.restype = C.c_int
        L.rxm_overflow_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.rxm_overflow_count.restype = C.c_int
        L.rxm_strerror.argtypes = [C.c_int]
        L.rxm_strerror.restype = C.c_char_p
        L.rxm_last_cuda_error.argtypes = []
        L.rxm_last_cuda_error.restype = C.c_char_p
        _lib = L
    return _lib


def strerror(status: int) -> str:
    return lib().rxm_strerror(status).decode()


class Tables:
    """Host tables parsed from the text form (rxm_tables_parse)."""

    def __init__(self, text: str):
        self.text = text
        self._p = C.POINTER(RxmTables)()
        raw = text.encode()
        st = lib().rxm_tables_parse(raw, len(raw), C.byref(self._p))
        if st != RXM_OK:
            raise RxmError(st, "rxm_tables_parse")

    @classmethod
    def load(cls, path: str) -> "Tables":
        with open(path) as f:
            return cls(f.read())

    @property
    def ptr(self):
        return self._p

    @property
    def c(self) -> RxmTables:
        return self._p.contents

    def format(self) -> str:
        need = C.c_size_t(0)
        lib().rxm_tables_format(self._p, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        st = lib().rxm_tables_format(self._p, buf, need.value, C.byref(need))
        if st != RXM_OK:
            raise RxmError(st, "rxm_tables_format")
        return buf.value.decode()

    def __del__(self):
        try:
            if self._p:
                lib().rxm_tables_release(self._p)
                self._p = None
        except Exception:
            pass


class Matcher:
    """Device matcher for one automaton: rxm_tables_upload / rxm_match_batch / rxm_free."""

    def __init__(self, tables: Tables, device: int = 0, engine=None, flags: int = 0, k3_tile: int = 0):
        """engine: None (the planner's choice) or "k1" / "bitset" / "k2" / "k3" / "k4" (rxm_tables_upload_opts);
        flags: OPT_*; k3_tile: 0 or 8 / 16 / 32."""
        self.tables = tables
        self.device = device
        self._h = C.c_void_p()
        self.saw_exit = C.c_int(0)  # set by the last match_text_* call
        if engine is None and not flags and not k3_tile:
            st = lib().rxm_tables_upload(tables.ptr, device, C.byref(self._h))
        else:
            o = RxmUploadOpts(abi_version=1, engine=ENGINE_IDS[engine] if engine else 0, flags=flags, k3_tile=k3_tile)
            st = lib().rxm_tables_upload_opts(tables.ptr, device, C.byref(o), C.byref(self._h))
        if st != RXM_OK:
            raise RxmError(st, "rxm_tables_upload: " + lib().rxm_last_cuda_error().decode())

    def plan(self) -> RxmPlanInfo:
        info = RxmPlanInfo()
        st = lib().rxm_plan_query(self._h, C.byref(info))
        if st != RXM_OK:
            raise RxmError(st, "rxm_plan_query")
        return info

    def match_host(self, chars: np.ndarray, offsets: np.ndarray) -> np.ndarray:
        """Host buffers in, host bits out (H2D + kernel + D2H inside the call)."""
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.empty(n, dtype=np.uint8)
        st = lib().rxm_match_batch(self._h, chars.ctypes.data if chars.size else None,
                                   offsets.ctypes.data, n, out.ctypes.data, None)
        if st != RXM_OK:
            raise RxmError(st, "rxm_match_batch: " + lib().rxm_last_cuda_error().decode())
        return out

    def match_ptrs(self, chars_ptr: int, offsets_ptr: int, n: int, out_ptr: int,
                   stream: int = 0) -> None:
        """Raw pointers (device or host), e.g. torch tensors' data_ptr()."""
        st = lib().rxm_match_batch(self._h, chars_ptr, offsets_ptr, n, out_ptr, stream)
        if st != RXM_OK:
            raise RxmError(st, "rxm_match_batch: " + lib().rxm_last_cuda_error().decode())

    def match_text_host(self, text: bytes, cap: Optional[int] = None) -> np.ndarray:
        """Raw whitespace-delimited text in (host), bits of the tokens before `exit` out:
        rxm_match_text (tokenised on the device)."""
        buf = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(0, dtype=np.uint8)
        cap = (len(text) + 1) // 2 if cap is None else cap
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_uint64(0)
        st = lib().rxm_match_text(self._h, buf.ctypes.data if buf.size else None, buf.size,
                                  out.ctypes.data, cap, C.byref(n), C.byref(self.saw_exit), None)
        if st != RXM_OK:
            e = RxmError(st, "rxm_match_text: " + lib().rxm_last_cuda_error().decode())
            e.n_tokens = n.value
            raise e
        return out[:n.value].copy()

    def match_text_ptrs(self, text_ptr: int, nbytes: int, out_ptr: int, cap: int, stream: int = 0) -> int:
        """Raw pointers (both device or both host); returns the number of tokens matched."""
        n = C.c_uint64(0)
        st = lib().rxm_match_text(self._h, text_ptr, nbytes, out_ptr, cap, C.byref(n),
                                  C.byref(self.saw_exit), stream)
        if st != RXM_OK:
            e = RxmError(st, "rxm_match_text: " + lib().rxm_last_cuda_error().decode())
            e.n_tokens = n.value
            raise e
        return n.value

    def set_concurrency(self, handles: int):
        """rxm_set_concurrency: `handles` matchers run at once on this device (streams of their own)."""
        st = lib().rxm_set_concurrency(self._h, handles)
        if st != RXM_OK:
            raise RxmError(st, "rxm_set_concurrency")

    def launch_count(self) -> int:
        v = C.c_uint64(0)
        lib().rxm_launch_count(self._h, C.byref(v))
        return v.value

    def overflow_count(self) -> int:
        v = C.c_uint64(0)
        lib().rxm_overflow_count(self._h, C.byref(v))
        return v.value

    def close(self):
        if self._h:
            lib().rxm_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
