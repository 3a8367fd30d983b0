"""TEST INFRASTRUCTURE ONLY: ctypes front-end of oracle/oracle.c (the CPU restatement of the
reference's parse + induction phases) plus an independent numpy definition of the BCR BWT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  Nothing under grlbwt_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

# level scalars / arrays (must match the enums in oracle.c)
N_IN, P, D, SUM_LEN, MAX_FREQ, TOT_PHRASES, ALPHABET, CELL_BYTES, N_PRE, PARSE_LEN, N_STR, LONGEST = range(12)
A_PARSE, A_STR_PTRS, A_DICT_SYMS, A_DICT_LEN, A_DICT_FREQ, A_DICT_META, A_PRE_SYM, A_PRE_LEN, A_RULE_L, A_RULE_R, \
    A_HAS_HOCC, A_IS_SUFFIX = range(12)
S_N_SYMS, S_N_STRINGS, S_LONGEST, S_MIN, S_MAX, S_MAX_SYM_FREQ, S_SEP = range(7)

_lib = None


def build() -> str:
    """Compile oracle.c -> _build/liboracle.so (gcc only; no reference sources needed)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_int)]
        L.oracle_par_phase.restype = C.c_int
        L.oracle_par_phase.argtypes = [C.c_void_p]
        L.oracle_level_scalar.restype = C.c_uint64
        L.oracle_level_scalar.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_level_array.restype = C.POINTER(C.c_uint64)
        L.oracle_level_array.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.oracle_stat.restype = C.c_uint64
        L.oracle_stat.argtypes = [C.c_void_p, C.c_int]
        L.oracle_ind_phase.restype = C.c_int
        L.oracle_ind_phase.argtypes = [C.c_void_p]
        L.oracle_bwt_runs.restype = C.c_uint64
        L.oracle_bwt_runs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_write_rl_bwt.restype = C.c_uint64
        L.oracle_write_rl_bwt.argtypes = [C.c_void_p, C.c_char_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class Oracle:
    """One collection through the restated par_phase (+ optionally ind_phase)."""

    def __init__(self, text: np.ndarray):
        text = np.ascontiguousarray(text)
        assert text.dtype in (np.uint8, np.uint16, np.uint32, np.uint64)
        err = C.c_int(0)
        self._h = lib().oracle_create(text.ctypes.data, text.size, text.dtype.itemsize, C.byref(err))
        if not self._h:
            raise ValueError({-1: "empty input or bad symbol width", -2: "ill formed collection"}.get(err.value, "error"))
        self.n_rounds = 0

    def par_phase(self) -> int:
        self.n_rounds = lib().oracle_par_phase(self._h)
        return self.n_rounds

    def stat(self, what: int) -> int:
        return int(lib().oracle_stat(self._h, what))

    def scalar(self, level: int, what: int) -> int:
        return int(lib().oracle_level_scalar(self._h, level, what))

    def array(self, level: int, what: int) -> np.ndarray:
        cnt = C.c_uint64(0)
        p = lib().oracle_level_array(self._h, level, what, C.byref(cnt))
        if cnt.value == 0:
            return np.zeros(0, np.uint64)
        return np.ctypeslib.as_array(p, shape=(cnt.value,)).copy()

    def ind_phase(self):
        """-> (syms u64[r], lens u64[r], sb, fb)"""
        rc = lib().oracle_ind_phase(self._h)
        assert rc == 0
        ps, pl = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
        sb, fb = C.c_uint64(0), C.c_uint64(0)
        n = lib().oracle_bwt_runs(self._h, C.byref(ps), C.byref(pl), C.byref(sb), C.byref(fb))
        syms = np.ctypeslib.as_array(ps, shape=(n,)).copy()
        lens = np.ctypeslib.as_array(pl, shape=(n,)).copy()
        return syms, lens, int(sb.value), int(fb.value)

    def write_rl_bwt(self, path: str) -> int:
        return int(lib().oracle_write_rl_bwt(self._h, path.encode()))

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# Independent definition (no grammar, no induction): BCR BWT by numpy prefix doubling.
# Terminators are ordered by string index; the symbol preceding a string's first symbol is the
# separator that ends the same string.
# ---------------------------------------------------------------------------------------------
def bcr_bwt(text: np.ndarray) -> np.ndarray:
    T = np.asarray(text).astype(np.int64)
    n = T.size
    sep = T[-1]
    is_sep = T == sep
    n_str = int(is_sep.sum())
    sid = np.cumsum(is_sep) - is_sep  # index of the string each cell belongs to
    # each string is a cyclic unit: successor of its terminator is its own first symbol
    starts = np.concatenate(([0], np.flatnonzero(is_sep)[:-1] + 1))
    nxt = np.arange(1, n + 1, dtype=np.int64)
    nxt[np.flatnonzero(is_sep)] = starts
    prv = np.empty(n, np.int64)
    prv[nxt] = np.arange(n, dtype=np.int64)
    key = np.where(is_sep, sid, T + n_str)
    _, rank = np.unique(key, return_inverse=True)
    rank = rank.astype(np.int64)
    hop = nxt.copy()
    k = 1
    while True:
        pair = rank * (n + 1) + rank[hop]
        u, rank = np.unique(pair, return_inverse=True)
        rank = rank.astype(np.int64)
        if u.size == n or k > 2 * n:
            break
        hop = hop[hop]
        k *= 2
    sa = np.argsort(rank, kind="stable")
    return T[prv[sa]]


def rle(a: np.ndarray):
    a = np.asarray(a)
    if a.size == 0:
        return a[:0].astype(np.uint64), np.zeros(0, np.uint64)
    b = np.flatnonzero(np.concatenate(([True], a[1:] != a[:-1])))
    lens = np.diff(np.concatenate((b, [a.size])))
    return a[b].astype(np.uint64), lens.astype(np.uint64)


def read_rl_bwt(path_or_bytes):
    """Parse a .rl_bwt (bwt_io.h:377-382,448-490) -> (syms u64, lens u64, sb, fb)."""
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    sb = int.from_bytes(raw[0:8], "little")
    fb = int.from_bytes(raw[8:16], "little")
    body = np.frombuffer(raw, np.uint8, offset=16).reshape(-1, sb + fb)
    def le(cols):
        out = np.zeros(cols.shape[0], np.uint64)
        for i in range(cols.shape[1]):
            out |= cols[:, i].astype(np.uint64) << np.uint64(8 * i)
        return out
    return le(body[:, :sb]), le(body[:, sb:]), sb, fb


def rl_bwt_bytes(syms, lens, sb, fb) -> bytes:
    syms = np.asarray(syms, np.uint64)
    lens = np.asarray(lens, np.uint64)
    rec = np.zeros((syms.size, sb + fb), np.uint8)
    for i in range(sb):
        rec[:, i] = (syms >> np.uint64(8 * i)) & np.uint64(0xFF)
    for i in range(fb):
        rec[:, sb + i] = (lens >> np.uint64(8 * i)) & np.uint64(0xFF)
    return int(sb).to_bytes(8, "little") + int(fb).to_bytes(8, "little") + rec.tobytes()
