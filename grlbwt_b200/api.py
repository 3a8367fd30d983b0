"""ctypes bindings of include/grlgpu.h (device parse phase) and include/grlbwt.h (host side).

There is no fallback: if the shared libraries are missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
FLAG_SMALL_TABLE, FLAG_FORCE_SLOW_SCAN, FLAG_KEEP_DICT, FLAG_FORCE_UNCACHED, FLAG_SMALL_PILOT, FLAG_FORCE_DOUBLING = 1, 2, 4, 8, 16, 32
CELL = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


class GrlGpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"grlgpu status {status}: {msg}")
        self.status = status


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_syms", "n_strings", "longest_string", "min_sym", "max_sym", "max_sym_freq", "sep_sym")]


class Round(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("round", "n_in", "n_strings", "parse_len", "n_phrases", "dict_syms", "max_freq", "alphabet",
                                          "tot_phrases", "n_pre_runs", "algorithmic_bytes")] + \
               [(k, C.c_uint32) for k in ("cell_bytes_in", "cell_bytes_out", "sym_bytes", "done")] + \
               [(k, C.c_float) for k in ("device_ms", "text_pass_ms", "dict_ms", "rewrite_ms")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Slice(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("rank_base", "tot_local", "pre_first", "n_pre_local", "exchange_bytes", "n_in_local", "parse_len_local")] + \
               [("sym_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class BwtResult(C.Structure):
    _fields_ = [("n_runs", C.c_uint64), ("sb", C.c_uint64), ("fb", C.c_uint64), ("syms", C.POINTER(C.c_uint64)),
                ("lens", C.POINTER(C.c_uint64)), ("n_rounds", C.c_uint64), ("h2d_ms", C.c_double), ("par_phase_ms", C.c_double),
                ("ind_phase_ms", C.c_double), ("device_ms", C.c_double), ("algorithmic_bytes", C.c_uint64), ("induced_on_device", C.c_uint64)]


_gpu = None
_host = None


def lib_gpu():
    global _gpu
    if _gpu is None:
        path = os.path.join(LIB_DIR, "libgrlgpu.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C grlbwt_b200` (or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(path)
        vp, u64 = C.c_void_p, C.c_uint64
        L.grlgpu_create.argtypes = [C.POINTER(vp), C.c_int, u64]
        L.grlgpu_destroy.argtypes = [vp]
        L.grlgpu_set_text.argtypes = [vp, vp, u64, C.c_int]
        L.grlgpu_set_text_device.argtypes = [vp, vp, u64, C.c_int]
        L.grlgpu_stats.argtypes = [vp, C.POINTER(Stats)]
        L.grlgpu_round.argtypes = [vp, C.POINTER(Round)]
        L.grlgpu_fetch_level.argtypes = [vp, vp, vp, vp, vp, vp]
        L.grlgpu_fetch_level_async.argtypes = [vp, vp, vp, vp, vp, vp]
        L.grlgpu_fetch_level32.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
        L.grlgpu_fetch_wait.argtypes = [vp]
        L.grlgpu_level_checksum.argtypes = [vp, vp]
        L.grlgpu_fetch_parse.argtypes = [vp, vp]
        L.grlgpu_fetch_str_ptrs.argtypes = [vp, vp]
        L.grlgpu_fetch_dictionary.argtypes = [vp, vp, vp, vp, vp]
        L.grlgpu_strerror.restype = C.c_char_p
        L.grlgpu_strerror.argtypes = [C.c_int]
        L.grlgpu_last_error.restype = C.c_char_p
        L.grlgpu_last_error.argtypes = [vp]
        L.grlgpu_create_on_stream.argtypes = [C.POINTER(vp), C.c_int, u64, vp]
        L.grlgpu_profile_enable.argtypes = [vp, C.c_int]
        L.grlgpu_profile_reset.argtypes = [vp]
        L.grlgpu_launch_count.argtypes = [vp]
        L.grlgpu_launch_count.restype = u64
        L.grlgpu_profile_entry.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(u64), C.POINTER(C.c_double), C.POINTER(u64)]
        L.grlgpu_histogram.argtypes = [vp, vp]
        L.grlgpu_nccl_unique_id.argtypes = [vp]
        L.grlgpu_comm_create_nccl.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int, C.c_int]
        L.grlgpu_comm_create_ipc.argtypes = [C.POINTER(vp), C.c_char_p, C.c_int, C.c_int, C.c_int]
        L.grlgpu_can_peer.argtypes = [C.c_int, C.c_int]
        L.grlgpu_local_group_create.argtypes = [C.POINTER(vp), C.c_int]
        L.grlgpu_local_group_abort.argtypes = [vp]
        L.grlgpu_local_group_destroy.argtypes = [vp]
        L.grlgpu_comm_create_local.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int]
        L.grlgpu_comm_destroy.argtypes = [vp]
        L.grlgpu_comm_info.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.c_char_p, C.c_int]
        L.grlgpu_comm_times.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.grlgpu_set_peers.argtypes = [vp, vp, C.c_int]
        L.grlgpu_mg_stats.argtypes = [vp, vp, C.POINTER(Stats)]
        L.grlgpu_mg_round.argtypes = [vp, vp, C.POINTER(Round)]
        L.grlgpu_mg_slice_info.argtypes = [vp, C.POINTER(Slice)]
        L.grlgpu_mg_fetch_slice.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        L.grlgpu_mg_slice_checksum.argtypes = [vp, vp]
        L.grlgpu_text_begin.argtypes = [vp, u64, C.c_int, u64]
        L.grlgpu_text_stage.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
        L.grlgpu_text_commit.argtypes = [vp, u64]
        L.grlgpu_text_end.argtypes = [vp]
        L.grlgpu_level_park.argtypes = [vp, C.c_int, vp]
        L.grlgpu_copy_to_host.argtypes = [C.c_int, vp, vp, u64]
        L.grlgpu_selftest_scan.argtypes = [vp, u64, vp, vp]
        L.grlgpu_selftest_sort.argtypes = [vp, vp, u64, C.c_int]
        L.grlgpu_selftest_compact.argtypes = [vp, vp, u64, vp, vp]
        L.grlgpu_selftest_ipc_rendezvous.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        _gpu = L
    return _gpu


def lib_host():
    global _host
    if _host is None:
        lib_gpu()
        path = os.path.join(LIB_DIR, "libgrlbwt.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make -C grlbwt_b200`")
        L = C.CDLL(path)
        L.grlbwt_build.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(BwtResult)]
        L.grlbwt_free_result.argtypes = [C.POINTER(BwtResult)]
        L.grlbwt_build_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.grlbwt_last_error.restype = C.c_char_p
        L.grlbwt_build_mg.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(BwtResult)]
        L.grlbwt_build_to.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(BwtResult)]
        L.grlbwt_build_packed.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                                          C.POINTER(BwtResult)]
        L.grlbwt_last_digests.argtypes = [C.c_void_p, C.c_uint64]
        L.grlbwt_last_digests.restype = C.c_uint64
        L.grlbwt_last_exchange_bytes.restype = C.c_uint64
        L.grlbwt_last_comm.restype = C.c_char_p
        _host = L
    return _host


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class GrlGpu:
    """One device context = the GPU parse strategy (mirrors the duck-typed strategy concept used by
    par_round, exact_par_phase.cpp:374-497: get_phrases / map / parse_text, fused into round())."""

    def __init__(self, device: int = 0, flags: int = 0, stream: int = 0):
        self._L = lib_gpu()
        self._h = C.c_void_p()
        self._check(self._L.grlgpu_create_on_stream(C.byref(self._h), device, flags, C.c_void_p(stream) if stream else None), ctx=False)
        self.last = None
        self._keep = None
        self._sym_bytes = 1
        self._round_started = False

    def _check(self, rc, ctx=True):
        if rc != 0:
            msg = self._L.grlgpu_strerror(rc).decode()
            if ctx and self._h:
                msg += " | " + self._L.grlgpu_last_error(self._h).decode()
            raise GrlGpuError(rc, msg)

    def set_text(self, text: np.ndarray):
        text = np.ascontiguousarray(text)
        if text.dtype not in (np.uint8, np.uint16, np.uint32, np.uint64):
            raise GrlGpuError(-1, "symbol width must be 1, 2, 4 or 8 bytes")
        self._keep = text
        self._sym_bytes, self._round_started, self.last = text.dtype.itemsize, False, None
        self._check(self._L.grlgpu_set_text(self._h, _ptr(text) if text.size else None, text.size, text.dtype.itemsize))

    def set_text_device(self, dev_ptr: int, n_syms: int, sym_bytes: int):
        self._keep, self._sym_bytes, self._round_started, self.last = None, sym_bytes, False, None
        self._check(self._L.grlgpu_set_text_device(self._h, C.c_void_p(dev_ptr), n_syms, sym_bytes))

    def stats(self) -> Stats:
        s = Stats()
        self._check(self._L.grlgpu_stats(self._h, C.byref(s)))
        return s

    def round(self) -> Round:
        r = Round()
        self._check(self._L.grlgpu_round(self._h, C.byref(r)))
        self.last = r
        self._round_started = True
        return r

    def level_checksum(self):
        """four order-insensitive sums over the last round's rules / hocc marks / preliminary BWT (see grlgpu.h)"""
        out = np.zeros(4, np.uint64)
        self._check(self._L.grlgpu_level_checksum(self._h, _ptr(out)))
        return [int(x) for x in out]

    def fetch_wait(self):
        self._check(self._L.grlgpu_fetch_wait(self._h))

    def fetch_level(self, arena: np.ndarray | None = None, widen: bool = True, async_: bool = False, offset: int = 0, narrow_len: bool = False):
        """-> dict(rule_l, rule_r, has_hocc, pre_sym, pre_len) as numpy arrays.
        arena: optional uint8 buffer (e.g. pinned host memory) the arrays are carved from, so the copies are
        direct DMA; widen=False keeps rule/pre_sym in the device's element width (sym_bytes) instead of u64;
        narrow_len: 32-bit run lengths (grlgpu_fetch_level32) whenever the round's n_in + parse_len < 2^32."""
        r = self.last
        st = np.uint32 if r.sym_bytes == 4 else np.uint64
        narrow = bool(narrow_len) and r.n_in + r.parse_len < (1 << 32)
        sizes = [(r.tot_phrases, st), (r.tot_phrases, st), (r.tot_phrases, np.uint8), (r.n_pre_runs, st), (r.n_pre_runs, np.uint32 if narrow else np.uint64)]
        if arena is None:
            arrs = [np.zeros(n, dt) for n, dt in sizes]
        else:
            arrs, off = [], int(offset)
            for n, dt in sizes:
                nb = n * np.dtype(dt).itemsize
                off = (off + 15) & ~15
                if off + nb > arena.size:
                    raise GrlGpuError(-1, "fetch arena too small")
                arrs.append(arena[off:off + nb].view(dt))
                off += nb
        rl, rr, hh, ps, pl = arrs
        if narrow:
            self._check(self._L.grlgpu_fetch_level32(self._h, _ptr(rl), _ptr(rr), _ptr(hh), _ptr(ps), _ptr(pl), int(async_)))
        else:
            fn = self._L.grlgpu_fetch_level_async if async_ else self._L.grlgpu_fetch_level
            self._check(fn(self._h, _ptr(rl), _ptr(rr), _ptr(hh), _ptr(ps), _ptr(pl)))
        if arena is not None:
            self.arena_end = (off + 15) & ~15
        if widen and arena is None:
            rl, rr, ps = rl.astype(np.uint64), rr.astype(np.uint64), ps.astype(np.uint64)
        return {"rule_l": rl, "rule_r": rr, "has_hocc": hh, "pre_sym": ps, "pre_len": pl}

    def fetch_parse(self, arena: np.ndarray | None = None) -> np.ndarray:
        r = self.last
        if arena is None:
            out = np.zeros(r.parse_len, CELL[r.cell_bytes_out])
        else:
            out = arena[: r.parse_len * r.cell_bytes_out].view(CELL[r.cell_bytes_out])
        self._check(self._L.grlgpu_fetch_parse(self._h, _ptr(out)))
        return out

    def fetch_str_ptrs(self) -> np.ndarray:
        out = np.zeros(self.last.n_strings + 1, np.uint64)
        self._check(self._L.grlgpu_fetch_str_ptrs(self._h, _ptr(out)))
        return out

    def fetch_dictionary(self):
        r = self.last
        syms, lens = np.zeros(r.dict_syms, np.uint64), np.zeros(r.n_phrases, np.uint64)
        freqs, metas = np.zeros(r.n_phrases, np.uint64), np.zeros(r.n_phrases, np.uint64)
        self._check(self._L.grlgpu_fetch_dictionary(self._h, _ptr(syms), _ptr(lens), _ptr(freqs), _ptr(metas)))
        return syms, lens, freqs, metas

    # ---- multi-GPU rounds: every rank makes the same calls in the same order (include/grlgpu.h) ----
    def histogram(self) -> np.ndarray:
        h = np.zeros(256, np.uint64)
        self._check(self._L.grlgpu_histogram(self._h, _ptr(h)))
        return h

    def set_peers(self, devices):
        d = np.asarray(devices, np.int32)
        self._check(self._L.grlgpu_set_peers(self._h, _ptr(d), d.size))

    def mg_stats(self, comm) -> Stats:
        s = Stats()
        self._check(self._L.grlgpu_mg_stats(self._h, comm.handle, C.byref(s)))
        return s

    def mg_round(self, comm) -> Round:
        r = Round()
        self._check(self._L.grlgpu_mg_round(self._h, comm.handle, C.byref(r)))
        self.last = r
        self._round_started = True
        return r

    def mg_slice_info(self) -> Slice:
        s = Slice()
        self._check(self._L.grlgpu_mg_slice_info(self._h, C.byref(s)))
        return s

    def mg_slice_checksum(self):
        out = np.zeros(4, np.uint64)
        self._check(self._L.grlgpu_mg_slice_checksum(self._h, _ptr(out)))
        return [int(x) for x in out]

    def mg_fetch_slice(self, arena: np.ndarray | None = None, offset: int = 0, narrow_len: bool = False, async_: bool = False):
        """this rank's part of the last level -> dict of numpy arrays (carved from `arena`, e.g. pinned memory, when given)"""
        sl = self.mg_slice_info()
        st = np.uint32 if sl.sym_bytes == 4 else np.uint64
        sizes = [(sl.tot_local, st), (sl.tot_local, st), (sl.tot_local, np.uint8), (sl.n_pre_local, st), (sl.n_pre_local, np.uint32 if narrow_len else np.uint64)]
        if arena is None:
            arrs = [np.zeros(n, dt) for n, dt in sizes]
        else:
            arrs, off = [], int(offset)
            for n, dt in sizes:
                nb = n * np.dtype(dt).itemsize
                off = (off + 15) & ~15
                if off + nb > arena.size:
                    raise GrlGpuError(-1, "fetch arena too small")
                arrs.append(arena[off:off + nb].view(dt))
                off += nb
            self.arena_end = (off + 15) & ~15
        rl, rr, hh, ps, pl = arrs
        self._check(self._L.grlgpu_mg_fetch_slice(self._h, _ptr(rl), _ptr(rr), _ptr(hh), _ptr(ps), _ptr(pl), 4 if narrow_len else 8, int(async_)))
        return {"rule_l": rl, "rule_r": rr, "has_hocc": hh, "pre_sym": ps, "pre_len": pl, "rank_base": sl.rank_base, "pre_first": sl.pre_first}

    def fetch_parse_local(self, n_cells: int, arena: np.ndarray | None = None) -> np.ndarray:
        """multi-GPU: this rank's part of the current parse (parse_len_local cells of grlgpu_slice_t)"""
        r = self.last
        out = np.zeros(n_cells, CELL[r.cell_bytes_out]) if arena is None else arena[: n_cells * r.cell_bytes_out].view(CELL[r.cell_bytes_out])
        self._check(self._L.grlgpu_fetch_parse(self._h, _ptr(out)))
        return out

    def profile_enable(self, on: bool = True):
        self._check(self._L.grlgpu_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._check(self._L.grlgpu_profile_reset(self._h))

    def launch_count(self) -> int:
        return int(self._L.grlgpu_launch_count(self._h))

    def profile(self):
        """-> {kernel name: (launches, total_ms, model_bytes)} measured with CUDA events on the launch stream"""
        out, i = {}, 0
        name = C.create_string_buffer(128)
        n, ms, by = C.c_uint64(), C.c_double(), C.c_uint64()
        while self._L.grlgpu_profile_entry(self._h, i, name, 128, C.byref(n), C.byref(ms), C.byref(by)) == 0:
            out[name.value.decode()] = (int(n.value), float(ms.value), int(by.value))
            i += 1
        return out

    def close(self):
        if self._h:
            self._L.grlgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_bwt(text: np.ndarray, device: int = 0, n_threads: int = 1, verbose: bool = False):
    """Whole construction (device parse phase + host induction) -> (syms u64, lens u64, sb, fb, info dict)."""
    L = lib_host()
    text = np.ascontiguousarray(text)
    res = BwtResult()
    rc = L.grlbwt_build(_ptr(text), text.size, text.dtype.itemsize, device, n_threads, int(verbose), C.byref(res))
    if rc != 0:
        raise GrlGpuError(rc, L.grlbwt_last_error().decode())
    try:
        syms = np.ctypeslib.as_array(res.syms, shape=(res.n_runs,)).copy()
        lens = np.ctypeslib.as_array(res.lens, shape=(res.n_runs,)).copy()
        info = {k: getattr(res, k) for k in ("n_rounds", "h2d_ms", "par_phase_ms", "ind_phase_ms", "device_ms", "algorithmic_bytes", "induced_on_device")}
        return syms, lens, int(res.sb), int(res.fb), info
    finally:
        L.grlbwt_free_result(C.byref(res))


def build_bwt_to(text: np.ndarray, out_syms: np.ndarray, out_lens: np.ndarray, devices=(0,), n_threads: int = 1, comm: int = 0):
    """Whole construction with the run-length BWT delivered into caller-owned uint32 arrays (e.g. views of pinned memory that is
    reused from call to call). -> (n_runs, sb, fb, info)"""
    L = lib_host()
    text = np.ascontiguousarray(text)
    assert out_syms.dtype == np.uint32 and out_lens.dtype == np.uint32 and out_syms.flags.c_contiguous and out_lens.flags.c_contiguous
    dv = np.asarray(devices, np.int32)
    res = BwtResult()
    rc = L.grlbwt_build_to(_ptr(text), text.size, text.dtype.itemsize, _ptr(dv), dv.size, n_threads, comm, _ptr(out_syms), _ptr(out_lens),
                           min(out_syms.size, out_lens.size), C.byref(res))
    if rc != 0:
        raise GrlGpuError(rc, L.grlbwt_last_error().decode())
    info = {k: getattr(res, k) for k in ("n_rounds", "h2d_ms", "par_phase_ms", "ind_phase_ms", "device_ms", "algorithmic_bytes", "induced_on_device")}
    return int(res.n_runs), int(res.sb), int(res.fb), info


def build_bwt_packed(text: np.ndarray, out_image: np.ndarray, devices=(0,), n_threads: int = 1, comm: int = 0):
    """Whole construction with the run-length BWT delivered as the image of the .rl_bwt file ([sb u64][fb u64] + records of sb + fb
    bytes) in a caller-owned uint8 array; the records are packed on the device. -> (image_bytes, n_runs, sb, fb, info)"""
    L = lib_host()
    text = np.ascontiguousarray(text)
    assert out_image.dtype == np.uint8 and out_image.flags.c_contiguous
    dv = np.asarray(devices, np.int32)
    res = BwtResult()
    nb = C.c_uint64()
    rc = L.grlbwt_build_packed(_ptr(text), text.size, text.dtype.itemsize, _ptr(dv), dv.size, n_threads, comm, _ptr(out_image), out_image.size, C.byref(nb), C.byref(res))
    if rc != 0:
        raise GrlGpuError(rc, L.grlbwt_last_error().decode())
    info = {k: getattr(res, k) for k in ("n_rounds", "h2d_ms", "par_phase_ms", "ind_phase_ms", "device_ms", "algorithmic_bytes", "induced_on_device")}
    return int(nb.value), int(res.n_runs), int(res.sb), int(res.fb), info


def parse_rl_bwt(image) -> tuple:
    """The .rl_bwt format (reference include/bwt_io.h:377-382,448-490) as numpy arrays: `image` is a file path or the bytes /
    uint8 array of a file image (what build_bwt_packed fills). -> (syms uint64[r], lens uint64[r], sb, fb). Host-side helper."""
    raw = np.fromfile(image, np.uint8) if isinstance(image, (str, os.PathLike)) else np.frombuffer(image, np.uint8)
    if raw.size < 16:
        raise ValueError("truncated header")
    sb, fb = (int(x) for x in raw[:16].view(np.uint64))
    if not (1 <= sb <= 8 and 1 <= fb <= 8):
        raise ValueError(f"bad header widths sb={sb} fb={fb}")
    body = raw[16:]
    if body.size % (sb + fb):
        raise ValueError("truncated record at the end of the image")
    rec = body.reshape(-1, sb + fb)
    syms = np.zeros(rec.shape[0], np.uint64)
    lens = np.zeros(rec.shape[0], np.uint64)
    for i in range(sb):
        syms |= rec[:, i].astype(np.uint64) << np.uint64(8 * i)
    for i in range(fb):
        lens |= rec[:, sb + i].astype(np.uint64) << np.uint64(8 * i)
    return syms, lens, sb, fb


def build_bwt_file(inp: str, out: str, sym_bytes: int = 1, device: int = 0, n_threads: int = 1, verbose: bool = False):
    L = lib_host()
    rc = L.grlbwt_build_file(inp.encode(), out.encode(), sym_bytes, device, n_threads, int(verbose))
    if rc != 0:
        raise GrlGpuError(rc, L.grlbwt_last_error().decode())


def selftest_induce(levels, final_parse: np.ndarray, n_threads: int = 0):
    """levels: list of dicts with alphabet, tot, rule_l, rule_r (u64), has_hocc (u8), pre_sym, pre_len (u64). CPU only."""
    L = lib_host()
    n = len(levels)
    u64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8)
    keep = []

    def arr(key, dt, ptype):
        out = (ptype * n)()
        for i, lv in enumerate(levels):
            a = np.ascontiguousarray(lv[key], dt)
            if a.size == 0:
                a = np.zeros(1, dt)
            keep.append(a)
            out[i] = a.ctypes.data_as(ptype)
        return out

    alph = np.array([lv["alphabet"] for lv in levels], np.uint64)
    tot = np.array([lv["tot"] for lv in levels], np.uint64)
    npre = np.array([len(lv["pre_sym"]) for lv in levels], np.uint64)
    fp = np.ascontiguousarray(final_parse, np.uint64)
    res = BwtResult()
    L.grlbwt_selftest_induce.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(u64p), C.POINTER(u64p), C.POINTER(u8p), C.c_void_p,
                                         C.POINTER(u64p), C.POINTER(u64p), C.c_void_p, C.c_uint64, C.c_int, C.POINTER(BwtResult)]
    rc = L.grlbwt_selftest_induce(n, _ptr(alph), _ptr(tot), arr("rule_l", np.uint64, u64p), arr("rule_r", np.uint64, u64p),
                                  arr("has_hocc", np.uint8, u8p), _ptr(npre), arr("pre_sym", np.uint64, u64p), arr("pre_len", np.uint64, u64p),
                                  _ptr(fp), fp.size, n_threads, C.byref(res))
    if rc != 0:
        raise RuntimeError(L.grlbwt_last_error().decode())
    try:
        return (np.ctypeslib.as_array(res.syms, shape=(res.n_runs,)).copy(), np.ctypeslib.as_array(res.lens, shape=(res.n_runs,)).copy())
    finally:
        L.grlbwt_free_result(C.byref(res))


def selftest_scan(a: np.ndarray):
    a = np.ascontiguousarray(a, np.uint32)
    out, tot = np.zeros(a.size, np.uint64), np.zeros(1, np.uint64)
    rc = lib_gpu().grlgpu_selftest_scan(_ptr(a), a.size, _ptr(out), _ptr(tot))
    if rc != 0:
        raise GrlGpuError(rc, lib_gpu().grlgpu_strerror(rc).decode())
    return out, int(tot[0])


def selftest_sort(keys: np.ndarray, vals: np.ndarray, n_bits: int):
    k, v = np.ascontiguousarray(keys, np.uint64).copy(), np.ascontiguousarray(vals, np.uint32).copy()
    rc = lib_gpu().grlgpu_selftest_sort(_ptr(k), _ptr(v), k.size, n_bits)
    if rc != 0:
        raise GrlGpuError(rc, lib_gpu().grlgpu_strerror(rc).decode())
    return k, v


def selftest_compact(bits: np.ndarray, prev_bits, n_bits: int):
    b = np.ascontiguousarray(bits, np.uint32)
    pb = None if prev_bits is None else np.ascontiguousarray(prev_bits, np.uint32)
    out, cnt = np.zeros(max(1, n_bits), np.uint64), np.zeros(1, np.uint64)
    rc = lib_gpu().grlgpu_selftest_compact(_ptr(b), _ptr(pb), n_bits, _ptr(out), _ptr(cnt))
    if rc != 0:
        raise GrlGpuError(rc, lib_gpu().grlgpu_strerror(rc).decode())
    return out[: int(cnt[0])]
