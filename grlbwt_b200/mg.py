"""Multi-GPU plumbing for tests and bench.py: ctypes handles of the exchange layer (include/grlgpu.h, grlgpu_comm) and a
few helpers. The rounds, the exchanges and the per-rank host threads all live in the C/C++ libraries
(grlbwt_b200/csrc/mg2.cuh, comm.hpp, host/gpu_par_phase.hpp); nothing here computes.

Two ways to run N ranks:
  * one process per GPU (bench.py under torchrun): `comm_from_torch` -- on one box whose GPUs reach each other over NVLink the
    ranks exchange through CUDA IPC windows and a shared-memory rendezvous (`ipc_comm`; rank 0 picks the segment's name and
    torch.distributed broadcasts it); otherwise, or on request, NCCL (`nccl_comm_from_torch`: rank 0 makes the NCCL id,
    torch.distributed broadcasts its 128 bytes). Either way the communicator lives inside libgrlgpu.so;
  * one process, N host threads (the grlbwt CLI with --gpus N, `build_bwt_mg` here): NCCL on request (COMM_NCCL), by default
    in-process peer copies over NVLink (several ranks may also share one GPU: the N > 1 path on a 1-GPU box).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .api import BwtResult, GrlGpuError, _ptr, lib_gpu, lib_host

COMM_AUTO, COMM_LOCAL, COMM_NCCL = 0, 1, 2


class Comm:
    """one rank's end of the exchange layer"""

    def __init__(self, handle, group=None):
        self.handle, self._group = handle, group

    def info(self):
        L = lib_gpu()
        b, nb, ns = C.c_uint64(), C.c_uint64(), C.c_uint64()
        kind = C.create_string_buffer(128)
        L.grlgpu_comm_info(self.handle, C.byref(b), C.byref(nb), C.byref(ns), kind, 128)
        mb, ms = C.c_double(), C.c_double()
        L.grlgpu_comm_times(self.handle, C.byref(mb), C.byref(ms))
        return {"bytes_sent": int(b.value), "bulk_collectives": int(nb.value), "small_collectives": int(ns.value), "kind": kind.value.decode(),
                "ms_in_bulk": round(mb.value, 1), "ms_in_small": round(ms.value, 1)}

    def close(self):
        if self.handle:
            lib_gpu().grlgpu_comm_destroy(self.handle)
            self.handle = C.c_void_p()


def nccl_unique_id() -> np.ndarray:
    buf = np.zeros(128, np.uint8)
    rc = lib_gpu().grlgpu_nccl_unique_id(_ptr(buf))
    if rc != 0:
        raise GrlGpuError(rc, "NCCL is not available")
    return buf


def nccl_comm(id128: np.ndarray, rank: int, world: int, device: int) -> Comm:
    h = C.c_void_p()
    id128 = np.ascontiguousarray(id128, np.uint8)
    rc = lib_gpu().grlgpu_comm_create_nccl(C.byref(h), _ptr(id128), rank, world, device)
    if rc != 0:
        raise GrlGpuError(rc, "ncclCommInitRank failed")
    return Comm(h)


def nccl_comm_from_torch(dist, rank: int, world: int, device_index: int, torch) -> Comm:
    """one process per GPU: the id travels through torch.distributed (any backend), the communicator lives in libgrlgpu.so"""
    dev = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        t.copy_(torch.from_numpy(nccl_unique_id()))
    dist.broadcast(t, src=0)
    return nccl_comm(t.cpu().numpy(), rank, world, device_index)


def ipc_comm(session: str, rank: int, world: int, device: int) -> Comm:
    """one rank per process on one box: CUDA IPC windows + POSIX shared-memory rendezvous named `session` ("/name")"""
    h = C.c_void_p()
    L = lib_gpu()
    rc = L.grlgpu_comm_create_ipc(C.byref(h), session.encode(), rank, world, device)
    if rc != 0:
        raise GrlGpuError(rc, "IPC communicator: " + L.grlgpu_last_error(None).decode())
    return Comm(h)


def can_peer_all(devices) -> bool:
    L = lib_gpu()
    return all(L.grlgpu_can_peer(int(a), int(b)) == 1 for a in devices for b in devices)


def comm_from_torch(dist, rank: int, world: int, device_index: int, torch, kind: str = "auto") -> Comm:
    """one process per GPU. kind: "ipc", "nccl" or "auto" (= ipc when every rank runs on this box and every pair of their GPUs
    has peer access, else nccl; GRLBWT_COMM overrides). Every rank takes the same decision from the same facts."""
    kind = os.environ.get("GRLBWT_COMM", kind)
    if kind == "local":
        kind = "ipc"
    if kind not in ("auto", "ipc", "nccl"):
        raise ValueError(f"unknown exchange backend {kind!r}")
    auto = kind == "auto"
    if auto:
        one_box = int(os.environ.get("LOCAL_WORLD_SIZE", "0")) == world
        kind = "ipc" if one_box and can_peer_all(range(world)) else "nccl"
    if kind == "nccl":
        return nccl_comm_from_torch(dist, rank, world, device_index, torch)
    dev = torch.device("cuda", device_index) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.zeros(64, dtype=torch.uint8, device=dev)
    if rank == 0 and _shm_usable():   # (an all-zero name tells every rank that rank 0 cannot create the segment)
        name = f"/grlgpu-{os.getpid()}-{int.from_bytes(os.urandom(6), 'little'):x}".encode()
        t[: len(name)] = torch.frombuffer(bytearray(name), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=0)
    session = bytes(t.cpu().numpy()).rstrip(b"\0").decode()
    comm, err = None, None
    if session:
        try:
            comm = ipc_comm(session, rank, world, device_index)
        except GrlGpuError as e:   # CUDA IPC or the segment is not usable here: the constructor fails on every rank together
            err = e
    else:
        err = GrlGpuError(-5, "POSIX shared memory is not usable on this box")
    ok = torch.tensor([1 if comm is not None else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 1:
        return comm
    if comm is not None:
        comm.close()
    if not auto:
        raise err if err is not None else GrlGpuError(-5, "the IPC backend failed on another rank")
    return nccl_comm_from_torch(dist, rank, world, device_index, torch)


def _shm_usable() -> bool:
    try:
        path = f"/dev/shm/grlgpu-probe-{os.getpid()}"
        fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
        os.close(fd)
        os.unlink(path)
        return True
    except OSError:
        return False


def shard_bounds(text: np.ndarray, n_ranks: int):
    """Contiguous ranges of WHOLE strings balanced by symbol count (same rule as host/gpu_par_phase.hpp: boundary r is the end
    of the first string that ends at or after cell r*n/G). -> list of (begin, end) cell offsets, possibly fewer than n_ranks."""
    n, sep = int(text.size), text[-1]
    b = [0]
    for r in range(1, n_ranks):
        pos = max(r * n // n_ranks, b[-1])
        hit = np.flatnonzero(text[pos:] == sep)
        if hit.size == 0:
            break
        end = pos + int(hit[0]) + 1
        if end >= n:
            break
        if end > b[-1]:
            b.append(end)
    b.append(n)
    return [(b[i], b[i + 1]) for i in range(len(b) - 1)]


def build_bwt_mg(text: np.ndarray, devices, n_threads: int = 1, comm: int = COMM_AUTO, verbose: bool = False):
    """Whole construction over len(devices) ranks inside ONE process (host threads): -> (syms, lens, sb, fb, info).
    info["digests"]: per round [tot_phrases, pre-BWT runs, parse length, distinct phrases, dictionary symbols, 4 checksums]."""
    L = lib_host()
    text = np.ascontiguousarray(text)
    dv = np.asarray(devices, np.int32)
    res = BwtResult()
    rc = L.grlbwt_build_mg(_ptr(text), text.size, text.dtype.itemsize, _ptr(dv), dv.size, n_threads, comm, int(verbose), C.byref(res))
    if rc != 0:
        raise GrlGpuError(rc, L.grlbwt_last_error().decode())
    try:
        syms = np.ctypeslib.as_array(res.syms, shape=(res.n_runs,)).copy()
        lens = np.ctypeslib.as_array(res.lens, shape=(res.n_runs,)).copy()
        info = {k: getattr(res, k) for k in ("n_rounds", "h2d_ms", "par_phase_ms", "ind_phase_ms", "device_ms", "algorithmic_bytes", "induced_on_device")}
        info["digests"] = last_digests()
        info["exchange_bytes"] = int(L.grlbwt_last_exchange_bytes())
        info["comm"] = L.grlbwt_last_comm().decode()
        return syms, lens, int(res.sb), int(res.fb), info
    finally:
        L.grlbwt_free_result(C.byref(res))


def last_digests():
    L = lib_host()
    n = int(L.grlbwt_last_digests(None, 0))
    out = np.zeros((max(n, 1), 9), np.uint64)
    L.grlbwt_last_digests(_ptr(out), n)
    return [[int(x) for x in row] for row in out[:n]]


def check_against_oracle(text: np.ndarray, n_ranks: int, device: int = 0):
    """N in-process ranks on one device vs the oracle: per-round scalars and the final run-length BWT (sanitize driver, tests)"""
    from oracle import oracle as O
    o = O.Oracle(text)
    R = o.par_phase()
    syms, lens, sb, fb, info = build_bwt_mg(text, [device] * n_ranks, n_threads=2, comm=COMM_LOCAL)
    osyms, olens, osb, ofb = o.ind_phase()
    assert info["n_rounds"] == R, (info["n_rounds"], R)
    for lv, d in enumerate(info["digests"]):
        assert d[0] == o.scalar(lv, O.TOT_PHRASES) and d[2] == o.scalar(lv, O.PARSE_LEN), (lv, d)
        assert d[3] == o.scalar(lv, O.D) and d[4] == o.scalar(lv, O.SUM_LEN), (lv, d)
    assert (sb, fb) == (osb, ofb) and np.array_equal(syms, osyms) and np.array_equal(lens, olens)
    o.close()
    return info
