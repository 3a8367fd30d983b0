"""CPU suite for the N > 1 path: the host-side logic that does not need a device -- the sharding rule of the C++ host
(whole strings, balanced by symbols), the rank-independent generators bench.py shards the workload with, and the
torch.distributed plumbing (gloo, world_size 2) that carries the NCCL id and merges the per-rank digests."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import grlbwt_b200 as G
from grlbwt_b200 import mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_shard_bounds(arr, n_ranks):
    L = G.lib_host()
    L.grlbwt_selftest_shard_bounds.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    out = np.zeros(n_ranks + 1, np.uint64)
    ns = C.c_int()
    rc = L.grlbwt_selftest_shard_bounds(arr.ctypes.data, arr.size, arr.dtype.itemsize, n_ranks, out.ctypes.data, C.byref(ns))
    assert rc == 0
    return [int(x) for x in out[: ns.value + 1]]


@pytest.mark.parametrize("name", ["test_byte_alphabet", "test_2bytes_alphabet", "with_empty", "only_empty", "single_sep", "dna_500", "u32_small_sigma", "u64_rand",
                                  "mixed_reads", "fuzz_3", "fuzz_40"])
def test_shards_are_whole_strings_and_cover_the_text(all_cases, name):
    arr = all_cases[name]
    sep = arr[-1]
    n_strings = int((arr == sep).sum())
    for n_ranks in (1, 2, 3, 4, 8):
        b = c_shard_bounds(arr, n_ranks)
        assert b[0] == 0 and b[-1] == arr.size and all(x < y for x, y in zip(b, b[1:]))
        assert len(b) - 1 <= min(n_ranks, n_strings)
        for e in b[1:]:
            assert arr[e - 1] == sep                       # every shard ends with a separator: whole strings only
        assert [(x, y) for x, y in zip(b, b[1:])] == mg.shard_bounds(arr, n_ranks)   # the Python mirror bench.py / tests use
        if n_strings >= 64 * n_ranks:                        # balanced by symbols when strings are short next to a shard
            sizes = np.diff(b)
            assert sizes.max() <= 1.5 * arr.size / n_ranks + int(np.diff(np.flatnonzero(arr == sep)).max(initial=1))


def test_rank_independent_slices_of_the_bench_generator():
    """bench.py's split rule: N ranks take contiguous read ranges of the SAME global collection"""
    sys.path.insert(0, ROOT)
    import bench
    for total, world in ((10, 3), (50_000_000, 8), (7, 7), (1000, 1)):
        spans = [bench.split_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["GRL_ROOT"])
sys.path.insert(0, os.path.join(os.environ["GRL_ROOT"], "tests"))
import gen
from grlbwt_b200 import mg
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# 1. the 128-byte id travels from rank 0 to everyone (what nccl_comm_from_torch does before ncclCommInitRank)
t = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    t.copy_(torch.arange(128, dtype=torch.uint8))
dist.broadcast(t, src=0)
assert t.tolist() == list(range(128))
# 2. shards of whole strings: every rank derives the same bounds and takes its own range; together they are the text
text = gen.dna_reads(3000, 150, seed=42)
bounds = mg.shard_bounds(text, world)
lo, hi = bounds[rank]
n_local = torch.tensor([hi - lo, int((text[lo:hi] == 10).sum())], dtype=torch.int64)
dist.all_reduce(n_local)
assert n_local.tolist() == [text.size, 3000]
# 3. per-rank digests (sums mod 2^64) merge by addition, whatever the number of ranks
vals = torch.from_numpy((text[lo:hi].astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64).copy())
s = vals.sum().reshape(1)
dist.all_reduce(s)
full = (text.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64).sum()
assert int(s.item()) == int(full)
dist.destroy_process_group()
print("WORKER_OK", rank)
'''


def test_gloo_world_size_2_plumbing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, GRL_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
                        str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-1500:]
    assert r.stdout.count("WORKER_OK") == 2


RENDEZVOUS_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GRL_ROOT"])
from grlbwt_b200.api import lib_gpu
L = lib_gpu()
rc = L.grlgpu_selftest_ipc_rendezvous(os.environ["GRL_SESSION"].encode(), int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))
if rc != 0:
    print("ERR", rc, L.grlgpu_last_error(None).decode())
sys.exit(0 if rc == 0 else 3)
'''


def _run_rendezvous(tmp_path, world, rounds, session, extra_env=None, skip_rank=None, timeout=120):
    script = tmp_path / "rv.py"
    script.write_text(RENDEZVOUS_WORKER)
    env = dict(os.environ, GRL_ROOT=ROOT, GRL_SESSION=session, **(extra_env or {}))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(rounds)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world) if r != skip_rank]
    return [(p.wait(timeout=timeout), p.stdout.read()) for p in procs]


def test_ipc_rendezvous_between_processes(tmp_path):
    """the "ipc" exchange backend's shared-memory segment (barriers + small all-gathers in 64 KB pieces) between 5 processes, no GPU"""
    session = f"/grlgpu-test-{os.getpid()}-a"
    res = _run_rendezvous(tmp_path, 5, 40, session)
    assert all(rc == 0 for rc, _ in res), res
    assert not os.path.exists("/dev/shm" + session)   # rank 0 unlinks the name once everyone is attached


def test_ipc_rendezvous_times_out_when_a_rank_is_missing(tmp_path):
    """a rank that never arrives: the others give up after GRLGPU_IPC_TIMEOUT_S instead of hanging, and report it"""
    session = f"/grlgpu-test-{os.getpid()}-b"
    res = _run_rendezvous(tmp_path, 3, 4, session, extra_env={"GRLGPU_IPC_TIMEOUT_S": "2"}, skip_rank=2)
    assert all(rc == 3 for rc, _ in res), res
    assert any("timed out" in out or "aborted" in out for _, out in res), res
    try:
        os.unlink("/dev/shm" + session)   # rank 0 could not unlink it: not everyone attached
    except OSError:
        pass
