"""GPU suite (-m gpu) for the multi-GPU parse rounds with the PARTITIONED dictionary (grlbwt_b200/csrc/mg2.cuh), through
the C ABI of the host library. On a 1-GPU box the ranks are host threads that share the device and exchange through
in-process peer copies; with >= 2 GPUs the same rounds run over NCCL (one rank per GPU, in-process threads and one
process per GPU). Bar: the bytes of the .rl_bwt and the per-round digests do not depend on the number of ranks, and equal
the reference's goldens."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import grlbwt_b200 as G
from grlbwt_b200 import mg
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def single_gpu_digests(arr):
    G.build_bwt(arr)
    return mg.last_digests()


def check_case(name, arr, g, n_ranks, devices=None, comm=mg.COMM_LOCAL, ref_digests=None):
    syms, lens, sb, fb, info = mg.build_bwt_mg(arr, devices or [0] * n_ranks, n_threads=2, comm=comm)
    tag = f"{name} x{n_ranks}"
    assert (sb, fb) == (g["sb"], g["fb"]), tag
    assert info["n_rounds"] == len(g["rounds"]), tag
    for lv, (d, gr) in enumerate(zip(info["digests"], g["rounds"])):
        assert d[0] == gr["tot_phrases"] and d[3] == gr["lms_phrases"], f"{tag} round {lv + 1}: {d} vs {gr}"
    if ref_digests is not None:
        assert info["digests"] == ref_digests, f"{tag}: per-round digests differ from the 1-GPU run\n{info['digests']}\n{ref_digests}"
    raw = O.rl_bwt_bytes(syms, lens, sb, fb)
    assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"], tag
    return info


def test_mg_reference_fixtures(golden, all_cases):
    for name in ("test_2bytes_alphabet", "test_byte_alphabet"):
        ref = single_gpu_digests(all_cases[name])
        for n in (2, 3, 4):
            check_case(name, all_cases[name], golden[name], n, ref_digests=ref)


def test_mg_corner_cases(golden, all_cases):
    for name, arr in all_cases.items():
        if name.startswith("test_") or name.startswith("fuzz_") or name in ("reads_100k", "rep_50x200k", "u16_2M", "mixed_reads"):
            continue
        ref = single_gpu_digests(arr)
        for n in (2, 3):
            check_case(name, arr, golden[name], n, ref_digests=ref)


def test_mg_fuzz(golden, all_cases):
    for i in range(0, 120, 3):
        name = f"fuzz_{i}"
        check_case(name, all_cases[name], golden[name], 2 + i % 3)


@pytest.mark.parametrize("name", ["reads_100k", "rep_50x200k", "u16_2M", "mixed_reads"])
def test_mg_config_shapes(golden, all_cases, name):
    ref = single_gpu_digests(all_cases[name])
    for n in (2, 4, 8):
        info = check_case(name, all_cases[name], golden[name], n, ref_digests=ref)
        assert info["exchange_bytes"] > 0


def test_mg_more_ranks_than_strings(golden, all_cases):
    """fewer strings than ranks: the host falls back to as many ranks as there are strings"""
    for name in ("mississippi", "single_sep", "homopolymer"):
        check_case(name, all_cases[name], golden[name], 4)


def test_mg_cli(tmp_path, golden, all_cases):
    exe = os.path.join(G.LIB_DIR, "grlbwt")
    for name, a in (("test_byte_alphabet", 1), ("test_2bytes_alphabet", 2)):
        inp = tmp_path / (name + ".txt")
        all_cases[name].tofile(inp)
        r = subprocess.run([exe, str(inp), "-a", str(a), "-t", "4", "-T", str(tmp_path), "-g", "0,0,0", "--comm", "local"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert hashlib.sha256((tmp_path / (name + ".rl_bwt")).read_bytes()).hexdigest() == golden[name]["rl_bwt_sha256"]


def n_gpus():
    import torch
    return torch.cuda.device_count()


def test_mg_nccl_threads(golden, all_cases):
    """one rank per GPU inside one process, NCCL grouped send/recv"""
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    k = min(n_gpus(), 8)
    for name in ("test_byte_alphabet", "reads_100k", "rep_50x200k", "u16_2M", "mutated_200x5k"):
        ref = single_gpu_digests(all_cases[name])
        info = check_case(name, all_cases[name], golden[name], k, devices=list(range(k)), comm=mg.COMM_NCCL, ref_digests=ref)
        assert "NCCL" in info["comm"]
        check_case(name, all_cases[name], golden[name], k, devices=list(range(k)), comm=mg.COMM_LOCAL, ref_digests=ref)   # NVLink peer copies


NCCL_WORKER = r'''
import hashlib, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.environ["GRL_ROOT"]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen
import grlbwt_b200 as G
from grlbwt_b200 import mg
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
kind = os.environ["GRL_COMM"]
if kind == "ipc_shared":   # every rank on GPU 0 (the 1-GPU box): torch.distributed over gloo, exchanges through CUDA IPC
    lr = 0
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
    comm = mg.comm_from_torch(dist, rank, world, 0, torch, kind="ipc")
else:
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = mg.comm_from_torch(dist, rank, world, lr, torch, kind=kind)
assert ("IPC" in comm.info()["kind"]) == kind.startswith("ipc"), comm.info()
arr = {"reads": lambda: gen.dna_reads(100000, 150, seed=42), "rep": lambda: gen.repetitive_genomes(50, 200000, seed=7)}[os.environ["GRL_CASE"]]()
lo, hi = mg.shard_bounds(arr, world)[rank]
ctx = G.GrlGpu(lr)
ctx.set_text(arr[lo:hi])
ctx.mg_stats(comm)
digests, cs_all = [], []
while True:
    r = ctx.mg_round(comm)
    cs = torch.tensor(np.array(ctx.mg_slice_checksum(), np.uint64).view(np.int64), device="cuda" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(cs)
    digests.append([r.tot_phrases, r.n_pre_runs, r.parse_len, r.n_phrases, r.dict_syms] + [int(x) for x in cs.cpu().numpy().view(np.uint64)])
    sl = ctx.mg_slice_info()
    if r.done:
        part = ctx.fetch_parse_local(sl.parse_len_local).astype(np.uint64)
        break
parts = [None] * world
dist.all_gather_object(parts, part)
if rank == 0:
    print("RESULT " + json.dumps({"digests": digests, "parse_sha": hashlib.sha256(np.concatenate(parts).tobytes()).hexdigest()}))
comm.close(); ctx.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("case", ["reads", "rep"])
@pytest.mark.parametrize("kind", ["nccl", "ipc", "ipc_shared"])
def test_mg_process_per_gpu(tmp_path, case, kind):
    """bench.py's arrangement: torchrun, one process per GPU; the NCCL id / the name of the IPC rendezvous segment is broadcast
    through torch.distributed. "ipc_shared": three processes share GPU 0, so the IPC backend is also covered on a 1-GPU box."""
    if kind != "ipc_shared" and n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    import gen
    import json
    arr = {"reads": lambda: gen.dna_reads(100000, 150, seed=42), "rep": lambda: gen.repetitive_genomes(50, 200000, seed=7)}[case]()
    ref = single_gpu_digests(arr)
    with G.GrlGpu(0) as ctx:
        ctx.set_text(arr)
        while True:
            r = ctx.round()
            if r.done:
                ref_sha = hashlib.sha256(ctx.fetch_parse().astype(np.uint64).tobytes()).hexdigest()
                break
    script = tmp_path / "w.py"
    script.write_text(NCCL_WORKER)
    k = 3 if kind == "ipc_shared" else min(n_gpus(), 8)
    env = dict(os.environ, GRL_ROOT=ROOT, GRL_CASE=case, GRL_COMM=kind, GRLGPU_IPC_TIMEOUT_S="120")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(k), "--master-addr", "127.0.0.1", "--master-port", "29741",
                        str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][0]
    got = json.loads(line[7:])
    assert got["digests"] == ref
    assert got["parse_sha"] == ref_sha
