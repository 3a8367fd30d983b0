"""N > 1 ranks: the orchestration of grlbwt_b200/multigpu.py.
CPU (not gpu): world_size 2 and 3 under gloo with the pure-Python test double (tests/cpu_engine.py).
GPU (-m gpu): world_size 2 with the real device engine; NCCL when the box has >= 2 GPUs, else both ranks share
GPU 0 and the collectives are staged through the host under gloo (the device code is the same)."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def test_merge_seams_equals_a_sequential_merge():
    """the O(ranks) merge of per-rank preliminary-BWT slices (two symbols per slice decide) == merging run by run"""
    import numpy as np
    import torch
    from grlbwt_b200.multigpu import merge_seams
    rng = np.random.default_rng(5)
    for _ in range(2000):
        slices = []
        for _r in range(int(rng.integers(1, 7))):
            syms = []
            for _k in range(int(rng.integers(0, 4))):
                x = int(rng.integers(0, 3))
                while syms and syms[-1] == x:
                    x = int(rng.integers(0, 3))
                syms.append(x)
            slices.append((syms, [int(v) for v in rng.integers(1, 9, size=len(syms))]))
        PS = torch.tensor([x for s, _ in slices for x in s] + [99], dtype=torch.int32)   # + padding as alloc() leaves it
        PL = torch.tensor([x for _, l in slices for x in l] + [7], dtype=torch.int64)
        S, L = merge_seams(PS, PL, [len(s) for s, _ in slices])
        es, el = [], []
        for s_, l_ in slices:
            for x, y in zip(s_, l_):
                if es and es[-1] == x:
                    el[-1] += y
                else:
                    es.append(x); el.append(y)
        assert S.tolist() == es and L.tolist() == el


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _case(name):
    import gen
    cases = gen.small_cases()
    if name in cases:
        return cases[name]
    for n, a in gen.fuzz_cases():
        if n == name:
            return a
    if name == "test_2bytes_alphabet":
        return gen.fixture_2bytes_alphabet()
    if name == "rep_50x200k":
        return gen.repetitive_genomes(50, 200000, seed=7)
    if name == "reads_100k":
        return gen.dna_reads(100000, 150, seed=42)
    if name == "u16_2M":
        return gen.int_alphabet(2000000, np.uint16, 65535, 1000, seed=11)
    raise KeyError(name)


def _worker(rank, world, port, backend, engine_kind, names, out_path, dist_rank=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import grlbwt_b200 as G
    from grlbwt_b200 import multigpu as M
    from oracle import oracle as O
    dist.init_process_group(backend, rank=rank, world_size=world)
    results = {}
    try:
        for name in names:
            arr = _case(name)
            if int((arr == arr[-1]).sum()) < world:
                continue  # fewer strings than ranks: nothing to shard
            b, e = M.shard_bounds(arr, world)[rank]
            shard = np.ascontiguousarray(arr[b:e])
            if engine_kind == "cpu":
                from cpu_engine import CpuEngine
                eng, ctx = CpuEngine(shard), None
                eng.distribute = dist_rank
            else:
                dev = rank if backend == "nccl" else 0
                torch.cuda.set_device(dev)
                stream = torch.cuda.Stream(device=dev)  # the library and torch's collectives share one stream
                torch.cuda.set_stream(stream)
                ctx = G.GrlGpu(dev, G.FLAG_FORCE_DIST_RANK if dist_rank else 0, stream=stream.cuda_stream)
                ctx.set_text(shard)
                eng = M.GpuEngine(ctx, torch.device("cuda", dev))
            res = M.par_phase_distributed(eng)
            if rank == 0:
                st = res["stats"]
                levels = [{"alphabet": L["alphabet"], "tot": L["tot"], "rule_l": L["rule_l"], "rule_r": L["rule_r"], "has_hocc": L["has_hocc"],
                           "pre_sym": L["pre_sym"], "pre_len": L["pre_len"]} for L in res["levels"]]
                syms, lens = G.selftest_induce(levels, res["final_parse"])
                sb = -(-int(st["max_sym"] + 4).bit_length() // 8)
                fb = -(-int(st["max_sym_freq"]).bit_length() // 8)
                raw = O.rl_bwt_bytes(syms, lens, sb, fb)
                results[name] = {"sha": hashlib.sha256(raw).hexdigest(), "sb": sb, "fb": fb, "rounds": len(res["rounds"]),
                                 "tot": [r["tot_phrases"] for r in res["rounds"]], "d": [r["n_phrases"] for r in res["rounds"]],
                                 "ranking": [r.get("ranking") for r in res["rounds"]]}
            if ctx is not None:
                ctx.close()
        if rank == 0:
            import json
            with open(out_path, "w") as f:
                json.dump(results, f)
    finally:
        dist.destroy_process_group()


def run_ranks(world, backend, engine_kind, names, tmp_path, dist_rank=False):
    out = str(tmp_path / f"mg_{engine_kind}_{world}_{int(dist_rank)}.json")
    mp.spawn(_worker, args=(world, free_port(), backend, engine_kind, names, out, dist_rank), nprocs=world, join=True)
    import json
    return json.load(open(out))


def check(results, golden):
    for name, r in results.items():
        g = golden[name]
        assert (r["sb"], r["fb"]) == (g["sb"], g["fb"]), name
        assert r["rounds"] == len(g["rounds"]), name
        assert r["tot"] == [x["tot_phrases"] for x in g["rounds"]], name
        assert r["d"] == [x["lms_phrases"] for x in g["rounds"]], name
        assert r["sha"] == g["rl_bwt_sha256"], name


CPU_NAMES = ["dna_500", "ac_short_3000", "with_empty", "u16_small_sigma", "test_2bytes_alphabet", "high_bytes", "fuzz_3", "fuzz_11", "fuzz_42",
             "fuzz_77", "long_phrases"]


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_rounds_gloo_cpu(golden, tmp_path, world):
    names = [n for n in CPU_NAMES if not (world == 3 and n in ("long_phrases",))]
    check(run_ranks(world, "gloo", "cpu", names, tmp_path), golden)


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_ranking_gloo_cpu(golden, tmp_path, world):
    """the ranking itself split over the ranks (rank_sort / rank_apply / all-reduce / rank_finish / level slices)"""
    names = [n for n in CPU_NAMES if not (world == 3 and n in ("long_phrases",))]
    res = run_ranks(world, "gloo", "cpu", names, tmp_path, dist_rank=True)
    check(res, golden)
    assert all(r == "distributed" for v in res.values() for r in v["ranking"])


def test_shard_bounds_whole_strings():
    from grlbwt_b200 import multigpu as M
    rng = np.random.default_rng(0)
    import gen
    arr = gen.random_collection(rng, 3, 50, 40)
    for w in (1, 2, 3, 7):
        b = M.shard_bounds(arr, w)
        assert b[0][0] == 0 and b[-1][1] == arr.size and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all(e > s and arr[e - 1] == arr[-1] for s, e in b)
    with pytest.raises(ValueError):
        M.shard_bounds(np.frombuffer(b"AC\n", np.uint8), 2)


GPU_NAMES = ["dna_500", "mutated_200x5k", "ac_short_3000", "with_empty", "only_empty", "u16_rand", "u32_rand", "u64_rand", "test_2bytes_alphabet",
             "high_bytes", "long_phrases", "homopolymers_multi", "fuzz_3", "fuzz_11", "rep_50x200k", "reads_100k", "u16_2M"]


@pytest.mark.gpu
def test_distributed_rounds_gpu_two_ranks(golden, tmp_path):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    check(run_ranks(2, backend, "gpu", GPU_NAMES, tmp_path), golden)


@pytest.mark.gpu
def test_distributed_ranking_gpu_two_ranks(golden, tmp_path):
    """dictionary ranking partitioned by first-key range over 2 ranks, forced on small dictionaries too"""
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    res = run_ranks(2, backend, "gpu", GPU_NAMES, tmp_path, dist_rank=True)
    check(res, golden)
    assert any(r == "distributed" for v in res.values() for r in v["ranking"])
