#!/usr/bin/env python
"""bench.py -- parse-phase throughput of the B200 path (BASELINE.json metric: input MB/s to BCR BWT;
parse-round HBM GB/s vs peak) on the C2 workload: synthetic DNA reads, 150 bp, '\\n'-separated.

A "step" = one pass of the hot path over the collection: collection statistics + every parse round
(boundary scan, dedup, dictionary ranking, pre-BWT/grammar, rewrite) until each string is one
metasymbol, i.e. what exact_algo::par_phase does in the reference.
  value : text already resident in HBM (grlgpu_set_text_device), nothing copied.
  e2e   : through the C ABI with HOST buffers: pinned text -> device every step, every level's
          artefacts (rules, hocc marks, preliminary BWT) and the final parse copied back to host.
  --impl reference : the UNMODIFIED reference parse phase (oracle/_ref/ref_harness = its par_phase,
          built from /root/reference by oracle/Makefile) on the host cores, on a bounded sample.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "input MB/s to BCR BWT (parse phase: all rounds on device)"
READ_LEN = 150


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=50_000_000, help="reads of the collection (C2: 50M x 150 bp = 7.55 GB)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"], help="c2: random reads (the metric's config); c3: repetitive genomes (BASELINE.json configs[2])")
    ap.add_argument("--copies", type=int, default=1000, help="c3: genome copies")
    ap.add_argument("--genome", type=int, default=4_000_000, help="c3: genome length")
    ap.add_argument("--sample-reads", type=int, default=100_000, help="reads of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------- reference / CPU baseline
def cpu_parse_phase(sample_reads, threads, repeats=1):
    """time the reference's own parse phase on `sample_reads` C2-shaped reads; -> (MB/s, kind, sample description)"""
    import numpy as np
    import gen
    arr = gen.dna_reads(sample_reads, READ_LEN, seed=42)
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    desc = f"{sample_reads} reads x {READ_LEN} bp ({arr.nbytes / 1e6:.1f} MB) of the same generator"
    secs = []
    if os.path.exists(harness):
        with tempfile.TemporaryDirectory(dir="/tmp") as td:
            inp = os.path.join(td, "sample.txt")
            arr.tofile(inp)
            for _ in range(repeats):
                r = subprocess.run([harness, "parse", inp, "1", str(threads), td], capture_output=True, text=True, cwd=td)
                t = [ln for ln in r.stdout.splitlines() if ln.startswith("PAR_PHASE_SECONDS")]
                if r.returncode != 0 or not t:
                    raise RuntimeError("reference harness failed: " + (r.stdout + r.stderr)[-400:])
                secs.append(float(t[0].split()[1]))
        return [arr.nbytes / 1e6 / s for s in secs], "reference", desc + f"; reference par_phase, -t {threads}", threads
    from oracle import oracle as O  # port (single thread)
    for _ in range(repeats):
        t0 = time.time()
        o = O.Oracle(arr)
        o.par_phase()
        secs.append(time.time() - t0)
        o.close()
    return [arr.nbytes / 1e6 / s for s in secs], "port", desc + "; oracle/oracle.c restatement, 1 thread", 1


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    total = args.warmup + args.steps
    rates, kind, desc, cores = cpu_parse_phase(args.sample_reads, threads, repeats=total)
    timed = rates[args.warmup:]
    # the reference's whole program (parse + induction + output) on the same sample, once, for context
    bwt_total = None
    exe = os.path.join(ROOT, "oracle", "_ref", "grlbwt_ref")
    if os.path.exists(exe):
        import gen
        arr = gen.dna_reads(args.sample_reads, READ_LEN, seed=42)
        with tempfile.TemporaryDirectory(dir="/tmp") as td:
            inp = os.path.join(td, "sample.txt")
            arr.tofile(inp)
            t0 = time.time()
            r = subprocess.run([exe, inp, "-t", str(threads), "-T", td], cwd=td, capture_output=True, text=True)
            dt = time.time() - t0
            if r.returncode == 0:
                bwt_total = {"value": round(arr.nbytes / 1e6 / dt, 3), "unit": "MB/s", "sample": desc, "what": "reference grlbwt CLI, text file -> .rl_bwt file"}
    sample_mb = args.sample_reads * (READ_LEN + 1) / 1e6
    ms = [sample_mb / r * 1e3 for r in timed]
    val = sample_mb * len(timed) / (sum(ms) / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 3), "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(sum(ms) / len(ms), 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": {"workload": f"C2-shaped sample: {desc}", "timing": "host wall clock inside the harness"},
            "cpu_baseline": {"value": round(val, 3), "unit": "MB/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": round(val, 3), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "bwt_total": bwt_total}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------- our arm
def make_reads_on_device(torch, n_reads, seed, device):
    """C2 generator on the device (uniform ACGT + '\\n'); same shape as tests/gen.dna_reads, torch RNG"""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    out = torch.empty((n_reads, READ_LEN + 1), dtype=torch.uint8, device=device)
    chunk = 2_000_000
    for i in range(0, n_reads, chunk):
        j = min(n_reads, i + chunk)
        idx = torch.randint(0, 4, (j - i, READ_LEN), generator=g, device=device, dtype=torch.int64)
        out[i:j, :READ_LEN] = lut[idx]
        del idx
    out[:, READ_LEN] = 10
    return out.reshape(-1)


def make_genomes_on_device(torch, n_copies, genome_len, seed, device, snp=1e-3, dele=1e-4):
    """C3 generator on the device: copies of one random genome with substitutions and single-base deletions
    (same shape as tests/gen.repetitive_genomes; every rank derives the SAME base genome from `seed`)"""
    gb = torch.Generator(device=device)
    gb.manual_seed(7)
    base = torch.randint(0, 4, (genome_len,), generator=gb, device=device, dtype=torch.int64)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=device)
    nl = torch.tensor([10], dtype=torch.uint8, device=device)
    parts = []
    for _ in range(n_copies):
        s = base.clone()
        m = torch.rand(genome_len, generator=g, device=device) < snp
        s[m] = (s[m] + torch.randint(1, 4, (int(m.sum()),), generator=g, device=device)) % 4
        keep = torch.rand(genome_len, generator=g, device=device) >= dele
        parts.append(lut[s[keep]])
        parts.append(nl)
    return torch.cat(parts)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import grlbwt_b200 as G
    from grlbwt_b200 import multigpu as M

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the parse phase has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json (measured copy)") if "hbm_gbs" in peaks else (6650.0, "fallback 6.65 TB/s")

    # strong scaling: the C2 collection (args.reads reads) is split into contiguous ranges of whole reads, one per rank
    if args.workload == "c2":
        my_reads = args.reads // world + (1 if rank < args.reads % world else 0)
        text = make_reads_on_device(torch, my_reads, 42 + rank, dev)
        n = text.numel()
        n_total = args.reads * (READ_LEN + 1)
        wl_desc = f"C2: {args.reads} reads x {READ_LEN} bp uniform ACGT + newline ({n_total / 1e9:.3f} GB), BASELINE.json configs[1]"
    else:
        my_copies = args.copies // world + (1 if rank < args.copies % world else 0)
        text = make_genomes_on_device(torch, my_copies, args.genome, 1000 + rank, dev)
        n = text.numel()
        tot = torch.tensor([n], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        n_total = int(tot.item())
        wl_desc = f"C3: {args.copies} copies of a {args.genome} bp genome, 0.1% SNPs + 0.01% deletions ({n_total / 1e9:.3f} GB), BASELINE.json configs[2]"
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)   # the library issues every kernel on this stream, so torch events bracket it
    torch.cuda.set_stream(stream)
    ctx = G.GrlGpu(local_rank, 0, stream=stream.cuda_stream)
    engine = M.GpuEngine(ctx, dev)
    arena = [None]
    e2e_parts = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_phase(fetch, collect=None):
        """all rounds on the text set in ctx; fetch: copy every level's artefacts + the final parse to the host"""
        d2h, a_off = 0, 0
        if world == 1:
            ctx.stats()
            while True:
                r = ctx.round()
                if collect is not None:
                    collect.append(r.as_dict())
                if fetch:  # every level lands in its own slice of the pinned arena while the next round computes
                    e2e_parts.setdefault("round_ms", []).append(round(r.device_ms, 1))
                    ctx.fetch_level(arena[0], async_=True, offset=a_off, narrow_len=True)  # 32-bit run lengths where they fit, as the C++ host does
                    a_off = ctx.arena_end
                    d2h += r.tot_phrases * (2 * r.sym_bytes + 1) + r.n_pre_runs * (r.sym_bytes + (4 if r.n_in + r.parse_len < (1 << 32) else 8))
                if r.done:
                    if fetch:
                        t_tail = time.perf_counter()
                        ctx.fetch_wait()
                        d2h += ctx.fetch_parse(arena[0][a_off:]).nbytes
                        e2e_parts["tail_wait_ms"] = (time.perf_counter() - t_tail) * 1e3
                    return d2h
        st = M.global_stats(engine)
        while True:
            info, done = M.distributed_round(engine, st["n_strings"], want_level=fetch)
            if collect is not None:
                collect.append(info)
            if fetch and rank == 0:
                engine.fetch_level(arena[0])
                d2h += info["tot_phrases"] * (2 * info["sym_bytes"] + 1) + info["n_pre_runs"] * (info["sym_bytes"] + 8)
            if done:
                if fetch:
                    fp = M.gather_final_parse(engine)
                    d2h += 0 if fp is None else fp.nbytes
                return d2h

    def step_resident(collect=None):
        ctx.set_text_device(text.data_ptr(), n, 1)
        run_phase(False, collect)

    # ---- value: text resident in HBM ----
    for _ in range(args.warmup):
        step_resident()
    rounds_info = []
    step_resident(rounds_info)           # untimed: per-round figures
    ctx.profile_reset(); ctx.profile_enable(True)
    step_resident()                      # untimed: per-kernel CUDA-event timing enabled
    prof = ctx.profile()
    ctx.profile_enable(False); ctx.profile_reset()

    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = n_total * args.steps / 1e6 / (ms_total / 1e3)

    # ---- e2e: host buffers through the C ABI ----
    e2e = None
    if not args.no_e2e:
        host_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        host_text.copy_(text)
        torch.cuda.synchronize()
        host_np = host_text.numpy()
        # pinned landing zone for the level artefacts (largest level of the workload, measured in the resident steps)
        need = sum(r["tot_phrases"] * 17 + r["n_pre_runs"] * 16 + 256 for r in rounds_info) + rounds_info[-1]["parse_len"] * 8 * world + (1 << 20)
        if world > 1 and rank != 0:
            need = 1 << 20  # the levels of a multi-rank job are assembled and fetched on rank 0 only
        arena[0] = torch.empty(int(need), dtype=torch.uint8, pin_memory=True).numpy()
        d2h_bytes = [0]

        def step_e2e():
            e2e_parts.clear()
            t_a = time.perf_counter()
            ctx.set_text(host_np)
            t_b = time.perf_counter()
            d2h_bytes[0] = run_phase(True)
            e2e_parts["h2d_ms"] = (t_b - t_a) * 1e3
            e2e_parts["rounds_and_fetch_ms"] = (time.perf_counter() - t_b) * 1e3

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([wall_ms, float(n), float(d2h_bytes[0])], device=dev, dtype=torch.float64)
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            wall_ms, h2d_b, d2h_b = float(tm[0]), int(t[1]), int(t[2])
        else:
            h2d_b, d2h_b = n, d2h_bytes[0]
        e2e = {"value": round(n_total * args.steps / 1e6 / (wall_ms / 1e3), 3), "unit": "MB/s", "h2d_bytes_per_step": int(h2d_b),
               "d2h_bytes_per_step": int(d2h_b), "ms_per_step": round(wall_ms / args.steps, 3),
               "timing": "host wall clock between stream synchronisations, max over ranks (fetches are host-blocking)",
               "last_step_parts_ms": {k: (v if isinstance(v, list) else round(v, 1)) for k, v in e2e_parts.items()}}
    ctx.close()

    # ---- roofline of the dominant kernel (CUDA events on the launch stream, live, one profiled step) ----
    roof = None
    if prof:
        name, (nl, ms, by) = max(prof.items(), key=lambda kv: kv[1][1])
        achieved = by / 1e9 / (ms / 1e3) if ms > 0 else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
        if os.path.exists(tpath) and args.workload == "c2" and args.reads == 50_000_000 and world == 1:
            tj = json.load(open(tpath))
            if tj.get("kernel") == name:   # ncu DRAM bytes of the same kernel on the same workload, per launch
                traffic, traffic_src = round(tj["traffic_bytes_per_launch"]), "profiles/r01_ncu_traffic.json (ncu dram__bytes_read+write, mean over the step's launches)"
        roof = {"bound": "hbm", "kernel": name, "launches_per_step": nl, "avg_launch_ms": round(ms / max(nl, 1), 4), "achieved": round(achieved, 1),
                "peak": hbm_peak, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "bytes_per_launch": round(by / max(nl, 1)), "peak_source": peak_src,
                "share_of_step": round(ms / sum(v[1] for v in prof.values()), 4),
                "bytes_model": "expected DRAM bytes of the kernel's launches (SURVEY.md 8d traffic table; DESIGN.md kernels section)"}
    alg_bytes = sum(r["algorithmic_bytes"] for r in rounds_info)
    round_ms = sum(r["device_ms"] for r in rounds_info)
    keys = ("round", "n_in", "parse_len", "n_phrases", "dict_syms", "tot_phrases", "n_pre_runs", "device_ms", "text_pass_ms", "dict_ms", "rewrite_ms",
            "algorithmic_bytes", "exchange_bytes_sent", "gather_bytes", "ranking")
    parse_rounds = {"algorithmic_GB": round(alg_bytes / 1e9, 3), "device_ms": round(round_ms, 3),
                    "achieved_GBps": round(alg_bytes / 1e6 / round_ms, 1) if round_ms else None,
                    "frac_of_measured_peak": round(alg_bytes / 1e6 / round_ms / hbm_peak, 4) if round_ms else None,
                    "frac_of_nominal_8TBps": round(alg_bytes / 1e6 / round_ms / 8000.0, 4) if round_ms else None,
                    "scope": "rank 0" if world > 1 else "whole job",
                    "per_round": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in keys} for r in rounds_info]}
    kernels = {k: {"launches": v[0], "ms": round(v[1], 3), "model_GBps": round(v[2] / 1e6 / v[1], 1) if v[1] > 0 else None}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}

    # ---- whole construction (device parse phase + multi-threaded host induction) on a bounded sample, for context ----
    bwt_total = None
    if rank == 0 and world == 1 and not args.no_e2e and args.workload == "c2":
        try:
            import gen
            sample = gen.dna_reads(min(args.reads, 2_000_000), READ_LEN, seed=42)
            thr = os.cpu_count() or 1
            G.build_bwt(sample[: 151 * 1000], n_threads=thr)  # warm the host library
            _, lens_, _, _, info = G.build_bwt(sample, device=local_rank, n_threads=thr)
            tot_ms = info["h2d_ms"] + info["par_phase_ms"] + info["ind_phase_ms"]
            bwt_total = {"value": round(sample.nbytes / 1e6 / (tot_ms / 1e3), 3), "unit": "MB/s", "sample": f"{sample.size // 151} reads x {READ_LEN} bp ({sample.nbytes / 1e6:.0f} MB)",
                         "h2d_ms": round(info["h2d_ms"], 1), "parse_phase_ms": round(info["par_phase_ms"], 1), "induction_ms": round(info["ind_phase_ms"], 1),
                         "host_threads": thr, "bwt_runs": int(lens_.size), "what": "host text -> run-length BCR BWT in host memory (grlbwt_build), file I/O excluded"}
        except Exception as e:
            bwt_total = {"value": None, "error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rates, kind, desc, cores = cpu_parse_phase(args.sample_reads, os.cpu_count() or 1, repeats=1)
            cpu = {"value": round(rates[0], 3), "unit": "MB/s", "cores": cores, "kind": kind, "sample": desc}
        except Exception as e:  # the baseline is reporting only
            cpu = {"value": None, "unit": "MB/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": wl_desc,
                           "reads": args.reads if args.workload == "c2" else None, "cache": "the text of every round-1 pass (7.55 GB at the default size) exceeds the 126 MB L2",
                           "parallelism": "1 GPU" if world == 1 else
                           f"{world} ranks: contiguous ranges of whole reads per rank; per round one hash-partitioned all-to-all-v of the local "
                           f"dictionaries + one all-gather-v of the deduplicated global dictionary over NCCL; large dictionaries are ranked "
                           f"distributed (suffix entries partitioned by first-key range, three all-reduces, metasymbols returned to the "
                           f"requesters by a reverse all-to-all-v), small ones replicated"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "bwt_total": bwt_total,
                "parse_rounds": parse_rounds, "kernels": kernels}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
