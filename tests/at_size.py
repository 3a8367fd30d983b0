#!/usr/bin/env python
"""At-size verification through the C++ CLI (SURVEY.md 8d "parity at sizes the oracle cannot
reach"). For every BASELINE.json config at its stated size (or the largest prefix the box's RAM allows):
    generator (tests/gen.py, fixed seeds) -> file -> grlbwt CLI (-a/-b/-t as the config says) -> .rl_bwt
    -> bwt_check -k 1000 (header rule, sum of run lengths, maximal runs, per-symbol totals, separator block in string
       order, LF-inversion of 1000 strings against the text)
and, on a >= 300 MB prefix of every shape, sha256(.rl_bwt) against the UNMODIFIED reference (oracle/_ref/grlbwt_ref)
run on the same box. Needs a GPU; run on the B200 box:  gpurun -- python tests/at_size.py --out gpurun_out/at_size.json
Nothing here is imported by the product; the reference binary is only executed as the checker."""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402

from oracle import refbin  # noqa: E402

LIB = os.path.join(ROOT, "grlbwt_b200", "lib")
REF = refbin.ref_path("grlbwt_ref")


def mem_available_gb():
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            return int(ln.split()[1]) / 1e6
    return 0.0


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def write_c2(path, reads):
    """C2 generator streamed to disk in blocks (same bytes as gen.dna_reads(reads, 150, 42): one RNG, 2^20-read chunks)"""
    rng = np.random.default_rng(42)
    with open(path, "wb") as f:
        chunk = 1 << 20
        for i in range(0, reads, chunk):
            j = min(reads, i + chunk)
            out = np.empty((j - i, 151), np.uint8)
            out[:, :150] = gen.DNA[rng.integers(0, 4, size=(j - i, 150), dtype=np.uint8)]
            out[:, 150] = 10
            out.tofile(f)


def make_input(name, path, scale):
    t0 = time.time()
    if name == "c2":
        write_c2(path, scale)
    elif name == "c3":
        gen.repetitive_genomes(scale, 4_000_000, seed=7).tofile(path)
    elif name == "c4":
        gen.int_alphabet(scale, np.uint16, 65535, 1000, seed=11).tofile(path)
    elif name == "c5":
        gen.mixed_reads(scale, max(1, scale * 150 // 10 // 10000), 150, 10000, seed=5).tofile(path)
    else:
        raise ValueError(name)
    return time.time() - t0


CLI_FLAGS = {"c2": ["-a", "1"], "c3": ["-a", "1", "-b", "2"], "c4": ["-a", "2", "-b", "2"], "c5": ["-a", "1"]}


def run_ours(name, inp, workdir, threads, gpus, timeout):
    out = os.path.join(workdir, f"{name}_ours.rl_bwt")
    cmd = [os.path.join(LIB, "grlbwt"), inp, "-o", out, "-t", str(threads), "-T", workdir] + CLI_FLAGS[name]
    if gpus > 1:
        cmd += ["--gpus", str(gpus)]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=workdir)
    dt = time.time() - t0
    return out, r, dt


def run_check(name, inp, rl, timeout):
    a = CLI_FLAGS[name][1]
    t0 = time.time()
    r = subprocess.run([os.path.join(LIB, "bwt_check"), inp, rl, "-a", a, "-k", "1000"], capture_output=True, text=True, timeout=timeout)
    return r, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c3,c4,c2")
    ap.add_argument("--prefix-configs", default="c2,c3,c4,c5", help="shapes whose >=300 MB prefix is compared with the reference's sha256")
    ap.add_argument("--c2-reads", type=int, default=0, help="0: as many of the 50M reads as the host RAM allows")
    ap.add_argument("--c3-copies", type=int, default=1000)
    ap.add_argument("--c4-cells", type=int, default=1_000_000_000)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--scratch", default="")
    ap.add_argument("--out", default="gpurun_out/at_size.json")
    ap.add_argument("--timeout", type=int, default=1500)
    args = ap.parse_args()

    threads = os.cpu_count() or 1
    avail = mem_available_gb()
    scratch = args.scratch or ("/dev/shm" if shutil.disk_usage("/dev/shm").free > 80e9 else "/tmp")
    work = tempfile.mkdtemp(prefix="at_size_", dir=scratch)
    report = {"host": {"cores": threads, "mem_available_gb": round(avail, 1), "scratch": scratch,
                       "scratch_free_gb": round(shutil.disk_usage(scratch).free / 1e9, 1)}, "at_size": [], "vs_reference": []}
    print(json.dumps(report["host"]), flush=True)

    # ---- reference runs on >= 300 MB prefixes, in the background (they are CPU-only and mostly serial) ----
    prefix_scale = {"c2": 2_000_000, "c3": 100, "c4": 150_000_000, "c5": 2_000_000}
    ref_threads, ref_results = [], {}

    def ref_job(name):
        inp = os.path.join(work, f"{name}_prefix.txt")
        gen_s = make_input(name, inp, prefix_scale[name])
        entry = {"config": name, "scale": prefix_scale[name], "bytes": os.path.getsize(inp), "gen_s": round(gen_s, 1)}
        try:
            out_ref = os.path.join(work, f"{name}_ref.rl_bwt")
            t0 = time.time()
            r = subprocess.run([REF, inp, "-o", out_ref, "-t", "4", "-T", work] + CLI_FLAGS[name], capture_output=True, text=True, timeout=args.timeout, cwd=work)
            entry["reference_s"] = round(time.time() - t0, 1)
            entry["reference_rc"] = r.returncode
            if r.returncode == 0:
                entry["reference_sha256"] = sha256_file(out_ref)
                entry["rl_bwt_bytes"] = os.path.getsize(out_ref)
                os.remove(out_ref)
        except Exception as e:  # noqa: BLE001
            entry["reference_error"] = str(e)[:200]
        ref_results[name] = entry

    report["host"]["reference_build"] = refbin.ref_flags()
    if not args.no_reference and REF:
        for name in [c for c in args.prefix_configs.split(",") if c]:
            th = threading.Thread(target=ref_job, args=(name,))
            th.start()
            ref_threads.append(th)

    # ---- stated sizes through our CLI + bwt_check ----
    for name in [c for c in args.configs.split(",") if c]:
        if name == "c2":
            scale = args.c2_reads or int(min(50_000_000, max(2_000_000, (mem_available_gb() - 40) * 1e9 / 14000)))
        elif name == "c3":
            scale = args.c3_copies
        elif name == "c4":
            scale = args.c4_cells
        else:
            scale = 2_000_000
        inp = os.path.join(work, f"{name}.txt")
        entry = {"config": name, "scale": scale}
        try:
            entry["gen_s"] = round(make_input(name, inp, scale), 1)
            entry["bytes"] = os.path.getsize(inp)
            out, r, dt = run_ours(name, inp, work, threads, args.gpus, args.timeout)
            entry["cli_s"] = round(dt, 2)
            entry["cli_rc"] = r.returncode
            entry["MBps_file_to_file"] = round(entry["bytes"] / 1e6 / dt, 1)
            entry["cli_tail"] = (r.stdout + r.stderr)[-1500:]
            timing = [ln for ln in r.stdout.splitlines() if ln.startswith("Timing (ms):")]
            entry["cli_timing"] = timing[-1] if timing else None   # text to device / parse phase / induction / writing, as the tool reports them
            if r.returncode == 0:
                entry["rl_bwt_bytes"] = os.path.getsize(out)
                with open(out, "rb") as f:
                    hdr = np.frombuffer(f.read(16), np.uint64)
                entry["header_sb_fb"] = [int(hdr[0]), int(hdr[1])]
                entry["rl_bwt_sha256"] = sha256_file(out)
                c, cdt = run_check(name, inp, out, args.timeout)
                entry["bwt_check"] = c.stdout.strip()[-300:]
                entry["bwt_check_rc"] = c.returncode
                entry["bwt_check_s"] = round(cdt, 1)
                os.remove(out)
        except Exception as e:  # noqa: BLE001
            entry["error"] = str(e)[:300]
        if os.path.exists(inp):
            os.remove(inp)
        report["at_size"].append(entry)
        print(json.dumps(entry), flush=True)

    # ---- the same prefixes through our CLI, compared with the reference ----
    for th in ref_threads:
        th.join()
    for name, entry in ref_results.items():
        inp = os.path.join(work, f"{name}_prefix.txt")
        try:
            out, r, dt = run_ours(name, inp, work, threads, args.gpus, args.timeout)
            entry["ours_s"] = round(dt, 2)
            entry["ours_rc"] = r.returncode
            if r.returncode == 0:
                entry["ours_sha256"] = sha256_file(out)
                entry["identical"] = entry.get("reference_sha256") == entry["ours_sha256"]
            else:
                entry["ours_tail"] = (r.stdout + r.stderr)[-600:]
        except Exception as e:  # noqa: BLE001
            entry["ours_error"] = str(e)[:200]
        report["vs_reference"].append(entry)
        print(json.dumps(entry), flush=True)

    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    ok = all(e.get("bwt_check_rc") == 0 for e in report["at_size"]) and all(e.get("identical") for e in report["vs_reference"])
    print("AT_SIZE", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
