"""Generate tests/golden/golden.json by running the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference) on the seeded inputs of tests/gen.py.

Run in the build container only (needs /root/reference for the build, not at test time):
    make -C oracle ref && python tests/golden/make_golden.py
The GPU box and the test-suite only read the committed golden.json + fixture files.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "grlbwt_ref")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
REF_DATA = "/root/reference/test_data"


def run_case(name, arr, threads=1):
    a = arr.dtype.itemsize
    with tempfile.TemporaryDirectory(dir="/tmp") as td:
        inp = os.path.join(td, "in.bin")
        arr.tofile(inp)
        rec = {"name": name, "alph_bytes": a, "n_syms": int(arr.size),
               "input_sha256": hashlib.sha256(arr.tobytes()).hexdigest()}
        r = subprocess.run([REF, inp, "-a", str(a), "-t", str(threads), "-T", td], cwd=td, capture_output=True, text=True)
        out = os.path.join(td, "in.rl_bwt")
        if r.returncode != 0 or not os.path.exists(out):
            rec["reference_failed"] = (r.stdout + r.stderr)[-300:]
            return rec
        raw = open(out, "rb").read()
        rec["rl_bwt_sha256"] = hashlib.sha256(raw).hexdigest()
        rec["rl_bwt_bytes"] = len(raw)
        rec["sb"] = int.from_bytes(raw[0:8], "little")
        rec["fb"] = int.from_bytes(raw[8:16], "little")
        dd = os.path.join(td, "dump")
        os.makedirs(dd)
        r = subprocess.run([HARNESS, "dump", inp, str(a), str(threads), dd], cwd=td, capture_output=True, text=True)
        if r.returncode != 0:
            rec["harness_failed"] = (r.stdout + r.stderr)[-300:]
            return rec
        rounds = []
        for line in open(os.path.join(dd, "rounds.txt")):
            if line.startswith("#"):
                kv = line[1:].split()
                rec["stats"] = {kv[i]: int(kv[i + 1]) for i in range(0, len(kv), 2)}
                continue
            kv = line.split()
            d = {kv[i]: int(kv[i + 1]) for i in range(0, len(kv), 2)}
            k = d["round"]
            d["parse_sha256"] = hashlib.sha256(open(os.path.join(dd, f"parse_r{k}.bin"), "rb").read()).hexdigest()
            sp = np.fromfile(os.path.join(dd, f"str_ptrs_r{k}.bin"), np.int64)
            d["str_ptrs_sha256"] = hashlib.sha256(sp.astype(np.uint64).tobytes()).hexdigest()
            rounds.append(d)
        rec["rounds"] = rounds
        return rec


def main():
    # fixtures copied from the reference's test_data (data, not source)
    gz = os.path.join(HERE, "test_byte_alphabet.txt.gz")
    if not os.path.exists(gz):
        with open(os.path.join(REF_DATA, "test_byte_alphabet.txt"), "rb") as f, gzip.GzipFile(gz, "wb", mtime=0) as g:
            g.write(f.read())
    b2 = os.path.join(HERE, "test_2bytes_alphabet.bin")
    if not os.path.exists(b2):
        shutil.copyfile(os.path.join(REF_DATA, "test_2bytes_alphabet.txt"), b2)

    golden = {"generator": "tests/golden/make_golden.py", "reference": "ddiazdom/grlBWT @ /root/reference (v1.0.1 alpha)",
              "cases": {}}
    cases = {"test_byte_alphabet": gen.fixture_byte_alphabet(), "test_2bytes_alphabet": gen.fixture_2bytes_alphabet()}
    cases.update(gen.small_cases())
    for name, arr in gen.fuzz_cases():
        cases[name] = arr
    cases["reads_100k"] = gen.dna_reads(100000, 150, seed=42)           # 15.1 MB prefix-shaped C2
    cases["rep_50x200k"] = gen.repetitive_genomes(50, 200000, seed=7)   # 10 MB C3-shaped
    cases["u16_2M"] = gen.int_alphabet(2000000, np.uint16, 65535, 1000, seed=11)  # C4-shaped
    cases["mixed_reads"] = gen.mixed_reads(20000, 30, 150, 10000, seed=5)  # C5-shaped
    for name, arr in cases.items():
        rec = run_case(name, arr)
        golden["cases"][name] = rec
        print(name, rec.get("rl_bwt_sha256", "FAILED")[:16], len(rec.get("rounds", [])), flush=True)
    # thread-count independence of the reference itself
    for name in ("test_byte_alphabet", "rep_50x200k"):
        r4 = run_case(name, cases[name], threads=4)
        assert r4["rl_bwt_sha256"] == golden["cases"][name]["rl_bwt_sha256"]
        assert [x["parse_sha256"] for x in r4["rounds"]] == [x["parse_sha256"] for x in golden["cases"][name]["rounds"]]
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
