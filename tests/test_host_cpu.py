"""CPU suite: host-side logic (induction phase, .rl_bwt writer, CLI argument handling) and the C-ABI
surface (libraries load and export every declared symbol; no compute calls without a GPU)."""
import ctypes
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import grlbwt_b200 as G
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_levels(o, R):
    levels = []
    for lv in range(R):
        levels.append({"alphabet": o.scalar(lv, O.ALPHABET), "tot": o.scalar(lv, O.TOT_PHRASES),
                       "rule_l": o.array(lv, O.A_RULE_L), "rule_r": o.array(lv, O.A_RULE_R),
                       "has_hocc": o.array(lv, O.A_HAS_HOCC).astype(np.uint8),
                       "pre_sym": o.array(lv, O.A_PRE_SYM), "pre_len": o.array(lv, O.A_PRE_LEN)})
    return levels


def test_host_induction_matches_oracle_and_reference(golden, all_cases):
    """ind_phase.hpp fed with the oracle's level artefacts must give the reference's .rl_bwt."""
    names = ["test_byte_alphabet", "test_2bytes_alphabet", "mutated_200x5k", "ac_short_3000", "u16_small_sigma", "with_empty",
             "long_phrases", "high_bytes"] + [f"fuzz_{i}" for i in range(0, 120, 7)]
    for name in names:
        arr = all_cases[name]
        o = O.Oracle(arr)
        R = o.par_phase()
        levels, fp, g = oracle_levels(o, R), o.array(R - 1, O.A_PARSE), golden[name]
        wide = any(int(L["rule_l"].max(initial=0)) >= 2**32 or int(L["rule_r"].max(initial=0)) >= 2**32 for L in levels)
        for threads in ((0,) if wide else (0, 1, 4)):   # 0 = sequential 64-bit path, else ind_phase_mt.hpp
            syms, lens = G.selftest_induce(levels, fp, threads)
            raw = O.rl_bwt_bytes(syms, lens, g["sb"], g["fb"])
            assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"], (name, threads)


@pytest.mark.parametrize("name", ["reads_100k", "rep_50x200k", "u16_2M", "mixed_reads"])
def test_multithreaded_induction_config_shapes(golden, all_cases, name):
    """every parallel step (tuple emission, radix distribution, prefix sums, range assembly, boundary merging) on inputs
    large enough to split across threads"""
    o = O.Oracle(all_cases[name])
    R = o.par_phase()
    levels, fp, g = oracle_levels(o, R), o.array(R - 1, O.A_PARSE), golden[name]
    for threads in (3, 8):
        syms, lens = G.selftest_induce(levels, fp, threads)
        raw = O.rl_bwt_bytes(syms, lens, g["sb"], g["fb"])
        assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"], (name, threads)


@pytest.mark.parametrize("name", ["rep_50x200k", "mixed_reads"])
def test_induction_with_more_threads_than_chunks(golden, all_cases, name):
    """thread counts far above what the small levels can feed: chunks past the end of an array are empty and must not
    touch the per-chunk partial sums (levels with 4096..T^2 runs used to lose the last chunk's total)"""
    o = O.Oracle(all_cases[name])
    R = o.par_phase()
    levels, fp, g = oracle_levels(o, R), o.array(R - 1, O.A_PARSE), golden[name]
    for threads in (96, 128, 258):
        syms, lens = G.selftest_induce(levels, fp, threads)
        raw = O.rl_bwt_bytes(syms, lens, g["sb"], g["fb"])
        assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"], (name, threads)


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(grl(?:gpu|bwt)_[a-z_0-9]+)\s*\(", src)))


def test_c_abi_exports_every_declared_symbol():
    gpu = ctypes.CDLL(os.path.join(G.LIB_DIR, "libgrlgpu.so"))
    names = declared_symbols("grlgpu.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(gpu, n), n
    host = ctypes.CDLL(os.path.join(G.LIB_DIR, "libgrlbwt.so"))
    for n in declared_symbols("grlbwt.h"):
        assert hasattr(host, n), n


def test_status_strings_and_argument_errors():
    L = G.lib_gpu()
    assert L.grlgpu_strerror(0) == b"ok"
    assert b"ill formed" in L.grlgpu_strerror(-2)
    assert L.grlgpu_create(None, 0, 0) == -1            # null out-pointer
    assert L.grlgpu_round(None, None) == -1
    assert L.grlgpu_stats(None, None) == -1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(G.GrlGpuError) as e:
        G.GrlGpu(0)
    assert e.value.status == -3
    with pytest.raises(G.GrlGpuError):
        G.build_bwt(np.frombuffer(b"ACGT\n", np.uint8))


def test_cli_flags_and_naming(tmp_path):
    exe = os.path.join(G.LIB_DIR, "grlbwt")
    assert os.path.exists(exe) and os.path.exists(os.path.join(G.LIB_DIR, "grlbwt-cli"))
    r = subprocess.run([exe, "-v"], capture_output=True, text=True)
    assert r.returncode == 0 and "v1.0.1" in r.stdout
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 106                           # TEXT is required (CLI11 RequiredError)
    r = subprocess.run([exe, str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 105                           # ExistingFile validator
    f = tmp_path / "x.txt"
    f.write_bytes(b"ACGT\n")
    for bad in (["-a", "3"], ["-b", "6"], ["-f", "1.5"], ["-T", str(tmp_path / "nodir")]):
        r = subprocess.run([exe, str(f)] + bad, capture_output=True, text=True)
        assert r.returncode == 105, bad
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    for flag in ("-o,--output-file", "-a,--alphabet", "-t,--threads", "-f,--hbuff", "-b,--run-len-bytes", "-T,--tmp", "-v,--version"):
        assert flag in r.stdout
