"""Seeded synthetic collections (numpy default_rng) shared by tests, golden generation and bench.py.

Every collection ends with the separator and the separator is the strictly smallest symbol, as the
reference requires (external/cdt/lib/utils.cpp:177-180).  Shapes follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import gzip
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DNA = np.frombuffer(b"ACGT", np.uint8)


def fixture_byte_alphabet() -> np.ndarray:
    """The reference's test_data/test_byte_alphabet.txt (2 956 004 B, 200 DNA strings)."""
    with gzip.open(os.path.join(HERE, "golden", "test_byte_alphabet.txt.gz"), "rb") as f:
        return np.frombuffer(f.read(), np.uint8).copy()


def fixture_2bytes_alphabet() -> np.ndarray:
    """The reference's test_data/test_2bytes_alphabet.txt (1000 x uint16, 10 strings, separator 0)."""
    return np.fromfile(os.path.join(HERE, "golden", "test_2bytes_alphabet.bin"), np.uint16)


def dna_reads(n_reads: int, read_len: int = 150, seed: int = 42) -> np.ndarray:
    """C2 shape: n_reads x read_len uniform ACGT, '\\n' after each read."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_reads, read_len + 1), np.uint8)
    chunk = 1 << 20
    for i in range(0, n_reads, chunk):
        j = min(n_reads, i + chunk)
        out[i:j, :read_len] = DNA[rng.integers(0, 4, size=(j - i, read_len), dtype=np.uint8)]
    out[:, read_len] = 10
    return out.reshape(-1)


def repetitive_genomes(n_copies: int, genome_len: int, seed: int = 7, snp: float = 1e-3, dele: float = 1e-4) -> np.ndarray:
    """C3 shape: copies of one random genome with substitutions and single-base deletions."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=genome_len, dtype=np.uint8)
    parts = []
    for _ in range(n_copies):
        s = base.copy()
        idx = np.flatnonzero(rng.random(genome_len) < snp)
        s[idx] = (s[idx] + rng.integers(1, 4, size=idx.size, dtype=np.uint8)) % 4
        keep = rng.random(genome_len) >= dele
        parts.append(DNA[s[keep]])
        parts.append(np.array([10], np.uint8))
    return np.concatenate(parts)


def int_alphabet(n: int, dtype=np.uint16, max_sym: int = 65535, every: int = 1000, seed: int = 11) -> np.ndarray:
    """C4 shape: uniform symbols in [1,max_sym], separator 0 at every `every`-th cell and at the end."""
    rng = np.random.default_rng(seed)
    t = rng.integers(1, max_sym + 1, size=n, dtype=np.uint64).astype(dtype)
    t[every - 1::every] = 0
    t[-1] = 0
    return t


def mixed_reads(n_short: int, n_long: int, short_len: int = 150, long_len: int = 10000, seed: int = 5) -> np.ndarray:
    """C5 shape: short reads with long reads interleaved deterministically."""
    rng = np.random.default_rng(seed)
    parts = []
    step = max(1, n_short // max(1, n_long))
    li = 0
    for i in range(n_short):
        parts.append(DNA[rng.integers(0, 4, size=short_len, dtype=np.uint8)])
        parts.append(np.array([10], np.uint8))
        if n_long and i % step == step - 1 and li < n_long:
            parts.append(DNA[rng.integers(0, 4, size=long_len, dtype=np.uint8)])
            parts.append(np.array([10], np.uint8))
            li += 1
    return np.concatenate(parts)


def random_collection(rng: np.random.Generator, sigma: int, n_strings: int, max_len: int, dtype=np.uint8,
                      base: int = 65, sep: int = 10, allow_empty: bool = True) -> np.ndarray:
    """Fuzz shape: short strings over a tiny alphabet (runs, duplicates, empty strings)."""
    parts = []
    for _ in range(n_strings):
        ln = int(rng.integers(0 if allow_empty else 1, max_len + 1))
        parts.append((rng.integers(0, sigma, size=ln) + base).astype(dtype))
        parts.append(np.array([sep], dtype))
    return np.concatenate(parts)


def small_cases():
    """name -> array.  Hand-made corner cases + seeded small collections (all oracle-sized)."""
    cases = {}
    cases["mississippi"] = np.frombuffer(b"mississippi\n", np.uint8).copy()
    cases["single_sep"] = np.frombuffer(b"\n", np.uint8).copy()
    cases["only_empty"] = np.frombuffer(b"\n\n\n\n", np.uint8).copy()
    cases["with_empty"] = np.frombuffer(b"ACGT\n\nAC\n\n\nTTTT\nACGT\n", np.uint8).copy()
    cases["homopolymer"] = np.frombuffer(b"A" * 29 + b"\n", np.uint8).copy()
    cases["homopolymers_multi"] = np.frombuffer(b"A" * 1000 + b"\n" + b"A" * 999 + b"\n" + b"C" * 70 + b"A" * 300 + b"\n", np.uint8).copy()
    cases["monotone"] = np.frombuffer(bytes(range(11, 255)) + b"\n" + bytes(range(254, 10, -1)) + b"\n", np.uint8).copy()
    cases["high_bytes"] = np.frombuffer(bytes([255, 254, 253, 252, 255, 11, 255, 254, 10, 252, 252, 255, 10]), np.uint8).copy()
    rng = np.random.default_rng(1)
    cases["ac_short_3000"] = random_collection(rng, 2, 3000, 12, allow_empty=False)
    rng = np.random.default_rng(2)
    parts = []
    for _ in range(500):
        parts.append(DNA[rng.integers(0, 4, size=int(rng.integers(30, 201)), dtype=np.uint8)])
        parts.append(np.array([10], np.uint8))
    cases["dna_500"] = np.concatenate(parts)
    cases["mutated_200x5k"] = repetitive_genomes(200, 5000, seed=3, snp=2e-3, dele=5e-4)
    cases["reads_2000x150"] = dna_reads(2000, 150, seed=42)
    cases["u16_rand"] = int_alphabet(50000, np.uint16, 65535, 1000, seed=11)
    cases["u16_small_sigma"] = random_collection(np.random.default_rng(4), 5, 400, 60, np.uint16, base=1000, sep=3)
    cases["u32_rand"] = int_alphabet(30000, np.uint32, (1 << 31) - 5, 500, seed=12)
    cases["u32_small_sigma"] = random_collection(np.random.default_rng(5), 3, 300, 80, np.uint32, base=70000, sep=0)
    cases["u64_rand"] = int_alphabet(20000, np.uint64, (1 << 33), 700, seed=13)
    cases["long_phrases"] = np.concatenate([np.frombuffer(b"AC" * 2000 + b"\n", np.uint8),
                                            np.frombuffer(b"ACGT" * 700 + b"A" * 3000 + b"\n", np.uint8),
                                            np.frombuffer(b"AC" * 1999 + b"\n", np.uint8)]).copy()
    return cases


def fuzz_cases(n_cases: int = 120, seed: int = 1234):
    rng = np.random.default_rng(seed)
    for i in range(n_cases):
        sigma = int(rng.integers(1, 5))
        n_strings = int(rng.integers(1, 40))
        max_len = int(rng.integers(1, 60))
        yield f"fuzz_{i}", random_collection(rng, sigma, n_strings, max_len)
