import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="session")
def all_cases():
    """name -> numpy array for every case that golden.json pins."""
    import gen
    import numpy as np
    cases = {"test_byte_alphabet": gen.fixture_byte_alphabet(), "test_2bytes_alphabet": gen.fixture_2bytes_alphabet()}
    cases.update(gen.small_cases())
    for name, arr in gen.fuzz_cases():
        cases[name] = arr
    cases["reads_100k"] = gen.dna_reads(100000, 150, seed=42)
    cases["rep_50x200k"] = gen.repetitive_genomes(50, 200000, seed=7)
    cases["u16_2M"] = gen.int_alphabet(2000000, np.uint16, 65535, 1000, seed=11)
    cases["mixed_reads"] = gen.mixed_reads(20000, 30, 150, 10000, seed=5)
    return cases
