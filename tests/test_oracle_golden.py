"""CPU suite: the oracle (oracle/oracle.c) against the reference's own outputs (tests/golden/golden.json,
produced by the unmodified reference via tests/golden/make_golden.py) and against the independent
numpy BCR-BWT definition."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O

CELL = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
BIG = ("reads_100k", "rep_50x200k", "u16_2M", "mixed_reads")


def check_case(arr, g):
    assert hashlib.sha256(arr.tobytes()).hexdigest() == g["input_sha256"], "generator drifted from golden input"
    o = O.Oracle(arr)
    st = g["stats"]
    assert o.stat(O.S_N_SYMS) == st["n_syms"] and o.stat(O.S_N_STRINGS) == st["n_strings"]
    assert o.stat(O.S_MIN) == st["min"] and o.stat(O.S_MAX) == st["max"]
    assert o.stat(O.S_MAX_SYM_FREQ) == st["max_sym_freq"] and o.stat(O.S_LONGEST) == st["longest"]
    R = o.par_phase()
    assert R == len(g["rounds"])
    for lv, gr in enumerate(g["rounds"]):
        assert o.scalar(lv, O.D) == gr["lms_phrases"]
        assert o.scalar(lv, O.TOT_PHRASES) == gr["tot_phrases"]
        assert o.scalar(lv, O.CELL_BYTES) == gr["cell_bytes"]
        assert o.scalar(lv, O.PARSE_LEN) == gr["parse_cells"]
        assert o.scalar(lv, O.LONGEST) == gr["longest"]
        parse = o.array(lv, O.A_PARSE).astype(CELL[gr["cell_bytes"]])
        assert hashlib.sha256(parse.tobytes()).hexdigest() == gr["parse_sha256"], f"parse of round {lv + 1}"
        assert hashlib.sha256(o.array(lv, O.A_STR_PTRS).tobytes()).hexdigest() == gr["str_ptrs_sha256"]
    syms, lens, sb, fb = o.ind_phase()
    assert (sb, fb) == (g["sb"], g["fb"])
    raw = O.rl_bwt_bytes(syms, lens, sb, fb)
    assert len(raw) == g["rl_bwt_bytes"]
    assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"]
    o.close()
    return syms, lens


def test_oracle_matches_reference_fixtures(golden, all_cases):
    for name in ("test_byte_alphabet", "test_2bytes_alphabet"):
        check_case(all_cases[name], golden[name])


def test_oracle_matches_reference_small_and_fuzz(golden, all_cases):
    n = 0
    for name, arr in all_cases.items():
        if name in BIG or name.startswith("test_"):
            continue
        syms, lens = check_case(arr, golden[name])
        # independent definition
        bs, bl = O.rle(O.bcr_bwt(arr))
        assert np.array_equal(bs, syms) and np.array_equal(bl, lens), name
        n += 1
    assert n >= 130


@pytest.mark.parametrize("name", BIG)
def test_oracle_matches_reference_config_shapes(golden, all_cases, name):
    check_case(all_cases[name], golden[name])


def test_rl_bwt_reader_roundtrip():
    syms = np.array([10, 65, 300, 65], np.uint64)
    lens = np.array([3, 70000, 1, 2], np.uint64)
    raw = O.rl_bwt_bytes(syms, lens, 2, 3)
    s2, l2, sb, fb = O.read_rl_bwt(raw)
    assert (sb, fb) == (2, 3) and np.array_equal(s2, syms) and np.array_equal(l2, lens)


def test_ill_formed_collection_rejected():
    with pytest.raises(ValueError):
        O.Oracle(np.frombuffer(b"ACGT\nAC", np.uint8))       # does not end with the separator
    with pytest.raises(ValueError):
        O.Oracle(np.frombuffer(b"AC\x01GT\nAC\n", np.uint8))  # separator is not the smallest symbol
