"""GPU suite (-m gpu): the CUDA path, called through the C ABI (include/grlgpu.h, include/grlbwt.h),
against the oracle on the same seeded inputs and against the reference's golden outputs. Bit-exact."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import grlbwt_b200 as G
from oracle import oracle as O

pytestmark = pytest.mark.gpu
CELL = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}
BIG = ("reads_100k", "rep_50x200k", "u16_2M", "mixed_reads")


# ------------------------------------------------------------------ primitives
@pytest.mark.parametrize("n", [1, 2, 31, 2048, 2049, 8192, 8193, 100003, (1 << 21) + 5])
def test_scan(n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 1000, size=n, dtype=np.uint32)
    out, tot = G.selftest_scan(a)
    ref = np.concatenate(([0], np.cumsum(a.astype(np.uint64))[:-1]))
    assert tot == int(a.astype(np.uint64).sum())
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("n,bits", [(2, 8), (100, 16), (4096, 64), (4097, 24), (100000, 40), ((1 << 20) + 17, 64), (300000, 3), (5_000_003, 33)])
def test_radix_sort(n, bits):
    rng = np.random.default_rng(n + bits)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    if bits < 64:
        keys &= np.uint64((1 << bits) - 1)
    if n > 1000:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)]  # duplicates -> stability matters
    vals = np.arange(n, dtype=np.uint32)
    k, v = G.selftest_sort(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


@pytest.mark.parametrize("n_bits", [1, 31, 32, 33, 65536, 65537, 1000003])
def test_bitmap_compact(n_bits):
    rng = np.random.default_rng(n_bits)
    bits = rng.random(n_bits) < 0.3
    prev = rng.random(n_bits) < 0.1
    def pack(b):
        pad = np.zeros((-len(b)) % 32, bool)
        return np.packbits(np.concatenate((b, pad)).reshape(-1, 32)[:, ::-1], axis=1).view(">u4").astype(np.uint32).reshape(-1)
    got = G.selftest_compact(pack(bits), None, n_bits)
    assert np.array_equal(got, np.flatnonzero(bits).astype(np.uint64))
    got = G.selftest_compact(pack(bits), pack(prev), n_bits)
    pos = np.flatnonzero(bits)
    flag = (pos == 0) | prev[np.maximum(pos, 1) - 1]
    assert np.array_equal(got, pos.astype(np.uint64) | (flag.astype(np.uint64) << np.uint64(63)))


# ------------------------------------------------------------------ rounds against the oracle
def compare_rounds(name, arr, g, flags=0, check_dict=True):
    o = O.Oracle(arr)
    R = o.par_phase()
    with G.GrlGpu(0, flags | (G.FLAG_KEEP_DICT if check_dict else 0)) as ctx:
        ctx.set_text(arr)
        st = ctx.stats()
        for k, w in (("n_syms", O.S_N_SYMS), ("n_strings", O.S_N_STRINGS), ("longest_string", O.S_LONGEST), ("min_sym", O.S_MIN),
                     ("max_sym", O.S_MAX), ("max_sym_freq", O.S_MAX_SYM_FREQ), ("sep_sym", O.S_SEP)):
            assert getattr(st, k) == o.stat(w), f"{name}: stats.{k}"
        for lv in range(R):
            r = ctx.round()
            tag = f"{name} round {lv + 1}"
            assert r.n_in == o.scalar(lv, O.N_IN), tag
            assert r.alphabet == o.scalar(lv, O.ALPHABET), tag
            assert r.parse_len == o.scalar(lv, O.PARSE_LEN), f"{tag}: parse_len {r.parse_len} vs {o.scalar(lv, O.PARSE_LEN)}"
            assert r.n_phrases == o.scalar(lv, O.D), f"{tag}: distinct phrases {r.n_phrases} vs {o.scalar(lv, O.D)}"
            assert r.dict_syms == o.scalar(lv, O.SUM_LEN), tag
            assert r.max_freq == o.scalar(lv, O.MAX_FREQ), tag
            if check_dict:
                syms, lens, freqs, metas = ctx.fetch_dictionary()
                assert np.array_equal(lens, o.array(lv, O.A_DICT_LEN)), f"{tag}: dictionary order/lengths"
                assert np.array_equal(syms, o.array(lv, O.A_DICT_SYMS)), f"{tag}: dictionary symbols"
                assert np.array_equal(freqs, o.array(lv, O.A_DICT_FREQ)), f"{tag}: phrase frequencies"
                assert np.array_equal(metas, o.array(lv, O.A_DICT_META)), f"{tag}: metasymbols"
            assert r.tot_phrases == o.scalar(lv, O.TOT_PHRASES), f"{tag}: tot_phrases {r.tot_phrases} vs {o.scalar(lv, O.TOT_PHRASES)}"
            assert r.cell_bytes_out == o.scalar(lv, O.CELL_BYTES), tag
            assert r.done == (1 if lv == R - 1 else 0), tag
            L = ctx.fetch_level()
            assert np.array_equal(L["pre_sym"], o.array(lv, O.A_PRE_SYM)), f"{tag}: pre-BWT symbols"
            assert np.array_equal(L["pre_len"], o.array(lv, O.A_PRE_LEN)), f"{tag}: pre-BWT lengths"
            assert np.array_equal(L["has_hocc"], o.array(lv, O.A_HAS_HOCC).astype(np.uint8)), f"{tag}: has_hocc"
            assert np.array_equal(L["rule_l"], o.array(lv, O.A_RULE_L)), f"{tag}: rule_l"
            assert np.array_equal(L["rule_r"], o.array(lv, O.A_RULE_R)), f"{tag}: rule_r"
            parse = ctx.fetch_parse()
            assert np.array_equal(parse.astype(np.uint64), o.array(lv, O.A_PARSE)), f"{tag}: parse"
            assert np.array_equal(ctx.fetch_str_ptrs(), o.array(lv, O.A_STR_PTRS)), f"{tag}: str_ptrs"
            if g is not None:
                gr = g["rounds"][lv]
                assert hashlib.sha256(parse.tobytes()).hexdigest() == gr["parse_sha256"], f"{tag}: parse vs reference dump"
                assert r.n_phrases == gr["lms_phrases"] and r.tot_phrases == gr["tot_phrases"], tag
        with pytest.raises(G.GrlGpuError):
            ctx.round()  # phase finished: GRLGPU_ERR_STATE
    o.close()


def test_rounds_reference_fixtures(golden, all_cases):
    for name in ("test_2bytes_alphabet", "test_byte_alphabet"):
        compare_rounds(name, all_cases[name], golden[name])


def test_rounds_corner_cases(golden, all_cases):
    for name, arr in all_cases.items():
        if name in BIG or name.startswith("test_") or name.startswith("fuzz_"):
            continue
        compare_rounds(name, arr, golden[name])


def test_rounds_fuzz(golden, all_cases):
    for name, arr in all_cases.items():
        if name.startswith("fuzz_"):
            compare_rounds(name, arr, golden[name])


@pytest.mark.parametrize("name", BIG)
def test_rounds_config_shapes(golden, all_cases, name):
    compare_rounds(name, all_cases[name], golden[name], check_dict=False)


def test_rounds_small_table_and_slow_scan(golden, all_cases):
    """regrow path of the phrase table and the long-run path of the boundary scan give the same bytes"""
    for name in ("mutated_200x5k", "homopolymers_multi", "long_phrases", "u16_rand", "dna_500", "fuzz_3", "fuzz_50"):
        compare_rounds(name, all_cases[name], golden[name], flags=G.FLAG_SMALL_TABLE | G.FLAG_FORCE_SLOW_SCAN)


def test_rounds_dedup_variants(golden, all_cases):
    """thread-per-phrase dedup for every phrase, and one-tile pilot + remainder (cached or not, chosen by the data)"""
    for name in ("mutated_200x5k", "u16_rand", "with_empty", "only_empty", "long_phrases", "reads_2000x150", "u64_rand", "fuzz_7"):
        compare_rounds(name, all_cases[name], golden[name], flags=G.FLAG_FORCE_UNCACHED)
    for name in ("mutated_200x5k", "u16_rand", "reads_2000x150", "u32_rand", "homopolymers_multi", "ac_short_3000"):
        compare_rounds(name, all_cases[name], golden[name], flags=G.FLAG_SMALL_PILOT)
    for name in ("rep_50x200k", "u16_2M"):
        compare_rounds(name, all_cases[name], golden[name], flags=G.FLAG_SMALL_PILOT | G.FLAG_SMALL_TABLE, check_dict=False)


def test_rounds_refinement_by_doubling(golden, all_cases):
    """suffix groups refined by prefix doubling on position-based ranks (the path long phrases take) on ordinary inputs too"""
    for name in ("test_byte_alphabet", "mutated_200x5k", "u16_rand", "ac_short_3000", "reads_2000x150", "with_empty", "fuzz_9", "fuzz_33"):
        compare_rounds(name, all_cases[name], golden[name], flags=G.FLAG_FORCE_DOUBLING)
    compare_rounds("rep_50x200k", all_cases["rep_50x200k"], golden["rep_50x200k"], flags=G.FLAG_FORCE_DOUBLING, check_dict=False)


def test_many_empty_strings_overflow_the_tile_list():
    """more phrase starts in one tile than the shared-memory list holds (runs of empty strings)"""
    arr = np.frombuffer(b"\n" * 70000 + b"ACGT\n" * 3000 + b"\n" * 40000, np.uint8).copy()
    compare_rounds("empty_runs", arr, None)


def test_long_equal_runs_cross_cta_boundaries():
    """homopolymer stretches longer than a CTA tile (8192 cells) + look-ahead: the summary/resolve path must kick in by itself"""
    parts = [b"ACGT" * 100 + b"A" * 20000 + b"C" + b"\n", b"T" * 50000 + b"\n", b"G" * 8192 + b"\n", b"CA" * 5000 + b"A" * 9000 + b"\n"]
    arr = np.frombuffer(b"".join(parts), np.uint8).copy()
    compare_rounds("long_runs", arr, None)
    syms, lens, sb, fb, _ = G.build_bwt(arr)
    bs, bl = O.rle(O.bcr_bwt(arr))
    assert np.array_equal(bs, syms) and np.array_equal(bl, lens)


# ------------------------------------------------------------------ whole construction against the reference
def test_full_bwt_all_cases(golden, all_cases):
    for name, arr in all_cases.items():
        g = golden[name]
        syms, lens, sb, fb, info = G.build_bwt(arr)
        assert (sb, fb) == (g["sb"], g["fb"]), name
        raw = O.rl_bwt_bytes(syms, lens, sb, fb)
        assert len(raw) == g["rl_bwt_bytes"], name
        assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"], name
        assert info["n_rounds"] == len(g["rounds"]), name


def test_host_and_device_induction_agree(golden, all_cases):
    """the induction phase on the device (default below 2^32 symbols) and on the host threads (GRLBWT_HOST_INDUCTION=1, and
    whenever a level needs 64-bit symbols) produce the reference's bytes"""
    names = ("test_byte_alphabet", "test_2bytes_alphabet", "mutated_200x5k", "with_empty", "only_empty", "long_phrases", "u32_rand", "reads_100k", "rep_50x200k",
             "u16_2M", "mixed_reads", "fuzz_5", "fuzz_77")
    for name in names:
        g = golden[name]
        for host in (False, True):
            if host:
                os.environ["GRLBWT_HOST_INDUCTION"] = "1"
            try:
                syms, lens, sb, fb, info = G.build_bwt(all_cases[name], n_threads=4)
            finally:
                os.environ.pop("GRLBWT_HOST_INDUCTION", None)
            assert bool(info["induced_on_device"]) == (not host), (name, host)
            assert hashlib.sha256(O.rl_bwt_bytes(syms, lens, sb, fb)).hexdigest() == g["rl_bwt_sha256"], (name, host)
    syms, lens, sb, fb, info = G.build_bwt(all_cases["u64_rand"])   # 64-bit symbols: the host induces
    assert not info["induced_on_device"]


def test_build_into_caller_buffers(golden, all_cases):
    """grlbwt_build_to: the run-length BWT lands in caller-owned 32-bit arrays (device induction: by DMA; host induction: copied)"""
    for name in ("test_byte_alphabet", "reads_100k", "u16_2M"):
        arr, g = all_cases[name], golden[name]
        out_s, out_l = np.zeros(arr.size, np.uint32), np.zeros(arr.size, np.uint32)
        for host in (False, True):
            if host:
                os.environ["GRLBWT_HOST_INDUCTION"] = "1"
            try:
                n_runs, sb, fb, info = G.build_bwt_to(arr, out_s, out_l, n_threads=4)
            finally:
                os.environ.pop("GRLBWT_HOST_INDUCTION", None)
            raw = O.rl_bwt_bytes(out_s[:n_runs].astype(np.uint64), out_l[:n_runs].astype(np.uint64), sb, fb)
            assert hashlib.sha256(raw).hexdigest() == g["rl_bwt_sha256"], (name, host)
    with pytest.raises(G.GrlGpuError):
        G.build_bwt_to(all_cases["reads_100k"], np.zeros(10, np.uint32), np.zeros(10, np.uint32))


def test_build_packed_image(golden, all_cases):
    """grlbwt_build_packed: the image of the .rl_bwt file in a caller-owned buffer (records packed on the device; host induction:
    packed by the host), bit-identical to the reference's output file; several ranks give the same image"""
    for name in ("test_byte_alphabet", "test_2bytes_alphabet", "reads_100k", "u16_2M", "mutated_200x5k", "with_empty", "only_empty", "u64_rand", "homopolymer"):
        arr, g = all_cases[name], golden[name]
        img = np.zeros(16 + arr.size * 16, np.uint8)
        for host in (False, True):
            if host:
                os.environ["GRLBWT_HOST_INDUCTION"] = "1"
            try:
                nb, n_runs, sb, fb, info = G.build_bwt_packed(arr, img, n_threads=4)
            finally:
                os.environ.pop("GRLBWT_HOST_INDUCTION", None)
            assert nb == 16 + n_runs * (sb + fb) and not (host and info["induced_on_device"]), (name, host, info)
            assert hashlib.sha256(img[:nb].tobytes()).hexdigest() == g["rl_bwt_sha256"], (name, host)
        nb2, _, _, _, _ = G.build_bwt_packed(arr, img, devices=[0, 0, 0], n_threads=4)
        assert hashlib.sha256(img[:nb2].tobytes()).hexdigest() == g["rl_bwt_sha256"], (name, "3 ranks")
    with pytest.raises(G.GrlGpuError):
        G.build_bwt_packed(all_cases["reads_100k"], np.zeros(64, np.uint8))


def test_async_level_fetch_matches_sync(golden, all_cases):
    """levels fetched on the copy stream while later rounds run equal the synchronously fetched ones"""
    for name in ("mutated_200x5k", "u16_rand", "reads_2000x150"):
        arr = all_cases[name]
        sync_levels = []
        with G.GrlGpu(0) as ctx:
            ctx.set_text(arr)
            while True:
                r = ctx.round()
                sync_levels.append(ctx.fetch_level())
                if r.done:
                    break
        arena = np.zeros(sum(L["rule_l"].size * 17 + L["pre_sym"].size * 16 + 256 for L in sync_levels), np.uint8)
        got, off = [], 0
        with G.GrlGpu(0) as ctx:
            ctx.set_text(arr)
            while True:
                r = ctx.round()
                got.append(ctx.fetch_level(arena, async_=True, offset=off))
                off = ctx.arena_end
                with pytest.raises(G.GrlGpuError):
                    ctx.fetch_level()            # the level was handed to the copy stream
                if r.done:
                    break
            ctx.fetch_wait()
        for a, b in zip(sync_levels, got):
            for k in ("rule_l", "rule_r", "has_hocc", "pre_sym", "pre_len"):
                assert np.array_equal(a[k], b[k].astype(a[k].dtype)), (name, k)
        # 32-bit run lengths (grlgpu_fetch_level32), synchronous and on the copy stream
        for async_ in (False, True):
            got32, off = [], 0
            with G.GrlGpu(0) as ctx:
                ctx.set_text(arr)
                while True:
                    r = ctx.round()
                    got32.append(ctx.fetch_level(arena, async_=async_, offset=off, narrow_len=True))
                    assert got32[-1]["pre_len"].dtype == np.uint32
                    off = ctx.arena_end
                    if r.done:
                        break
                ctx.fetch_wait()
            for a, b in zip(sync_levels, got32):
                for k in ("rule_l", "rule_r", "has_hocc", "pre_sym", "pre_len"):
                    assert np.array_equal(a[k], b[k].astype(a[k].dtype)), (name, k, async_)


def test_ill_formed_rejected():
    for bad in (b"ACGT\nAC", b"AC\x01GT\nAC\n"):
        with G.GrlGpu(0) as ctx:
            ctx.set_text(np.frombuffer(bad, np.uint8))
            with pytest.raises(G.GrlGpuError) as e:
                ctx.stats()
            assert e.value.status == -2
    with G.GrlGpu(0) as ctx:
        with pytest.raises(G.GrlGpuError):
            ctx.set_text(np.zeros(0, np.uint8))
        with pytest.raises(G.GrlGpuError):
            ctx.round()


def test_cli_end_to_end(tmp_path, golden, all_cases):
    exe = os.path.join(G.LIB_DIR, "grlbwt")
    for name, a in (("test_byte_alphabet", 1), ("test_2bytes_alphabet", 2)):
        inp = tmp_path / (name + ".txt")
        all_cases[name].tofile(inp)
        r = subprocess.run([exe, str(inp), "-a", str(a), "-t", "4", "-T", str(tmp_path)], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        out = tmp_path / (name + ".rl_bwt")           # extension replaced, relative to cwd (main.cpp:112-113)
        assert hashlib.sha256(out.read_bytes()).hexdigest() == golden[name]["rl_bwt_sha256"]
        assert "The resulting BCR BWT was stored in" in r.stdout
    bad = tmp_path / "bad.txt"
    bad.write_bytes(b"ACGT\nAC")
    r = subprocess.run([exe, str(bad)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 1 and "ill formed" in r.stdout


# ------------------------------------------------------------------ sizes the oracle does not reach: BWT invariants
@pytest.mark.parametrize("name,a,threads", [("reads_2M", 1, 16), ("rep_100x1M", 1, 4), ("u16_20M", 2, 8), ("mixed_200k", 1, 8)])
def test_cli_at_scale_satisfies_bwt_invariants(tmp_path, name, a, threads):
    """CLI end to end on config-shaped inputs of 40-300 MB, then bwt_check: header, sum of lengths, maximal runs,
    per-symbol totals, the separator block, and LF-inversion of 1000 strings against the text"""
    import gen
    arr = {"reads_2M": lambda: gen.dna_reads(2000000, 150, seed=42), "rep_100x1M": lambda: gen.repetitive_genomes(100, 1000000, seed=7),
           "u16_20M": lambda: gen.int_alphabet(20000000, np.uint16, 65535, 1000, seed=11),
           "mixed_200k": lambda: gen.mixed_reads(200000, 300, 150, 10000, seed=5)}[name]()
    inp = tmp_path / (name + ".txt")
    arr.tofile(inp)
    exe = os.path.join(G.LIB_DIR, "grlbwt")
    r = subprocess.run([exe, str(inp), "-a", str(a), "-t", str(threads), "-T", str(tmp_path)], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout + r.stderr)[-600:]
    r = subprocess.run([os.path.join(G.LIB_DIR, "bwt_check"), str(inp), str(tmp_path / (name + ".rl_bwt")), "-a", str(a), "-k", "1000"],
                       capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout
