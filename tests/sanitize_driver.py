"""Driver for `make sanitize`: a handful of small collections through every code path of the device library
(default, table regrow + long-run scan, thread-per-phrase dedup, one-tile pilot, prefix doubling), each checked
against the oracle, sized for compute-sanitizer (memcheck / racecheck slow kernels down 10-100x).
Run as:  GRLGPU_NO_POOL=1 compute-sanitizer --tool memcheck python tests/sanitize_driver.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen  # noqa: E402
import grlbwt_b200 as G  # noqa: E402
from oracle import oracle as O  # noqa: E402


def run_case(name, arr, flags):
    o = O.Oracle(arr)
    R = o.par_phase()
    with G.GrlGpu(0, flags) as ctx:
        ctx.set_text(arr)
        for lv in range(R):
            r = ctx.round()
            assert r.tot_phrases == o.scalar(lv, O.TOT_PHRASES), (name, flags, lv)
            L = ctx.fetch_level()
            assert np.array_equal(L["rule_r"], o.array(lv, O.A_RULE_R)), (name, flags, lv)
            assert np.array_equal(ctx.fetch_parse().astype(np.uint64), o.array(lv, O.A_PARSE)), (name, flags, lv)
    o.close()


def main():
    cases = gen.small_cases()
    cases["reads_300x150"] = gen.dna_reads(300, 150, seed=42)
    cases["empty_runs"] = np.frombuffer(b"\n" * 20000 + b"ACGT\n" * 300 + b"\n" * 100, np.uint8).copy()
    names = ["mississippi", "with_empty", "only_empty", "homopolymers_multi", "high_bytes", "dna_500", "mutated_200x5k", "u16_small_sigma",
             "u32_small_sigma", "u64_rand", "long_phrases", "reads_300x150", "empty_runs"]
    variants = [0, G.FLAG_SMALL_TABLE | G.FLAG_FORCE_SLOW_SCAN, G.FLAG_FORCE_UNCACHED, G.FLAG_SMALL_PILOT, G.FLAG_FORCE_DOUBLING]
    n = 0
    if "--quick" in sys.argv:  # only what the full runs in profiles/ predate: multi-rank rounds (records in source order, rule records) + the packed output
        from grlbwt_b200 import mg
        for name in ("dna_500", "u16_small_sigma"):
            mg.check_against_oracle(cases[name], n_ranks=3)
            n += 1
        for name in ("dna_500", "u16_small_sigma", "reads_300x150", "with_empty"):
            arr = cases[name]
            o = O.Oracle(arr)
            o.par_phase()
            osyms, olens, osb, ofb = o.ind_phase()
            img = np.zeros(16 + arr.size * 16, np.uint8)
            for devs in ([0], [0, 0]):
                nb, n_runs, sb, fb, info = G.build_bwt_packed(arr, img, devices=devs, n_threads=2)
                assert (sb, fb) == (osb, ofb) and n_runs == osyms.size, (name, devs, info)
                assert img[:nb].tobytes() == O.rl_bwt_bytes(osyms, olens, osb, ofb), (name, devs)
                n += 1
            o.close()
        print(f"sanitize driver ok (quick): {n} runs")
        return
    for name in names:
        for fl in variants:
            run_case(name, cases[name], fl)
            n += 1
    if "--mg" in sys.argv:  # multi-rank rounds with in-process ranks on one device
        from grlbwt_b200 import mg
        for name in ("dna_500", "mutated_200x5k", "u16_small_sigma"):
            mg.check_against_oracle(cases[name], n_ranks=3)
            n += 1
    print(f"sanitize driver ok: {n} runs")


if __name__ == "__main__":
    main()
