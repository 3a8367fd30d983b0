"""Ad-hoc timing of the host induction phase on oracle-produced levels of a C2-shaped sample (development aid)."""
import os
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gen
import grlbwt_b200 as G
from oracle import oracle as O

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cache = f"gpurun_out/levels_{n_reads}.pkl"
if os.path.exists(cache):
    levels, fp = pickle.load(open(cache, "rb"))
else:
    arr = gen.dna_reads(n_reads, 150, seed=42)
    t0 = time.time()
    o = O.Oracle(arr)
    n_rounds = o.par_phase()
    print(f"oracle par_phase: {time.time() - t0:.1f} s, {n_rounds} rounds", flush=True)
    levels = []
    for lv in range(n_rounds):
        levels.append({"alphabet": o.scalar(lv, O.ALPHABET), "tot": o.scalar(lv, O.TOT_PHRASES),
                       "rule_l": o.array(lv, O.A_RULE_L), "rule_r": o.array(lv, O.A_RULE_R),
                       "has_hocc": o.array(lv, O.A_HAS_HOCC).astype(np.uint8),
                       "pre_sym": o.array(lv, O.A_PRE_SYM), "pre_len": o.array(lv, O.A_PRE_LEN)})
    fp = o.array(n_rounds - 1, O.A_PARSE)
    o.close()
    os.makedirs("gpurun_out", exist_ok=True)
    pickle.dump((levels, fp), open(cache, "wb"))
for L in levels:
    print({k: (v.size if hasattr(v, "size") else v) for k, v in L.items()})
for rep in range(3):
    t0 = time.time()
    syms, lens = G.selftest_induce(levels, fp, n_threads=threads)
    print(f"induce: {time.time() - t0:.3f} s, runs {lens.size}, n {int(lens.sum())}", flush=True)
