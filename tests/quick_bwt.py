"""Ad-hoc timing of the whole construction (device parse phase + host induction) -- development aid."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gen
import grlbwt_b200 as G

for name, arr in (("reads_2M", gen.dna_reads(2000000, 150, seed=42)), ("rep_100x1M", gen.repetitive_genomes(100, 1000000, seed=7)),
                  ("u16_50M", gen.int_alphabet(50000000, np.uint16, 65535, 1000, seed=11))):
    for rep in range(2):
        t0 = time.time()
        syms, lens, sb, fb, info = G.build_bwt(arr, n_threads=16)
        dt = time.time() - t0
    print(f"{name}: {arr.nbytes/1e6:.1f} MB total {dt:.2f}s -> {arr.nbytes/1e6/dt:.1f} MB/s | h2d {info['h2d_ms']:.0f} ms, parse {info['par_phase_ms']:.0f} ms "
          f"(device {info['device_ms']:.0f}), induction {info['ind_phase_ms']:.0f} ms, runs {len(syms)}", flush=True)
