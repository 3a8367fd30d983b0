"""Ad-hoc timing of the device parse phase on C2/C3-shaped inputs (development aid, not a bench line)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import gen
import grlbwt_b200 as G


def run(name, arr, prof=False):
    with G.GrlGpu(0) as ctx:
        if prof:
            ctx.profile_enable(True)
        t0 = time.time()
        ctx.set_text(arr)
        st = ctx.stats()
        t1 = time.time()
        tot_ms, tot_b = 0.0, 0
        while True:
            r = ctx.round()
            tot_ms += r.device_ms
            tot_b += r.algorithmic_bytes
            print(f"  {name} r{r.round}: n={r.n_in} p={r.parse_len} d={r.n_phrases} nE={r.dict_syms} tot={r.tot_phrases} "
                  f"ms={r.device_ms:.2f} (text {r.text_pass_ms:.2f} dict {r.dict_ms:.2f} rw {r.rewrite_ms:.2f}) "
                  f"B_r={r.algorithmic_bytes / 1e6:.1f}MB -> {r.algorithmic_bytes / r.device_ms / 1e6:.1f} GB/s", flush=True)
            if prof:
                pr = ctx.profile()
                for k, (nl, ms, by) in sorted(pr.items(), key=lambda kv: -kv[1][1])[:int(__import__('os').environ.get('QT_TOP', '12'))]:
                    print(f"      {k:22s} x{nl:4d} {ms:9.3f} ms  model {by / 1e6:10.1f} MB  {by / max(ms, 1e-9) / 1e6:8.1f} GB/s")
                ctx.profile_reset()
            if r.done:
                break
        print(f"{name}: {arr.nbytes / 1e6:.1f} MB, set_text+stats {1e3 * (t1 - t0):.1f} ms, rounds {tot_ms:.1f} ms "
              f"=> {arr.nbytes / tot_ms / 1e3:.1f} MB/s parse-phase, {tot_b / tot_ms / 1e6:.1f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    reads = gen.dna_reads(int(2000000 * scale), 150, seed=42)
    run("reads", reads)
    run("reads", reads, prof=True)
    if which == "reads":
        sys.exit(0)
    run("repetitive", gen.repetitive_genomes(int(100 * scale), 1000000, seed=7), prof=True)
    run("u16", gen.int_alphabet(int(50000000 * scale), np.uint16, 65535, 1000, seed=11), prof=True)
