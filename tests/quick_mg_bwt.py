"""Ad-hoc timing of the whole construction on 1..N GPUs of one process (local P2P exchanges and NCCL) -- development aid.
usage: python tests/quick_mg_bwt.py [reads] [n_gpus]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import gen
import grlbwt_b200 as G
from grlbwt_b200 import mg

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
ngpu = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
thr = os.cpu_count() or 1
sample = torch.from_numpy(gen.dna_reads(reads, 150, seed=42)).pin_memory().numpy()
out_s = torch.empty(sample.size, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
out_l = torch.empty(sample.size, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
ref = None
n = 1
while n <= ngpu:
    for comm, name in ((mg.COMM_LOCAL, "local"), (mg.COMM_NCCL, "nccl")):
        if n == 1 and comm != mg.COMM_LOCAL:
            continue
        devs = list(range(n))
        G.build_bwt_to(sample[: 151 * 1000], out_s, out_l, devices=devs, n_threads=thr, comm=comm)
        for rep in range(2):
            t0 = time.perf_counter()
            n_runs, sb, fb, info = G.build_bwt_to(sample, out_s, out_l, devices=devs, n_threads=thr, comm=comm)
            wall = (time.perf_counter() - t0) * 1e3
            print(f"gpus {n} comm {name} rep {rep}: {sample.nbytes / 1e6 / (wall / 1e3):.1f} MB/s wall {wall:.0f} ms | h2d {info['h2d_ms']:.0f} parse {info['par_phase_ms']:.0f} "
                  f"(device {info['device_ms']:.0f}) induction {info['ind_phase_ms']:.0f} on_device {info['induced_on_device']} runs {n_runs}", flush=True)
        sig = (n_runs, int(out_s[:n_runs].astype(np.uint64).sum()), int(out_l[:n_runs].astype(np.uint64).sum()), hash(out_l[:n_runs].tobytes()))
        if ref is None:
            ref = sig
        print("   identical to the 1-GPU BWT:", sig == ref, flush=True)
    n *= 2
