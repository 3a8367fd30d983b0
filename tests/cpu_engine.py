"""TEST DOUBLE (tests only): a pure-Python engine with the interface of grlbwt_b200.multigpu.GpuEngine, so the
multi-rank orchestration (sharding, hash-partitioned all-to-all-v, owner merge, all-gather-v, termination,
final gather) runs under gloo on CPU tensors. It restates the round semantics of SURVEY.md App. A on small
inputs with Python containers; it is not a product path and is never imported by grlbwt_b200/."""
from __future__ import annotations

import functools
import zlib
from collections import Counter

import numpy as np
import torch

NP_CELL = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def sym_width(v):
    return int(v).bit_length()


def cmp_pg(a, b):  # A.2: proper prefix is GREATER
    for x, y in zip(a, b):
        if x != y:
            return -1 if x < y else 1
    return 0 if len(a) == len(b) else (1 if len(a) < len(b) else -1)


class CpuEngine:
    def __init__(self, shard: np.ndarray):
        self.device = torch.device("cpu")
        self.w = shard.dtype.itemsize
        self.cells = [int(x) for x in shard]      # raw cells
        self.first = True
        self.sep = self.cells[-1]
        ends = [i for i, c in enumerate(self.cells) if c == self.sep]
        self.str_ptrs = [0] + [e + 1 for e in ends]
        self.A = None
        self.is_suffix = None
        self._level = None
        self._local = None

    # ---- helpers ----
    def _val(self, c):
        return c if self.first else c >> 1

    def _rep(self, c):
        return 1 if self.first else c & 1

    def alloc(self, n, dtype):
        return torch.zeros(max(int(n), 1), dtype=dtype)

    def stats(self):
        v = self.cells
        n_str = len(self.str_ptrs) - 1
        longest = max(self.str_ptrs[i + 1] - self.str_ptrs[i] for i in range(n_str))
        return {"n_syms": len(v), "n_strings": n_str, "longest_string": longest, "min_sym": min(v), "max_sym": max(v),
                "max_sym_freq": max(Counter(v).values()) if self.w == 1 else len(v), "sep_sym": self.sep}

    def histogram(self):
        return np.bincount(np.array(self.cells, np.int64), minlength=256).astype(np.uint64)

    def set_alphabet(self, max_sym):
        self.A = max_sym + 1
        self.is_suffix = {self.sep: True}

    def cell_bytes(self):
        return self.w

    def _suffix(self, sym):
        return bool(self.is_suffix.get(sym, False))

    # ---- A.1 on the shard ----
    def _parse(self):
        cells, out = self.cells, []
        for s in range(len(self.str_ptrs) - 1):
            st, en = self.str_ptrs[s], self.str_ptrs[s + 1] - 1
            t_next, brk = 0, []
            for i in range(en - 1, st - 1, -1):
                a, b = self._val(cells[i]), self._val(cells[i + 1])
                t = (1 if a < b else 0) if a != b else t_next
                if a != b and t_next == 1 and t == 0 and self._rep(cells[i]) and self._rep(cells[i + 1]):
                    brk.append(i + 1)
                t_next = t
            b = [st] + sorted(brk)
            phr = [tuple(cells[b[k]:b[k + 1] + 1]) for k in range(len(b) - 1)] + [tuple(cells[b[-1]:en + 1])]
            out.append(phr)
        return out

    def local(self, n_ranks):
        per_str = self._parse()
        cnt = Counter(p for ph in per_str for p in ph)
        owner = {p: zlib.crc32(np.array(p, NP_CELL[self.w]).tobytes()) % n_ranks for p in cnt}
        order = sorted(cnt, key=lambda p: (owner[p], p))
        self._local = (per_str, cnt, order)
        per = []
        for g in range(n_ranks):
            ps = [p for p in order if owner[p] == g]
            per.append((len(ps), sum(len(p) for p in ps)))
        return per, sum(len(ph) for ph in per_str)

    def _fill(self, phrases, nums, lens, counts, cells):
        flat = np.array([c for p in phrases for c in p], NP_CELL[self.w])
        lens[: len(phrases)] = torch.tensor([len(p) for p in phrases], dtype=torch.int32)
        counts[: len(phrases)] = torch.tensor(nums, dtype=torch.int64)
        raw = torch.from_numpy(flat.view(np.uint8).copy())
        cells[: raw.numel()] = raw

    def pack(self, lens, counts, cells):
        _, cnt, order = self._local
        self._fill(order, [cnt[p] for p in order], lens, counts, cells)

    def _unpack(self, lens, nums, cells, m, n_cells):
        ln = lens[:m].tolist()
        flat = cells[: n_cells * self.w].numpy().view(NP_CELL[self.w])
        out, o = [], 0
        for L in ln:
            out.append(tuple(int(x) for x in flat[o:o + L]))
            o += L
        return out, [int(x) for x in nums[:m].tolist()]

    def merge(self, lens, counts, cells, m, n_cells):
        ph, nums = self._unpack(lens, counts, cells, m, n_cells)
        acc = Counter()
        for p, c in zip(ph, nums):
            acc[p] += c
        self._part = acc
        dense = {p: i for i, p in enumerate(acc)}  # pack_part emits the partition in this order
        self._recv_dense = [dense[p] for p in ph]
        return len(acc), sum(len(p) for p in acc)

    def pack_part(self, lens, freqs, cells):
        ph = list(self._part)
        self._fill(ph, [self._part[p] for p in ph], lens, freqs, cells)

    # ---- A.3 / A.5 on the global dictionary, A.4 on the shard ----
    def global_round(self, lens, freqs, cells, d, n_cells, done):
        ph, fr = self._unpack(lens, freqs, cells, d, n_cells)
        assert len(set(ph)) == len(ph), "the gathered dictionary holds a phrase twice"
        vals = {p: tuple(self._val(c) for c in p) for p in ph}
        freq = {vals[p]: f for p, f in zip(ph, fr)}
        A = self.A
        groups = {}
        for p, f in freq.items():
            for k in range(len(p)):
                suf = p[k:]
                if len(suf) == 1 and not self._suffix(suf[0]):
                    continue
                g = groups.setdefault(suf, {"left": set(), "full": False, "n": 0, "freq": 0})
                g["n"] += 1
                g["freq"] += f
                if k == 0:
                    g["full"] = True
                    g["left"].add(-1)
                else:
                    g["left"].add(p[k - 1])
        bwt_dummy, hocc_dummy = A + 1, A + 2
        rank, meta, pre, rank_of, has_hocc, reps = 0, {}, [], {}, [], []
        for suf in sorted(groups, key=functools.cmp_to_key(cmp_pg)):
            g = groups[suf]
            if len(g["left"]) > 1 or g["full"]:
                if g["full"]:
                    meta[suf] = (rank << 1) | (1 if freq[suf] > 1 else 0)
                rank_of[suf] = rank
                has_hocc.append(1 if g["n"] > 1 else 0)
                reps.append(suf)
                sym = hocc_dummy if g["n"] > 1 else bwt_dummy
                rank += 1
            else:
                sym = next(iter(g["left"]))
            if pre and pre[-1][0] == sym:
                pre[-1][1] += g["freq"]
            else:
                pre.append([sym, g["freq"]])
        tot = rank
        alph3, dummy = A + 3, A + 3 + tot + 1
        rl, rr = [], []
        for suf in reps:
            if len(suf) == 1:
                rl.append(dummy); rr.append(suf[0]); continue
            k = 1
            while True:
                s2 = suf[k:]
                marked = s2 in rank_of and has_hocc[rank_of[s2]]
                if marked or k == len(suf) - 1:
                    break
                k += 1
            if marked:
                rl.append(suf[k - 1]); rr.append(alph3 + rank_of[suf[k:]])
            else:
                rl.append(dummy); rr.append(suf[-1] if self._suffix(suf[-1]) else suf[-2])
        self._level = {"rule_l": np.array(rl, np.uint64), "rule_r": np.array(rr, np.uint64), "has_hocc": np.array(has_hocc, np.uint8),
                       "pre_sym": np.array([p[0] for p in pre], np.uint64), "pre_len": np.array([p[1] for p in pre], np.uint64)}
        new_suffix = {}
        for p in freq:
            new_suffix[meta[p] >> 1] = self._suffix(p[-1])
        # rewrite the shard
        per_str, cnt, _ = self._local
        new_cells, new_ptrs = [], [0]
        for phs in per_str:
            for p in phs:
                new_cells.append(meta[vals[p]] if p in vals else meta[tuple(self._val(c) for c in p)])
            new_ptrs.append(len(new_cells))
        bps = sym_width(tot) + 1
        info = {"alphabet": A, "tot_phrases": tot, "n_phrases": len(freq), "dict_syms": sum(len(p) for p in freq), "parse_len": len(new_cells),
                "n_in": len(self.cells), "done": int(done), "cell_bytes_out": 1 if bps <= 8 else 2 if bps <= 16 else 4 if bps <= 32 else 8}
        self.cells, self.str_ptrs, self.first = new_cells, new_ptrs, False
        self.w = info["cell_bytes_out"]
        self.A, self.is_suffix = tot, new_suffix
        return info

    # ---- the same round with the ranking split over the ranks (contiguous slices of the sorted suffix groups) ----
    def _global_groups(self, lens, freqs, cells, d, n_cells):
        ph, fr = self._unpack(lens, freqs, cells, d, n_cells)
        vals = [tuple(self._val(c) for c in p) for p in ph]
        off, o = [], 0
        for p in vals:
            off.append(o)
            o += len(p)
        groups = {}
        for idx, (p, f) in enumerate(zip(vals, fr)):
            for k in range(len(p)):
                suf = p[k:]
                if len(suf) == 1 and not self._suffix(suf[0]):
                    continue
                g = groups.setdefault(suf, {"left": set(), "full": None, "n": 0, "freq": 0, "members": []})
                g["n"] += 1
                g["freq"] += f
                g["members"].append(off[idx] + k)
                if k == 0:
                    g["full"] = idx
                    g["left"].add(-1)
                else:
                    g["left"].add(p[k - 1])
        order = sorted(groups, key=functools.cmp_to_key(cmp_pg))
        return ph, vals, fr, off, o, groups, order

    def rank_sort(self, lens, freqs, cells, d, n_cells, rank_id, n_ranks):
        if not getattr(self, "distribute", False):
            return [0, 0, 0, 0, 8]
        ph, vals, fr, off, nE, groups, order = self._global_groups(lens, freqs, cells, d, n_cells)
        lo, hi = rank_id * len(order) // n_ranks, (rank_id + 1) * len(order) // n_ranks
        A = self.A
        mine, pre = [], []
        for suf in order[lo:hi]:
            g = groups[suf]
            ranked = len(g["left"]) > 1 or g["full"] is not None
            sym = (A + 2 if g["n"] > 1 else A + 1) if ranked else next(iter(g["left"]))
            if ranked:
                mine.append(suf)
            if pre and pre[-1][0] == sym:
                pre[-1][1] += g["freq"]
            else:
                pre.append([sym, g["freq"]])
        self._dist = {"ph": ph, "vals": vals, "fr": fr, "off": off, "groups": groups, "mine": mine, "pre": pre}
        return [1, len(mine), len(pre), nE, 8]

    def rank_apply(self, rank_base, ph_meta, isn, erank1):
        D = self._dist
        for u, suf in enumerate(D["mine"]):
            g, r = D["groups"][suf], rank_base + u
            if g["full"] is not None:
                idx = g["full"]
                ph_meta[idx] = (r << 1) | (1 if D["fr"][idx] > 1 else 0)
                isn[r] = 1 if self._suffix(suf[-1]) else 0
            if g["n"] > 1:
                for e in g["members"]:
                    erank1[e] = r + 1

    def reply(self, part_base, ph_meta, reply):
        for k, i in enumerate(self._recv_dense):
            reply[k] = int(ph_meta[part_base + i])

    def rank_finish(self, rank_base, tot, n_pre, ph_meta, isn, erank1, done, local_meta=None):
        D, A = self._dist, self.A
        alph3, dummy = A + 3, A + 3 + tot + 1
        rl, rr, hh = [], [], []
        for suf in D["mine"]:
            g = D["groups"][suf]
            hh.append(1 if g["n"] > 1 else 0)
            e0 = g["members"][0]
            if len(suf) == 1:
                rl.append(dummy); rr.append(suf[0]); continue
            k = 1
            while True:
                marked = int(erank1[e0 + k]) != 0
                if marked or k == len(suf) - 1:
                    break
                k += 1
            if marked:
                rl.append(suf[k - 1]); rr.append(alph3 + int(erank1[e0 + k]) - 1)
            else:
                rl.append(dummy); rr.append(suf[-1] if self._suffix(suf[-1]) else suf[-2])
        self._slice = {"rule_l": rl, "rule_r": rr, "has_hocc": hh, "pre_sym": [p[0] for p in D["pre"]], "pre_len": [p[1] for p in D["pre"]]}
        per_str, cnt, order = self._local
        if local_meta is not None:  # the owners returned the metasymbols in pack order
            meta_of = {p: int(local_meta[k]) for k, p in enumerate(order)}
        else:
            gidx = {p: i for i, p in enumerate(D["ph"])}
            meta_of = {p: int(ph_meta[gidx[p]]) for p in cnt}
        new_cells, new_ptrs = [], [0]
        for phs in per_str:
            for p in phs:
                new_cells.append(meta_of[p])
            new_ptrs.append(len(new_cells))
        bps = sym_width(tot) + 1
        info = {"alphabet": A, "tot_phrases": tot, "n_phrases": len(D["ph"]), "dict_syms": sum(len(p) for p in D["ph"]), "parse_len": len(new_cells),
                "n_in": len(self.cells), "done": int(done), "n_pre_runs": n_pre, "cell_bytes_out": 1 if bps <= 8 else 2 if bps <= 16 else 4 if bps <= 32 else 8}
        self.cells, self.str_ptrs, self.first = new_cells, new_ptrs, False
        self.w = info["cell_bytes_out"]
        self.A = tot
        self.is_suffix = {r: bool(int(isn[r])) for r in range(tot)}
        return info

    def level_slice(self, rl, rr, hh, ps, pl):
        S = self._slice
        for t, key in ((rl, "rule_l"), (rr, "rule_r"), (hh, "has_hocc"), (ps, "pre_sym"), (pl, "pre_len")):
            if S[key]:
                t[: len(S[key])] = torch.tensor(S[key], dtype=t.dtype)

    level_override = None

    def fetch_level(self):
        if self.level_override is not None:  # assembled from per-rank slices by multigpu._ranked_distributed (tensors)
            from grlbwt_b200.multigpu import LEVEL_KEYS, unsigned_numpy
            return {k: unsigned_numpy(self.level_override[k]) for k in LEVEL_KEYS}
        return dict(self._level)

    def fetch_parse(self):
        return np.array(self.cells, NP_CELL[self.w])
