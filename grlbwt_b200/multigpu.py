"""Multi-GPU parse phase: one process per GPU, shards of whole strings, one
hash-partitioned all-to-all-v + one all-gather-v of the round's dictionary over NCCL per round.

torch.distributed is plumbing here: every device step before, between and after the collectives is a
call into libgrlgpu.so (grlgpu_mg_* in include/grlgpu.h). The orchestration is written against a small
engine interface so that the same code runs under gloo on CPU tensors with a test double
(tests/cpu_engine.py) -- that is how the N>1 logic is covered without GPUs.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

CELL_T = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}


def shard_bounds(text: np.ndarray, n_ranks: int):
    """Contiguous ranges of WHOLE strings balanced by symbol count (the reference's mt split,
    parsing_strategies.h:208-214, but byte-balanced). -> list of (begin, end) cell offsets."""
    sep = text[-1]
    ends = np.flatnonzero(text == sep) + 1  # one past each string
    if ends.size < n_ranks:
        raise ValueError(f"{ends.size} strings cannot be split over {n_ranks} ranks")
    bounds, begin = [], 0
    for r in range(n_ranks):
        if r == n_ranks - 1:
            end = int(text.size)
        else:
            target = (r + 1) * text.size // n_ranks
            k = int(np.searchsorted(ends, target))
            k = min(max(k, r), ends.size - (n_ranks - r))  # every remaining rank keeps at least one string
            end = int(ends[k])
            if end <= begin:
                end = int(ends[np.searchsorted(ends, begin + 1)])
        bounds.append((begin, end))
        begin = end
    return bounds


class GpuEngine:
    """The device side of one rank: a thin adapter from tensors to the C ABI."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device

    def alloc(self, n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype, device=self.device)

    def stats(self):
        s = self.ctx.stats()
        return {k: getattr(s, k) for k, _ in s._fields_}

    def histogram(self):
        return self.ctx.histogram()

    def set_alphabet(self, max_sym):
        self.ctx.mg_set_alphabet(max_sym)

    def cell_bytes(self):
        return self.ctx.cell_bytes()

    def local(self, n_ranks):
        return self.ctx.mg_local(n_ranks)

    def pack(self, lens, counts, cells):
        self.ctx.mg_pack(lens.data_ptr(), counts.data_ptr(), cells.data_ptr())

    def merge(self, lens, counts, cells, m, n_cells):
        return self.ctx.mg_merge(lens.data_ptr(), counts.data_ptr(), cells.data_ptr(), m, n_cells)

    def pack_part(self, lens, freqs, cells):
        self.ctx.mg_pack_part(lens.data_ptr(), freqs.data_ptr(), cells.data_ptr())

    def global_round(self, lens, freqs, cells, d, n_cells, done):
        return self.ctx.mg_global(lens.data_ptr(), freqs.data_ptr(), cells.data_ptr(), d, n_cells, done)

    def rank_sort(self, lens, freqs, cells, d, n_cells, rank_id, n_ranks):
        return self.ctx.mg_rank_sort(lens.data_ptr(), freqs.data_ptr(), cells.data_ptr(), d, n_cells, rank_id, n_ranks)

    def rank_apply(self, rank_base, ph_meta, isn, erank1):
        self.ctx.mg_rank_apply(rank_base, ph_meta.data_ptr(), isn.data_ptr(), erank1.data_ptr())

    def reply(self, part_base, ph_meta, reply):
        self.ctx.mg_reply(part_base, ph_meta.data_ptr(), reply.data_ptr())

    def rank_finish(self, rank_base, tot, n_pre, ph_meta, isn, erank1, done, local_meta=None):
        return self.ctx.mg_rank_finish(rank_base, tot, n_pre, ph_meta.data_ptr(), isn.data_ptr(), erank1.data_ptr(), done,
                                       None if local_meta is None else local_meta.data_ptr())

    def level_slice(self, rl, rr, hh, ps, pl):
        self.ctx.mg_level_slice(rl.data_ptr(), rr.data_ptr(), hh.data_ptr(), ps.data_ptr(), pl.data_ptr())

    level_override = None  # set by distributed_round when the level was assembled from per-rank slices (device tensors)

    def fetch_level(self, arena=None, **k):
        """level artefacts as numpy arrays; with `arena` (pinned uint8 numpy buffer) the copies land there directly"""
        ov = self.level_override
        if ov is None:
            return self.ctx.fetch_level(arena, **k) if arena is not None else self.ctx.fetch_level(**k)
        out, off = {}, 0
        for key in LEVEL_KEYS:
            t = ov[key]
            if arena is not None and t.is_cuda and t.numel():
                nb = t.numel() * t.element_size()
                off = (off + 15) & ~15
                if off + nb > arena.size:
                    raise ValueError("fetch arena too small")
                torch.from_numpy(arena[off:off + nb]).view(t.dtype).copy_(t)  # device -> (pinned) host, no staging copy
                out[key] = arena[off:off + nb].view(_UNSIGNED[t.dtype])
                off += nb
            else:
                out[key] = unsigned_numpy(t)
        return out

    def fetch_parse(self):
        return self.ctx.fetch_parse()


LEVEL_KEYS = ("rule_l", "rule_r", "has_hocc", "pre_sym", "pre_len")
_UNSIGNED = {torch.int32: np.uint32, torch.int64: np.uint64, torch.uint8: np.uint8}


def unsigned_numpy(t):
    """torch tensors carry the library's u32 / u64 values as int32 / int64 bit patterns: host copy, reinterpreted"""
    return t.cpu().numpy().view(_UNSIGNED[t.dtype])


def merge_seams(PS, PL, counts):
    """PS / PL: concatenation, in rank order, of every rank's preliminary-BWT runs (each slice maximal on its own).
    Equal symbols can only meet where two slices meet: at most len(counts) - 1 merges, decided from 2 symbols per slice."""
    n = sum(counts)
    PS, PL = PS[:n], PL[:n]
    spans, off = [], 0
    for c in counts:
        if c:
            spans.append((off, off + c - 1))
        off += c
    if len(spans) < 2:
        return PS, PL
    idx = torch.tensor([x for ab in spans for x in ab], dtype=torch.int64, device=PS.device)
    sy = PS[idx].cpu().tolist()
    drops, target, last_sym = [], spans[0][1], sy[1]
    for k in range(1, len(spans)):
        f, l = spans[k]
        if sy[2 * k] == last_sym:
            drops.append((f, target))      # the slice's first run continues the run `target`
            if l == f:
                continue                   # a one-run slice: the open run stays the same
        target, last_sym = l, sy[2 * k + 1]
    if not drops:
        return PS, PL
    PL = PL.clone()
    for j, t in drops:
        PL[t] += PL[j]
    pieces, a = [], 0
    for j, _ in drops:
        pieces.append((a, j))
        a = j + 1
    pieces.append((a, n))
    return torch.cat([PS[a:b] for a, b in pieces]), torch.cat([PL[a:b] for a, b in pieces])


def _staged():
    """gloo cannot move CUDA tensors in every collective: stage through the host there (tests on a 1-GPU box)"""
    return dist.get_backend() == "gloo"


def _fence(t):
    """NCCL collectives are ordered on torch's stream only; the library may run on another stream, so make the
    result visible to the host (and thereby to any stream) before it is handed to the library"""
    if t.is_cuda:
        torch.cuda.synchronize(t.device)


def _all_reduce(t, op=dist.ReduceOp.SUM):
    if _staged() and t.is_cuda:
        c = t.cpu()
        dist.all_reduce(c, op=op)
        t.copy_(c)
    else:
        dist.all_reduce(t, op=op)


def _broadcast(t, src):
    if _staged() and t.is_cuda:
        c = t.cpu()
        dist.broadcast(c, src=src)
        t.copy_(c)
    else:
        dist.broadcast(t, src=src)
        _fence(t)


def _all_gather_small(t):
    c = t.cpu() if (_staged() and t.is_cuda) else t
    out = [torch.empty_like(c) for _ in range(dist.get_world_size())]
    dist.all_gather(out, c)
    return [x.cpu().tolist() for x in out]


def _all_to_all_single(recv, send, recv_counts=None, send_counts=None):
    if _staged() and send.is_cuda:
        cs, cr = send.cpu(), torch.empty(recv.shape, dtype=recv.dtype)
        dist.all_to_all_single(cr, cs, recv_counts, send_counts)
        recv.copy_(cr)
    else:
        dist.all_to_all_single(recv, send, recv_counts, send_counts)
        _fence(recv)


def _all_to_all_v(send, send_counts, recv_counts, engine):
    recv = engine.alloc(sum(recv_counts), send.dtype)
    _all_to_all_single(recv[: sum(recv_counts)], send[: sum(send_counts)], list(recv_counts), list(send_counts))
    return recv


def _all_gather_v(part, counts, engine):
    """concatenation of every rank's `part[:counts[rank]]`, in rank order, on every rank"""
    out = engine.alloc(sum(counts), part.dtype)
    off = 0
    me = dist.get_rank()
    if not _staged() and part.is_cuda and max(counts) > 0:
        # NCCL: ONE all-gather of slices padded to the longest part (hash partitions are near-equal), then a
        # device-side compaction; one collective at NVSwitch bandwidth instead of a chain of broadcasts
        mx = max(counts)
        send = torch.empty(mx, dtype=part.dtype, device=part.device)
        send[: counts[me]].copy_(part[: counts[me]])
        padded = torch.empty(mx * len(counts), dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(padded, send)
        for r, c in enumerate(counts):
            if c:
                out[off: off + c].copy_(padded[r * mx: r * mx + c])
            off += c
        _fence(out)
        return out
    for r, c in enumerate(counts):
        if c:
            sl = out[off: off + c]
            if r == me:
                sl.copy_(part[:c])
            _broadcast(sl, r)
        off += c
    return out


def _ranked_distributed(engine, glens, gfreqs, gcells, d, n_cells, done, info5, want_level, exch):
    """the dictionary ranking split over the ranks by first-key range (grlgpu_mg_rank_*): two small all-gathers,
    three all-reduce(MAX), the metasymbol return (owners answer the packs they received: reverse all-to-all-v), and
    -- only when the level is wanted -- an all-gather-v of the level slices"""
    G, me = dist.get_world_size(), dist.get_rank()
    n_ranked, n_pre_loc, nE, sym_bytes = info5[1:5]
    allc = _all_gather_small(torch.tensor([n_ranked, n_pre_loc], dtype=torch.int64, device=engine.device))
    ranked_counts = [int(x[0]) for x in allc]
    pre_counts = [int(x[1]) for x in allc]
    base, tot = sum(ranked_counts[:me]), sum(ranked_counts)
    ph_meta = engine.alloc(d, torch.int64).zero_()
    isn = engine.alloc(tot, torch.uint8).zero_()
    erank1 = engine.alloc(nE, torch.int32).zero_()
    if ph_meta.is_cuda:
        torch.cuda.synchronize(ph_meta.device)
    engine.rank_apply(base, ph_meta, isn, erank1)
    for t in (ph_meta, isn, erank1):
        _all_reduce(t, dist.ReduceOp.MAX)
        _fence(t)
    sent_phr, recv_phr, part_phr = exch  # phrases per peer: sent to / received from as an owner; partition sizes in rank order
    local_meta = None
    if hasattr(engine, "reply"):
        reply = engine.alloc(sum(recv_phr), torch.int64)
        engine.reply(sum(part_phr[:me]), ph_meta, reply)
        local_meta = _all_to_all_v(reply, recv_phr, sent_phr, engine)
    info = engine.rank_finish(base, tot, sum(pre_counts), ph_meta, isn, erank1, done, local_meta)
    engine.level_override = None
    if want_level:
        st = torch.int32 if sym_bytes == 4 else torch.int64
        rl, rr = engine.alloc(n_ranked, st), engine.alloc(n_ranked, st)
        hh = engine.alloc(n_ranked, torch.uint8)
        ps, pl = engine.alloc(n_pre_loc, st), engine.alloc(n_pre_loc, torch.int64)
        engine.level_slice(rl, rr, hh, ps, pl)
        RL, RR, HH = _all_gather_v(rl, ranked_counts, engine), _all_gather_v(rr, ranked_counts, engine), _all_gather_v(hh, ranked_counts, engine)
        PS, PL = _all_gather_v(ps, pre_counts, engine), _all_gather_v(pl, pre_counts, engine)
        if me == 0:  # stays on the device until somebody fetches it
            S, Ln = merge_seams(PS, PL, pre_counts)
            engine.level_override = {"rule_l": RL[:tot], "rule_r": RR[:tot], "has_hocc": HH[:tot], "pre_sym": S, "pre_len": Ln}
            info["n_pre_runs"] = int(S.numel())
    info["ranking"] = "distributed"
    return info


def distributed_round(engine, n_strings_global: int, timings: dict | None = None, want_level: bool = True):
    """One parse round over all ranks. -> (round info dict, done)"""
    G = dist.get_world_size()
    w = engine.cell_bytes()
    per_owner, parse_len_local = engine.local(G)  # [(n_phrases, n_cells)] * G
    t = torch.tensor([parse_len_local], dtype=torch.int64, device=engine.device)
    _all_reduce(t)
    done = int(t.item()) == n_strings_global

    # ---- hash-partitioned all-to-all-v of the local dictionaries ----
    n_phr = [p[0] for p in per_owner]
    n_cel = [p[1] for p in per_owner]
    lens = engine.alloc(sum(n_phr), torch.int32)
    counts = engine.alloc(sum(n_phr), torch.int64)
    cells = engine.alloc(sum(n_cel) * w, torch.uint8)
    engine.pack(lens, counts, cells)
    sizes = torch.tensor([[a, b] for a, b in zip(n_phr, n_cel)], dtype=torch.int64, device=engine.device)
    rsizes = torch.empty_like(sizes)
    _all_to_all_single(rsizes, sizes)
    rs = rsizes.cpu().tolist()
    r_phr = [int(x[0]) for x in rs]
    r_cel = [int(x[1]) for x in rs]
    rlens = _all_to_all_v(lens, n_phr, r_phr, engine)
    rcounts = _all_to_all_v(counts, n_phr, r_phr, engine)
    rcells = _all_to_all_v(cells, [c * w for c in n_cel], [c * w for c in r_cel], engine)
    exchanged = (sum(n_phr) * 12 + sum(n_cel) * w)

    # ---- owner-side dedup, then all-gather-v of the deduplicated partitions ----
    d_part, c_part = engine.merge(rlens, rcounts, rcells, sum(r_phr), sum(r_cel))
    plens = engine.alloc(d_part, torch.int32)
    pfreqs = engine.alloc(d_part, torch.int64)
    pcells = engine.alloc(c_part * w, torch.uint8)
    engine.pack_part(plens, pfreqs, pcells)
    mine = torch.tensor([d_part, c_part], dtype=torch.int64, device=engine.device)
    allsz = _all_gather_small(mine)
    g_phr = [int(x[0]) for x in allsz]
    g_cel = [int(x[1]) for x in allsz]
    glens = _all_gather_v(plens, g_phr, engine)
    gfreqs = _all_gather_v(pfreqs, g_phr, engine)
    gcells = _all_gather_v(pcells, [c * w for c in g_cel], engine)
    gathered = (sum(g_phr) * 12 + sum(g_cel) * w)

    info5 = engine.rank_sort(glens, gfreqs, gcells, sum(g_phr), sum(g_cel), dist.get_rank(), G) if hasattr(engine, "rank_sort") else [0] * 5
    if info5[0]:
        info = _ranked_distributed(engine, glens, gfreqs, gcells, sum(g_phr), sum(g_cel), done, info5, want_level, (n_phr, r_phr, g_phr))
    else:
        info = engine.global_round(glens, gfreqs, gcells, sum(g_phr), sum(g_cel), done)
        info["ranking"] = "replicated"
        if hasattr(engine, "level_override"):
            engine.level_override = None
    info["exchange_bytes_sent"] = exchanged
    info["gather_bytes"] = gathered
    if timings is not None:
        timings.setdefault("rounds", []).append(info)
    return info, done


def global_stats(engine):
    """collection statistics over all ranks (collection_stats, utils.cpp:100-189) + the global alphabet set in the engine"""
    st = engine.stats()
    dev = engine.device
    red = torch.tensor([st["max_sym"], -int(st["min_sym"]), -int(st["sep_sym"]), int(st["sep_sym"])], dtype=torch.int64, device=dev)
    _all_reduce(red, dist.ReduceOp.MAX)
    max_sym, min_sym, sep_lo, sep_hi = int(red[0]), -int(red[1]), -int(red[2]), int(red[3])
    if sep_lo != sep_hi or sep_lo != min_sym:
        raise ValueError("the collection is ill formed: ranks disagree on the separator or it is not the smallest symbol")
    sums = torch.tensor([st["n_syms"], st["n_strings"]], dtype=torch.int64, device=dev)
    _all_reduce(sums)
    n_syms, n_strings = int(sums[0]), int(sums[1])
    longest = torch.tensor([st["longest_string"]], dtype=torch.int64, device=dev)
    _all_reduce(longest, dist.ReduceOp.MAX)
    if engine.cell_bytes() == 1:  # byte alphabet: highest count of the GLOBAL histogram (utils.cpp:161-175)
        h = torch.from_numpy(engine.histogram().astype(np.int64)).to(dev)
        _all_reduce(h)
        max_sym_freq = int(h.max())
    else:
        max_sym_freq = n_syms  # utils.cpp:117
    engine.set_alphabet(max_sym)
    return {"n_syms": n_syms, "n_strings": n_strings, "longest_string": int(longest[0]), "min_sym": min_sym, "max_sym": max_sym,
            "max_sym_freq": max_sym_freq, "sep_sym": sep_lo}


def gather_final_parse(engine):
    """final parse: one cell per string, gathered to rank 0 in rank (= string) order"""
    dev = engine.device
    fp = np.ascontiguousarray(engine.fetch_parse()).astype(np.int64)
    mine = torch.tensor([fp.size], dtype=torch.int64, device=dev)
    cnts = [int(x[0]) for x in _all_gather_small(mine)]
    part = torch.from_numpy(fp).to(dev)
    full = _all_gather_v(part, cnts, engine)
    return full[: sum(cnts)].cpu().numpy().astype(np.uint64) if dist.get_rank() == 0 else None


def par_phase_distributed(engine, collect_levels: bool = True):
    """The whole parse phase over all ranks. Every rank must have its shard set in the engine already.
    -> dict(stats, levels (rank 0 only), final_parse (rank 0 only, string order), rounds)"""
    me = dist.get_rank()
    gstats = global_stats(engine)
    levels, rounds = [], []
    while True:
        info, done = distributed_round(engine, gstats["n_strings"], want_level=collect_levels)
        rounds.append(info)
        if collect_levels and me == 0:
            L = engine.fetch_level()
            L["alphabet"], L["tot"] = info["alphabet"], info["tot_phrases"]
            levels.append(L)
        if done:
            break
    return {"stats": gstats, "levels": levels, "final_parse": gather_final_parse(engine), "rounds": rounds}
