// Induction phase on the host (fresh implementation of the behaviour of the reference's
// exact_algo::ind_phase, lib/exact_algo/exact_ind_phase.cpp:674-697; semantics: SURVEY.md App. B):
// from the deepest parse back to level 0, BWT_i is induced from BWT_{i+1}, the level's grammar
// rules / hocc marks and its preliminary BWT, all produced on the device by the parse phase.
// Everything is kept in host RAM as (symbol, length) run arrays; the semi-external block cache of
// the reference (include/bwt_io.h) is not reproduced. The `-b` byte cap of the reference only
// bounds its in-RAM run lengths and never changes the output, so lengths are plain u64 here.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

#include "rl_bwt_io.hpp"

namespace grlbwt {

// artefacts of one parse round (what the reference stores in dict_lev_k + pre_bwt_lev_k)
struct Level {
    uint64_t alphabet = 0;     // A of the round's input text
    uint64_t tot_phrases = 0;  // ranks of the round
    std::vector<uint64_t> rule_l, rule_r;  // exact_par_phase.cpp:33-87 conventions: alphabet A+3, dummy A+3+tot+1
    std::vector<uint8_t> has_hocc;
    std::vector<uint64_t> pre_sym, pre_len;  // maximal runs; A+1 = from BWT_{i+1}, A+2 = from the hocc buffer
};

// deepest level: the final parse (cells >> 1) in string order, run-length encoded (parse2bwt_int, exact_ind_phase.cpp:621-635)
template <class CellT>
inline RunList parse_to_bwt(const CellT* parse, uint64_t n) {
    RunList r;
    for (uint64_t i = 0; i < n; i++) r.push((uint64_t)parse[i] >> 1, 1);
    return r;
}

// one level step BWT_{i+1} -> BWT_i (infer_lvl_bwt, exact_ind_phase.cpp:111-386)
inline RunList induce_level(RunList& bwt, const Level& L) {
    const uint64_t A = L.alphabet, alph3 = A + 3, bwt_dummy = A + 1, hocc_dummy = A + 2, tot = L.tot_phrases;
    const uint64_t FROM_BWT = ~0ULL;
    const uint64_t n_runs = bwt.size();

    // counting pre-pass (compute_hocc_size, :42-109): upper bound of the runs appended to each hocc bucket
    std::vector<uint64_t> off(tot + 1, 0);
    for (uint64_t i = 0; i < n_runs; i++) {
        uint64_t P = bwt.sym[i];
        if (L.has_hocc[P]) off[P + 1]++;
        uint64_t r = L.rule_r[P];
        while (r >= alph3) {
            const uint64_t g = r - alph3;
            off[g + 1]++;
            r = L.rule_r[g];
        }
    }
    for (uint64_t g = 0; g < tot; g++) off[g + 1] += off[g];
    std::vector<uint64_t> fill(off.begin(), off.end() - 1);
    std::vector<uint64_t> hs(off[tot]), hl(off[tot], 0);
    auto append = [&](uint64_t g, uint64_t s, uint64_t f) {
        uint64_t& q = fill[g];
        if (q > off[g] && hs[q - 1] == s) hl[q - 1] += f;
        else { hs[q] = s; hl[q] = f; q++; }
    };

    // induction pass (:143-258): feed the buckets along the grammar chain; the run keeps the chain's terminal symbol
    for (uint64_t i = 0; i < n_runs; i++) {
        const uint64_t P = bwt.sym[i], f = bwt.len[i];
        if (L.has_hocc[P]) append(P, FROM_BWT, f);
        uint64_t l = L.rule_l[P], r = L.rule_r[P];
        while (r >= alph3) {
            const uint64_t g = r - alph3;
            append(g, l, f);
            l = L.rule_l[g];
            r = L.rule_r[g];
        }
        bwt.sym[i] = r;
    }

    // assembly along the preliminary BWT (:287-361)
    RunList out;
    uint64_t sp = 0;             // next run of the rewritten BWT_{i+1} stream
    uint64_t hg = 0, hq = 0;     // next hocc entry: bucket, slot
    if (tot) hq = off[0];
    auto take = [&](uint64_t f) {  // extract_rl_syms, :19-40
        while (f) {
            const uint64_t t = bwt.len[sp] < f ? bwt.len[sp] : f;
            out.push(bwt.sym[sp], t);
            f -= t;
            bwt.len[sp] -= t;
            if (bwt.len[sp] == 0) sp++;
        }
    };
    const uint64_t n_pre = L.pre_sym.size();
    for (uint64_t i = 0; i < n_pre; i++) {
        const uint64_t s = L.pre_sym[i];
        uint64_t f = L.pre_len[i];
        if (s == bwt_dummy) take(f);
        else if (s == hocc_dummy) {
            while (f) {
                while (hq == fill[hg]) {
                    if (++hg >= tot) throw std::runtime_error("induction: hocc buffer exhausted");
                    hq = off[hg];
                }
                const uint64_t t = hl[hq] < f ? hl[hq] : f;
                if (hs[hq] == FROM_BWT) take(t); else out.push(hs[hq], t);
                hl[hq] -= t;
                f -= t;
                if (hl[hq] == 0) hq++;
            }
        } else out.push(s, f);
    }
    return out;
}

// ind_phase (exact_ind_phase.cpp:674-697): levels[0] is round 1
template <class CellT>
inline RunList ind_phase(const std::vector<Level>& levels, const CellT* final_parse, uint64_t n_strings) {
    RunList bwt = parse_to_bwt<CellT>(final_parse, n_strings);
    for (size_t lv = levels.size(); lv-- > 0;) {
        RunList next = induce_level(bwt, levels[lv]);
        bwt = std::move(next);
    }
    return bwt;
}

}  // namespace grlbwt
