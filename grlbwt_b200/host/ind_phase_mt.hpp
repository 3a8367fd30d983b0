// Multi-threaded induction phase on the host. Same semantics as ind_phase.hpp (the
// reference's infer_lvl_bwt, lib/exact_algo/exact_ind_phase.cpp:111-386), reorganised so that every step is
// data parallel over std::thread workers:
//   A  chain expansion: each run of BWT_{i+1} follows its grammar chain and emits (bucket, left symbol, length)
//      tuples; the run keeps the chain's terminal symbol                         (:143-258)
//   B  the hocc buffer = the tuples in bucket order, input order inside a bucket = a STABLE parallel LSD radix
//      sort on the bucket id (11-bit digits, per-thread histograms)             (replaces compute_hocc_size :42-109
//      + the in-place bucket fills)
//   C  prefix sums: symbol offsets of the rewritten BWT_{i+1} stream, of the hocc buffer, and of the part of the
//      hocc buffer that redirects to the stream; from them the stream / hocc offset where every preliminary-BWT
//      run starts
//   D  assembly: ranges of preliminary-BWT runs are expanded independently into thread-local run lists
//      (solved runs, stream copies, hocc copies) and concatenated with boundary merging  (:287-361)
// Symbols are 32-bit here (levels whose alphabet needs 64 bits use the sequential ind_phase.hpp).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include <sys/mman.h>

#include "rl_bwt_io.hpp"

namespace grlbwt {

// uninitialised array: pages are first touched by the worker threads that fill them, not zeroed serially
template <class T>
struct RawBuf {
    T* p = nullptr;
    size_t n = 0;
    RawBuf() = default;
    explicit RawBuf(size_t count) { alloc(count); }
    RawBuf(const RawBuf&) = delete;
    RawBuf& operator=(const RawBuf&) = delete;
    RawBuf(RawBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    RawBuf& operator=(RawBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    void alloc(size_t count) {
        release();
        n = count;
        const size_t bytes = (count ? count : 1) * sizeof(T);
        constexpr size_t HUGE = (size_t)2 << 20;
        if (bytes >= 2 * HUGE && !getenv("GRLBWT_NO_HUGE")) {  // big arrays: 2 MB pages (512 x fewer first-touch faults and TLB misses in the scatter passes)
            p = (T*)aligned_alloc(HUGE, (bytes + HUGE - 1) / HUGE * HUGE);
            if (p) madvise((void*)p, (bytes + HUGE - 1) / HUGE * HUGE, MADV_HUGEPAGE);
        } else p = (T*)malloc(bytes);
        if (!p) throw std::bad_alloc();
    }
    void release() { free(p); p = nullptr; n = 0; }
    ~RawBuf() { release(); }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    void resize(size_t count) { alloc(count); }  // contents are NOT preserved and NOT initialised
    const T* begin() const { return p; }
    const T* end() const { return p + n; }
    template <class It>
    void assign(It first, It last) {
        alloc((size_t)(last - first));
        size_t i = 0;
        for (It it = first; it != last; ++it) p[i++] = (T)*it;
    }
};

struct Level32 {
    uint64_t alphabet = 0, tot_phrases = 0;
    // uninitialised buffers (no serial zero-fill of GBs before the device copies land in them)
    RawBuf<uint32_t> rule_l, rule_r;
    RawBuf<uint8_t> has_hocc;
    RawBuf<uint32_t> pre_sym;
    RawBuf<uint64_t> pre_len;    // run lengths, 64-bit ...
    RawBuf<uint32_t> pre_len32;  // ... or 32-bit when the level's lengths fit (one of the two is empty)
    inline uint64_t len(size_t i) const { return pre_len32.empty() ? pre_len[i] : (uint64_t)pre_len32[i]; }
};

struct Runs32 {
    std::vector<uint32_t> sym;
    std::vector<uint64_t> len;
    size_t size() const { return sym.size(); }
    inline void push(uint32_t s, uint64_t l) {
        if (l == 0) return;
        if (!sym.empty() && sym.back() == s) { len.back() += l; return; }
        sym.push_back(s);
        len.push_back(l);
    }
};
// a level's BWT as flat arrays (what the parallel steps read and the concatenation writes)
struct RunArr {
    RawBuf<uint32_t> sym;
    RawBuf<uint64_t> len;
    RawBuf<uint32_t> len32;  // the level-0 BWT of the device induction comes with 32-bit run lengths (len stays empty then)
    size_t n = 0;
    size_t size() const { return n; }
    uint64_t length(size_t i) const { return len32.empty() ? len[i] : (uint64_t)len32[i]; }
};

template <class F>
inline void parallel_chunks(size_t n_threads, size_t n_items, F&& fn, size_t min_items = 4096) {  // fn(tid, begin, end)
    if (n_threads <= 1 || n_items < min_items) { fn(0, 0, n_items); return; }
    std::vector<std::thread> th;
    const size_t per = (n_items + n_threads - 1) / n_threads;
    for (size_t t = 0; t < n_threads; t++) {
        const size_t b = std::min(n_items, t * per), e = std::min(n_items, b + per);
        if (b == e) break;  // chunks past the end are empty: no thread, so per-chunk state is only touched by chunks that exist
        th.emplace_back([&fn, t, b, e] { fn(t, b, e); });
    }
    for (auto& x : th) x.join();
}

// exclusive prefix sums of a[0..n) into out[0..n], out[n] = total (two-pass, parallel)
template <class GetT>
inline void parallel_prefix(size_t n_threads, size_t n, GetT&& get, RawBuf<uint64_t>& out) {
    out.alloc(n + 1);
    const size_t T = (n_threads <= 1 || n < 4096) ? 1 : n_threads;
    std::vector<uint64_t> part(T + 1, 0);
    parallel_chunks(T, n, [&](size_t t, size_t b, size_t e) {  // partials indexed by the chunk id the scheduler hands out
        uint64_t s = 0;
        for (size_t i = b; i < e; i++) s += get(i);
        part[t + 1] = s;
    });
    for (size_t t = 0; t < T; t++) part[t + 1] += part[t];
    parallel_chunks(T, n, [&](size_t t, size_t b, size_t e) {
        uint64_t s = part[t];
        for (size_t i = b; i < e; i++) { out[i] = s; s += get(i); }
    });
    out[n] = part[T];
}

// one entry of a level's hocc buffer: bucket g, symbol l (or FROM_BWT32), run length f. FT = uint32_t (12-byte tuples: a
// quarter less traffic in the sort and the prefix passes) whenever every run length of the level fits, else uint64_t
template <class FT>
struct HTupleT { uint32_t g, l; FT f; };
constexpr uint32_t FROM_BWT32 = 0xffffffffu;

// stable LSD radix sort of tuples by g (11-bit digits)
template <class HTuple>
inline void radix_sort_tuples(size_t n_threads, RawBuf<HTuple>& a, uint64_t max_key) {
    const size_t n = a.size();
    if (n < 2) return;
    RawBuf<HTuple> b(n);
    int key_bits = 0;
    while (key_bits < 32 && (max_key >> key_bits)) key_bits++;
    // as few passes as 11-bit digits allow, then digits of equal width: 15 key bits sort as 8 + 7, not 11 + 4 (fewer open
    // cache lines per thread in each scatter)
    const int n_pass = std::max(1, (key_bits + 10) / 11), BITS = std::max(1, (key_bits + n_pass - 1) / n_pass), NB = 1 << BITS;
    const size_t T = (n_threads <= 1 || n < (1u << 16)) ? 1 : n_threads;
    std::vector<uint64_t> hist(T * NB);
    HTuple* src = a.data();
    HTuple* dst = b.data();
    for (int shift = 0; shift < key_bits; shift += BITS) {
        std::fill(hist.begin(), hist.end(), 0);
        parallel_chunks(T, n, [&](size_t c, size_t bg, size_t en) {
            uint64_t* h = hist.data() + c * NB;
            for (size_t i = bg; i < en; i++) h[(src[i].g >> shift) & (NB - 1)]++;
        });
        uint64_t acc = 0;
        for (int d = 0; d < NB; d++)
            for (size_t t = 0; t < T; t++) { const uint64_t c = hist[t * NB + d]; hist[t * NB + d] = acc; acc += c; }
        parallel_chunks(T, n, [&](size_t c, size_t bg, size_t en) {
            uint64_t* h = hist.data() + c * NB;
            for (size_t i = bg; i < en; i++) dst[h[(src[i].g >> shift) & (NB - 1)]++] = src[i];
        });
        std::swap(src, dst);
    }
    if (src != a.data()) std::swap(a, b);
}

struct PhaseClock {  // GRLBWT_TRACE=1 prints the time of every step of every level to stderr
    bool on;
    std::chrono::steady_clock::time_point t;
    PhaseClock() : on(getenv("GRLBWT_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what, size_t n) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "    [induction] %-22s %10zu items %8.1f ms\n", what, n, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// one level step BWT_{i+1} -> BWT_i
template <class FT>
inline RunArr induce_level_t(RunArr& bwt, const Level32& L, size_t n_threads) {
    typedef HTupleT<FT> HTuple;
    PhaseClock clk;
    const uint64_t A = L.alphabet, alph3 = A + 3, bwt_dummy = A + 1, hocc_dummy = A + 2;
    const size_t m = bwt.size();
    const size_t T = std::max<size_t>(1, n_threads);
    uint32_t* bsym = bwt.sym.data();
    const uint64_t* blen = bwt.len.data();

    // ---- A: chain expansion. Pass 1 counts the tuples of every chunk, pass 2 writes them at the chunk's offset of the
    //         final array and leaves the chain's terminal symbol in the run ----
    const size_t n_chunks = (T == 1 || m < 4096) ? 1 : T * 8;
    const size_t per_chunk = (m + n_chunks - 1) / n_chunks;
    std::vector<uint64_t> cbase(n_chunks + 1, 0);
    auto run_chunks = [&](auto&& work) {
        if (n_chunks == 1) { work(0); return; }
        std::vector<std::thread> th;
        std::atomic<size_t> ticket{0};
        for (size_t t = 0; t < T; t++)
            th.emplace_back([&] { for (size_t c = ticket.fetch_add(1); c < n_chunks; c = ticket.fetch_add(1)) work(c); });
        for (auto& x : th) x.join();
    };
    run_chunks([&](size_t c) {
        const size_t b = std::min(m, c * per_chunk), e = std::min(m, b + per_chunk);
        uint64_t cnt = 0;
        for (size_t i = b; i < e; i++) {
            const uint32_t P = bsym[i];
            cnt += L.has_hocc[P];
            uint32_t r = L.rule_r[P];
            while (r >= alph3) { cnt++; r = L.rule_r[r - alph3]; }
        }
        cbase[c + 1] = cnt;
    });
    for (size_t c = 0; c < n_chunks; c++) cbase[c + 1] += cbase[c];
    const size_t n_tuples = cbase[n_chunks];
    RawBuf<HTuple> hocc(n_tuples);
    run_chunks([&](size_t c) {
        const size_t b = std::min(m, c * per_chunk), e = std::min(m, b + per_chunk);
        HTuple* out = hocc.data() + cbase[c];
        for (size_t i = b; i < e; i++) {
            const uint32_t P = bsym[i];
            const uint64_t f = blen[i];
            if (L.has_hocc[P]) *out++ = {P, FROM_BWT32, (FT)f};
            uint32_t l = L.rule_l[P], r = L.rule_r[P];
            while (r >= alph3) {
                const uint32_t g = (uint32_t)(r - alph3);
                *out++ = {g, l, (FT)f};
                l = L.rule_l[g];
                r = L.rule_r[g];
            }
            bsym[i] = r;
        }
    });
    clk.lap("chain expansion", m);

    // ---- B: bucket order, input order inside a bucket ----
    radix_sort_tuples(T, hocc, L.tot_phrases ? L.tot_phrases - 1 : 0);
    clk.lap("radix by bucket", n_tuples);

    // ---- C: offsets ----
    RawBuf<uint64_t> cum_s, cum_h, cum_hb;  // stream symbols before run i; hocc symbols before entry k; of which from the stream
    parallel_prefix(T, m, [&](size_t i) { return blen[i]; }, cum_s);
    {   // both prefix sums over the tuples in one two-pass sweep (the tuples are the big array of a level)
        cum_h.alloc(n_tuples + 1);
        cum_hb.alloc(n_tuples + 1);
        const size_t Tp = (T <= 1 || n_tuples < 4096) ? 1 : T;
        std::vector<uint64_t> ph(Tp + 1, 0), pb(Tp + 1, 0);
        parallel_chunks(Tp, n_tuples, [&](size_t c, size_t b, size_t e) {
            uint64_t sh = 0, sb = 0;
            for (size_t k = b; k < e; k++) { sh += hocc[k].f; sb += hocc[k].l == FROM_BWT32 ? hocc[k].f : 0; }
            ph[c + 1] = sh; pb[c + 1] = sb;
        });
        for (size_t t = 0; t < Tp; t++) { ph[t + 1] += ph[t]; pb[t + 1] += pb[t]; }
        parallel_chunks(Tp, n_tuples, [&](size_t c, size_t b, size_t e) {
            uint64_t sh = ph[c], sb = pb[c];
            for (size_t k = b; k < e; k++) {
                cum_h[k] = sh; cum_hb[k] = sb;
                sh += hocc[k].f; sb += hocc[k].l == FROM_BWT32 ? hocc[k].f : 0;
            }
        });
        cum_h[n_tuples] = ph[Tp]; cum_hb[n_tuples] = pb[Tp];
    }
    auto from_bwt_before = [&](uint64_t x) -> uint64_t {  // stream symbols consumed by the first x symbols of the hocc buffer
        if (x == 0) return 0;
        const size_t k = (size_t)(std::upper_bound(cum_h.data(), cum_h.data() + n_tuples + 1, x - 1) - cum_h.data()) - 1;  // entry holding symbol x-1
        return cum_hb[k] + (hocc[k].l == FROM_BWT32 ? x - cum_h[k] : 0);
    };
    const size_t n_pre = L.pre_sym.size();
    RawBuf<uint64_t> pre_h, pre_s, pre_o;  // hocc offset, stream offset, output symbol offset at the start of pre-run i
    parallel_prefix(T, n_pre, [&](size_t i) { return L.pre_sym[i] == hocc_dummy ? L.len(i) : 0; }, pre_h);
    if (pre_h[n_pre] != cum_h[n_tuples]) throw std::runtime_error("induction: hocc buffer and preliminary BWT disagree");
    parallel_prefix(T, n_pre, [&](size_t i) -> uint64_t {
        if (L.pre_sym[i] == bwt_dummy) return L.len(i);
        if (L.pre_sym[i] == hocc_dummy) return from_bwt_before(pre_h[i] + L.len(i)) - from_bwt_before(pre_h[i]);
        return 0;
    }, pre_s);
    if (pre_s[n_pre] != cum_s[m]) throw std::runtime_error("induction: stream and preliminary BWT disagree");
    parallel_prefix(T, n_pre, [&](size_t i) { return L.len(i); }, pre_o);
    clk.lap("offsets", n_pre);

    // ---- D: assembly. Parts are ranges of OUTPUT symbols of equal size (a long pre-run may be cut in the middle), so
    //         the work is balanced even when a few preliminary runs cover most of the level ----
    const uint64_t n_out = pre_o[n_pre];
    const size_t n_parts = (T == 1 || n_out < (1u << 16)) ? 1 : T * 8;
    std::vector<Runs32> parts(n_parts);
    auto assemble = [&](size_t p) {
        Runs32& out = parts[p];
        const uint64_t o_beg = n_out / n_parts * p, o_end = p + 1 == n_parts ? n_out : n_out / n_parts * (p + 1);
        if (o_beg >= o_end) return;
        {   // a part cannot hold more runs than output symbols, and the level cannot hold more than its sources together
            const uint64_t guess = std::min<uint64_t>(o_end - o_beg, (uint64_t)(n_tuples + m + n_pre) / n_parts + 1024);
            out.sym.reserve(guess);
            out.len.reserve(guess);
        }

        // first pre-run that reaches into [o_beg, o_end), and how many of its symbols lie before o_beg
        size_t i = (size_t)(std::upper_bound(pre_o.data(), pre_o.data() + n_pre + 1, o_beg) - pre_o.data()) - 1;
        uint64_t skip = o_beg - pre_o[i];
        // stream / hocc cursors at the start of pre-run i, advanced by `skip` symbols of that run
        uint64_t soff = pre_s[i], hoff = pre_h[i];
        if (skip) {
            if (L.pre_sym[i] == bwt_dummy) soff += skip;
            else if (L.pre_sym[i] == hocc_dummy) { soff += from_bwt_before(hoff + skip) - from_bwt_before(hoff); hoff += skip; }
        }
        size_t sp = soff < cum_s[m] ? (size_t)(std::upper_bound(cum_s.data(), cum_s.data() + m + 1, soff) - cum_s.data()) - 1 : m;
        uint64_t s_used = sp < m ? soff - cum_s[sp] : 0;  // symbols of run sp already consumed
        size_t hk = hoff < cum_h[n_tuples] ? (size_t)(std::upper_bound(cum_h.data(), cum_h.data() + n_tuples + 1, hoff) - cum_h.data()) - 1 : n_tuples;
        uint64_t h_used = hk < n_tuples ? hoff - cum_h[hk] : 0;
        auto take = [&](uint64_t f) {  // extract_rl_syms, :19-40
            while (f) {
                const uint64_t avail = blen[sp] - s_used, t = avail < f ? avail : f;
                out.push(bsym[sp], t);
                f -= t;
                s_used += t;
                if (s_used == blen[sp]) { sp++; s_used = 0; }
            }
        };
        uint64_t o = o_beg;
        for (; o < o_end; i++) {
            const uint32_t s = L.pre_sym[i];
            uint64_t f = L.len(i) - skip;
            skip = 0;
            if (f > o_end - o) f = o_end - o;
            o += f;
            if (s == bwt_dummy) take(f);
            else if (s == hocc_dummy) {
                while (f) {
                    const uint64_t avail = hocc[hk].f - h_used, t = avail < f ? avail : f;
                    if (hocc[hk].l == FROM_BWT32) take(t); else out.push(hocc[hk].l, t);
                    f -= t;
                    h_used += t;
                    if (h_used == hocc[hk].f) { hk++; h_used = 0; }
                }
            } else out.push(s, f);
        }
    };
    if (n_parts == 1) assemble(0);
    else {
        std::vector<std::thread> th;
        std::atomic<size_t> ticket{0};
        for (size_t t = 0; t < T; t++)
            th.emplace_back([&] { for (size_t p = ticket.fetch_add(1); p < n_parts; p = ticket.fetch_add(1)) assemble(p); });
        for (auto& x : th) x.join();
    }
    clk.lap("assembly", n_pre);
    // concatenate; runs are maximal inside a part, so merging is only needed where two parts meet
    struct Piece { size_t part, from, to, base; };
    std::vector<Piece> pieces;
    size_t out_total = 0;
    {
        int64_t tail = -1;  // part that owns the current last run of the output
        for (size_t p = 0; p < n_parts; p++) {
            const size_t sz = parts[p].size();
            if (sz == 0) continue;
            size_t from = 0;
            if (tail >= 0 && parts[(size_t)tail].sym.back() == parts[p].sym[0]) {
                parts[(size_t)tail].len.back() += parts[p].len[0];
                from = 1;
            }
            if (from < sz) {
                pieces.push_back({p, from, sz, out_total});
                out_total += sz - from;
                tail = (int64_t)p;
            }
        }
    }
    RunArr out;
    out.n = out_total;
    out.sym.alloc(out_total);
    out.len.alloc(out_total);
    parallel_chunks(std::min(T, std::max<size_t>(1, pieces.size())), pieces.size(), [&](size_t, size_t b, size_t e) {
        for (size_t q = b; q < e; q++) {
            const Piece& pc = pieces[q];
            memcpy(out.sym.data() + pc.base, parts[pc.part].sym.data() + pc.from, (pc.to - pc.from) * sizeof(uint32_t));
            memcpy(out.len.data() + pc.base, parts[pc.part].len.data() + pc.from, (pc.to - pc.from) * sizeof(uint64_t));
            std::vector<uint32_t>().swap(parts[pc.part].sym);
            std::vector<uint64_t>().swap(parts[pc.part].len);
        }
    }, 2);
    clk.lap("concat output", out_total);
    return out;
}

// deepest level: the final parse (cells >> 1) in string order, run-length encoded (parse2bwt_int, exact_ind_phase.cpp:621-635)
inline RunArr parse_to_bwt32(const uint64_t* parse, uint64_t n) {
    Runs32 r;
    for (uint64_t i = 0; i < n; i++) r.push((uint32_t)(parse[i] >> 1), 1);
    RunArr a;
    a.n = r.size();
    a.sym.alloc(a.n);
    a.len.alloc(a.n);
    if (a.n) { memcpy(a.sym.data(), r.sym.data(), a.n * 4); memcpy(a.len.data(), r.len.data(), a.n * 8); }
    return a;
}

// ind_phase (exact_ind_phase.cpp:674-697): levels[0] is round 1
inline RunArr induce_level_mt(RunArr& bwt, const Level32& L, size_t n_threads) {
    uint64_t longest = 0;  // the tuples copy run lengths of BWT_{i+1}
    for (size_t i = 0; i < bwt.size(); i++) longest = std::max(longest, bwt.len[i]);
    return longest < (1ull << 32) ? induce_level_t<uint32_t>(bwt, L, n_threads) : induce_level_t<uint64_t>(bwt, L, n_threads);
}

inline RunArr ind_phase_mt(const std::vector<Level32>& levels, const uint64_t* final_parse, uint64_t n_strings, size_t n_threads) {
    RunArr bwt = parse_to_bwt32(final_parse, n_strings);
    for (size_t lv = levels.size(); lv-- > 0;) {
        RunArr next = induce_level_mt(bwt, levels[lv], n_threads);
        bwt = std::move(next);
    }
    return bwt;
}

}  // namespace grlbwt
