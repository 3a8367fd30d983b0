// C ABI (include/grlgpu.h) and host-side sequencing of one parse round on one B200.
// Replaces the strategy + par_round pair of the reference (lib/exact_algo/exact_par_phase.cpp:265-497).
#include "../../include/grlgpu.h"
#include "util.cuh"
#include "primitives.cuh"
#include "parse_kernels.cuh"
#include "dict_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

using namespace grl;


namespace { struct MgRound; }

struct grlgpu_ctx {
    int device = 0;
    u64 flags = 0;
    cudaStream_t st = nullptr;
    bool own_stream = true;
    std::string last_error;
    DevicePool pool;  // declared before every DevBuf member: destroyed after them
    Profiler prof;

    // current text
    const void* text = nullptr;  // device
    DevBuf<u8> text_own;         // owns `text` unless borrowed
    u64 n = 0;
    int w = 0;
    bool first = true;
    DevBuf<u32> end_bits;   // string-end bitmap of the current text (rounds >= 2)
    DevBuf<u8> is_suffix;   // per symbol of the current alphabet (rounds >= 2)
    u64 alphabet = 0;       // A
    u64 n_strings = 0;
    u64 sep = 0;
    int round = 0;
    bool done = false;
    bool have_stats = false;
    grlgpu_stats_t stats{};
    u64 hist[256] = {0};  // byte alphabets: histogram of the (local) text

    // artefacts of the last round (device)
    int lvl_sym_bytes = 4;
    u64 lvl_tot = 0, lvl_npre = 0, lvl_n_in = ~0ull;  // upper bound of the sum of the level's run lengths (all ones: unknown)
    DevBuf<u8> rule_l, rule_r, has_hocc, pre_sym;
    DevBuf<u64> pre_len;

    // asynchronous level fetches: a second stream copies while the next round computes; the level's device buffers
    // are parked here (not returned to the pool) until grlgpu_fetch_wait
    cudaStream_t copy_st = nullptr;
    cudaEvent_t copy_ev = nullptr;
    std::vector<DevBuf<u8>> parked_u8;
    std::vector<DevBuf<u64>> parked_u64;
    std::vector<DevBuf<u32>> parked_u32;

    // multi-GPU round in flight (between grlgpu_mg_local and grlgpu_mg_global)
    MgRound* mg = nullptr;
    int mg_ranks = 0;

    // optional: dictionary of the last round kept for tests (GRLGPU_FLAG_KEEP_DICT)
    u64 kd_d = 0, kd_nE = 0, kd_nS = 0;
    DevBuf<u8> kd_D;
    DevBuf<u32> kd_off, kd_len, kd_order, kd_phr_of;
    DevBuf<u64> kd_freq, kd_meta;
};

namespace {

struct Timer {
    cudaEvent_t a{}, b{};
    cudaStream_t st;
    explicit Timer(cudaStream_t s) : st(s) {
        GRL_CUDA(cudaEventCreate(&a));
        GRL_CUDA(cudaEventCreate(&b));
    }
    ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { GRL_CUDA(cudaEventRecord(a, st)); }
    void stop() { GRL_CUDA(cudaEventRecord(b, st)); }
    float ms() {
        GRL_CUDA(cudaEventSynchronize(b));
        float t = 0;
        GRL_CUDA(cudaEventElapsedTime(&t, a, b));
        return t;
    }
};

template <class T>
T d2h_scalar(const T* dptr, cudaStream_t st) {
    static_assert(sizeof(T) % 4 == 0, "d2h_scalar reads whole words");
    T h;
    d2h_small(&h, dptr, sizeof(T), st);
    return h;
}

inline unsigned grid_for(u64 n, int threads) { return (unsigned)std::max<u64>(1, div_up(n, (u64)threads)); }

static __global__ void table_init_kernel(ulonglong2* table, u64 cap) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) table[i] = make_ulonglong2(HT_EMPTY, 0ULL);
}
static __global__ void next_start_bits_kernel(const u32* __restrict__ end_bits, u64 n, u32* __restrict__ start_bits) {
    // start_bits[q] = (q == 0) || end_bits[q-1]   (string starts of the current text)
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 n_words = (n + 31) / 32;
    if (w >= n_words) return;
    u32 cur = end_bits[w];
    u32 prev = w ? end_bits[w - 1] : 0x80000000u;
    u32 s = (cur << 1) | (prev >> 31);
    if (w == n_words - 1 && (n & 31)) s &= (1u << (n & 31)) - 1u;
    start_bits[w] = s;
}
template <class CellT>
__global__ void __launch_bounds__(256) first_round_end_bits_kernel(const CellT* __restrict__ text, u64 n, CellT sep, u32* __restrict__ end_bits) {
    // thread = 32 cells = one bitmap word, 16-byte vector loads
    const u64 word = (u64)blockIdx.x * blockDim.x + threadIdx.x, base = word * 32;
    if (base >= n) return;
    const u64 cnt = n - base;
    u32 bits = 0;
    if (cnt >= 32) {
        __align__(16) CellT c[32];
        constexpr int NV = 32 * sizeof(CellT) / 16;
        const uint4* src = reinterpret_cast<const uint4*>(text + base);
        uint4* dst = reinterpret_cast<uint4*>(c);
#pragma unroll
        for (int k = 0; k < NV; k++) dst[k] = src[k];
#pragma unroll
        for (int k = 0; k < 32; k++) bits |= (u32)(c[k] == sep) << k;
    } else {
        for (u64 k = 0; k < cnt; k++) bits |= (u32)(text[base + k] == sep) << k;
    }
    end_bits[word] = bits;
}
static __global__ void adjacent_max_diff_kernel(const u64* __restrict__ ptrs, u64 n_str, u64* out) {
    u64 m = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_str; i += (u64)gridDim.x * blockDim.x) {
        const u64 dlt = ptrs[i + 1] - ptrs[i];
        m = dlt > m ? dlt : m;
    }
    m = warp_max(m);
    if (lane_id() == 0 && m) atomicMax(out, m);
}
static __global__ void u32_to_u64_kernel(const u32* __restrict__ in, u64 n, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// state shared by the stages of one round
struct Round {
    grlgpu_ctx* c;
    cudaStream_t st;
    u64 n, p = 0, d = 0, nE = 0, cap = 0, G = 0, tot = 0, n_pre = 0, max_freq = 0, max_len = 0;
    DevBuf<u32> start_bits, end_bits_first;
    const u32* end_bits = nullptr;  // of the input text
    DevBuf<ulonglong2> table;
    DevBuf<u32> slot_of_phrase, occ_slots;
    DevBuf<u64> ph_pos, ph_freq;
    DevBuf<u32> ph_len, ph_off;
    DevBuf<u8> D_raw;
    DevBuf<u32> phr_of, rem, rank, order;
    DevBuf<ulonglong2> einfo;
    // suffix-order plan: K symbol codes of sym_bits bits per 64-bit key; ext_mode = refinement by key extension, over the
    // nS valid entries only (first keys + entry ids come out of the dictionary gather); else prefix doubling over all nE
    int sym_bits = 0, K = 1, spare = 0;  // spare: bits of the first key beyond the K whole codes (key extension only)
    bool ext_mode = false;
    u64 nS = 0;
    DevBuf<u64> keys;
    DevBuf<u32> vals;
    const void* dict_text = nullptr;  // text the dictionary's ph_pos point into (default: the context's text)
    u64* ph_meta = nullptr;           // if set, metasymbols go here (per phrase) instead of into the table
    explicit Round(grlgpu_ctx* ctx) : c(ctx), st(ctx->st), n(ctx->n) {}
};

// ---------------- text stage: boundary flags -> phrase starts -> dedup table -> distinct list ----------------
template <class CellT, bool FIRST>
void stage_flags(Round& R) {
    grlgpu_ctx* c = R.c;
    const CellT* text = (const CellT*)c->text;
    const u64 n_words = div_up(R.n, 32);
    const u64 n_blocks = div_up(n_words, LMS_THREADS);
    R.start_bits.alloc(n_words, R.st);
    u32* end_out = nullptr;
    if (FIRST) {
        R.end_bits_first.alloc(n_words, R.st);
        end_out = R.end_bits_first.p;
        R.end_bits = R.end_bits_first.p;
    } else R.end_bits = c->end_bits.p;
    const u32* end_in = FIRST ? nullptr : c->end_bits.p;
    DevBuf<u32> need_slow(1, R.st);
    need_slow.zero();
    const CellT sep = (CellT)c->sep;
    bool slow = (c->flags & GRLGPU_FLAG_FORCE_SLOW_SCAN) != 0;
    if (!slow) {
        GRL_LAUNCH("lms_flags", R.n * sizeof(CellT) + R.n / 4, (lms_flags_kernel<CellT, FIRST, 0>), (unsigned)n_blocks, LMS_THREADS, 0, R.st, text, R.n, sep, end_in, end_out, R.start_bits.p, nullptr, nullptr, need_slow.p);
        slow = d2h_scalar(need_slow.p, R.st) != 0;
    }
    if (slow) {  // a run of equal cells crosses a CTA boundary by more than the look-ahead
        DevBuf<u8> state(n_blocks, R.st), incoming(n_blocks, R.st);
        GRL_LAUNCH("lms_flags", R.n * sizeof(CellT) + R.n / 4, (lms_flags_kernel<CellT, FIRST, 1>), (unsigned)n_blocks, LMS_THREADS, 0, R.st, text, R.n, sep, end_in, end_out, R.start_bits.p, state.p, nullptr, need_slow.p);
        GRL_LAUNCH("lms_resolve", 0, lms_resolve_kernel, 1, 32, 0, R.st, state.p, incoming.p, n_blocks);
        GRL_LAUNCH("lms_flags", R.n * sizeof(CellT) + R.n / 4, (lms_flags_kernel<CellT, FIRST, 2>), (unsigned)n_blocks, LMS_THREADS, 0, R.st, text, R.n, sep, end_in, end_out, R.start_bits.p, nullptr, incoming.p, need_slow.p);
        GRL_CUDA(cudaStreamSynchronize(R.st));
    }
}

void dict_offsets(Round& R);

template <class CellT, class PosT>
void insert_uncached(Round& R, BitmapCompactor& bc, DevBuf<u8>& ps_raw, u64 j0, u64 cap, u32* overflow) {
    grlgpu_ctx* c = R.c;
    if (!ps_raw.p) {  // compacted phrase starts (top bit: starts a string), sentinel ps[p] = n | FLAG
        ps_raw.alloc((R.p + 1) * sizeof(PosT), R.st);
        bc.write<PosT>(R.end_bits, (PosT*)ps_raw.p);
        const PosT sentinel = (PosT)R.n | PosFlag<PosT>::FLAG;
        GRL_CUDA(cudaMemcpyAsync((PosT*)ps_raw.p + R.p, &sentinel, sizeof(PosT), cudaMemcpyHostToDevice, R.st));
        GRL_CUDA(cudaStreamSynchronize(R.st));
    }
    const u64 cnt = R.p - j0;
    GRL_LAUNCH("phrase_insert", (R.n + cnt) * sizeof(CellT) + cnt * (2 * sizeof(PosT) + 4 + 32), (phrase_insert_kernel<CellT, PosT>), grid_for(cnt, 256), 256, 0, R.st,
               (const CellT*)c->text, R.n, (const PosT*)ps_raw.p, j0, R.p, R.start_bits.p, R.end_bits, R.table.p, cap, R.slot_of_phrase.p, overflow);
}

template <class CellT>
void stage_dedup(Round& R) {
    grlgpu_ctx* c = R.c;
    const CellT* text = (const CellT*)c->text;
    // phrase numbering: per-tile popcounts of the start bitmap -> exclusive scan = index of a tile's first phrase
    constexpr int TW = fd_tile_words<CellT>();
    const u64 n_words = div_up(R.n, 32), n_tiles = div_up(n_words, TW);
    DevBuf<u32> tile_cnt(n_tiles, R.st);
    DevBuf<u64> tile_base(n_tiles, R.st), ptot(1, R.st);
    GRL_LAUNCH("tile_popc", R.n / 8, tile_popc_kernel, (unsigned)n_tiles, 256, 0, R.st, R.start_bits.p, n_words, TW, tile_cnt.p);
    exclusive_scan<u32, u64>(tile_cnt.p, tile_base.p, n_tiles, ptot.p, R.st);
    R.p = d2h_scalar(ptot.p, R.st);
    static int n_sm = 0;  // one per CellT instantiation
    if (!n_sm) {
        GRL_CUDA(cudaFuncSetAttribute(dedup_cached_kernel<CellT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fd_smem_bytes<CellT>()));
        GRL_CUDA(cudaFuncSetAttribute(dedup_cached_kernel<CellT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        GRL_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->device));
    }
    const unsigned fd_grid = (unsigned)std::min<u64>(n_tiles, (u64)n_sm * FD_CTAS_PER_SM);  // persistent CTAs, tiles strided
    // the first tiles always run through the cached kernel and report how many phrases missed the caches;
    // the rest of the text takes the cached kernel (duplicate-heavy) or the thread-per-phrase kernel (unique-heavy)
    const bool force_uncached = (c->flags & GRLGPU_FLAG_FORCE_UNCACHED) != 0;
    const u64 t_pilot = force_uncached ? 0 : std::min<u64>(n_tiles, (c->flags & GRLGPU_FLAG_SMALL_PILOT) ? 1ull : 4ull * fd_grid);
    u64 j_pilot = R.p;
    if (t_pilot < n_tiles) j_pilot = t_pilot ? d2h_scalar(tile_base.p + t_pilot, R.st) : 0;
    BitmapCompactor bc;
    DevBuf<u8> ps_raw;
    bool counted = false;

    R.slot_of_phrase.alloc(R.p, R.st);
    DevBuf<u32> overflow(1, R.st);
    DevBuf<u64> stats(2, R.st);
    // capacity (any multiple of 256; slot = mulhi(hash, cap)). "full" = 1.6 p slots: load <= 0.625, the pass cannot
    // overflow (beyond 2^31 slots, i.e. p > 1.3e9: start at 2^28 and regrow on demand). A text long enough to have a
    // pilot starts SMALL instead (1.6 x the pilot's phrases): if the pilot says duplicate-heavy the dictionary is tiny
    // next to p and the small table is kept (no 30 GB init + occupancy scan for a 16 k-phrase dictionary); if it says
    // unique-heavy the pass restarts on the full table with the thread-per-phrase kernel.
    const u64 cap_max = (1ull << 31) - 256;
    const u64 want = std::max<u64>(1024, (R.p + R.p / 2 + R.p / 10 + 255) / 256 * 256);
    const u64 cap_full = want <= cap_max ? want : (1ull << 28);
    const bool has_rest = t_pilot < n_tiles;
    u64 cap = cap_full;
    if (t_pilot && has_rest) cap = std::min<u64>(cap_full, std::max<u64>(1ull << 20, (j_pilot + j_pilot / 2 + j_pilot / 10 + 255) / 256 * 256));
    if (c->flags & GRLGPU_FLAG_SMALL_TABLE) cap = 1024;
    int decided = has_rest ? (t_pilot ? -1 : 0) : 1;  // -1: ask the pilot, 1: cached kernel for the rest, 0: thread-per-phrase
    for (;;) {
        if (cap > cap_max) throw Error(GRLGPU_ERR_LIMIT, "phrase table would exceed 2^31 slots");
        R.table.alloc(cap, R.st);
        GRL_LAUNCH("table_init", cap * 16, table_init_kernel, grid_for(cap, 256), 256, 0, R.st, R.table.p, cap);
        overflow.zero();
        stats.zero();
        const bool run_pilot = t_pilot && decided != 0;
        if (run_pilot) {
            const u64 cells = std::min<u64>(R.n, t_pilot * TW * 32);
            GRL_LAUNCH("dedup_cached", cells * sizeof(CellT) + cells / 4 + j_pilot * 4, (dedup_cached_kernel<CellT>), fd_grid, FD_THREADS, fd_smem_bytes<CellT>(), R.st,
                       text, R.n, R.start_bits.p, R.end_bits, tile_base.p, (u64)0, t_pilot, R.table.p, cap, R.slot_of_phrase.p, overflow.p, stats.p);
        }
        if (has_rest) {
            if (decided == -1) {
                u64 hs[2];
                d2h_small(hs, stats.p, 16, R.st);
                GRL_CUDA(cudaStreamSynchronize(R.st));
                decided = hs[0] * 2 <= hs[1] ? 1 : 0;  // cached: at most half of the pilot's phrases had to go to the global table
                if (decided == 0 && cap < cap_full && !(c->flags & GRLGPU_FLAG_SMALL_TABLE)) { cap = cap_full; continue; }
            }
            if (decided == 1) {
                const u64 cells = R.n - t_pilot * TW * 32;
                GRL_LAUNCH("dedup_cached", cells * sizeof(CellT) + cells / 4 + (R.p - j_pilot) * 4, (dedup_cached_kernel<CellT>), fd_grid, FD_THREADS, fd_smem_bytes<CellT>(),
                           R.st, text, R.n, R.start_bits.p, R.end_bits, tile_base.p, t_pilot, n_tiles, R.table.p, cap, R.slot_of_phrase.p, overflow.p, (u64*)nullptr);
            } else {
                if (!counted) {
                    const u64 pc = bc.count(R.start_bits.p, R.n, R.st);
                    if (pc != R.p) throw Error(GRLGPU_ERR_STATE, "phrase count mismatch between tile counts and compaction");
                    counted = true;
                }
                const u64 j0 = run_pilot ? j_pilot : 0;
                if (R.n < (1ull << 31)) insert_uncached<CellT, u32>(R, bc, ps_raw, j0, cap, overflow.p);
                else insert_uncached<CellT, u64>(R, bc, ps_raw, j0, cap, overflow.p);
            }
        }
        bool ovf = d2h_scalar(overflow.p, R.st) != 0;
        u64 d = 0;
        if (!ovf) {
            DevBuf<u32> occ_bits(cap / 32, R.st);
            GRL_LAUNCH("table_occupancy", cap * 16, table_occupancy_kernel, (unsigned)(cap / 256), 256, 0, R.st, R.table.p, cap, occ_bits.p);
            BitmapCompactor oc;
            d = oc.count(occ_bits.p, cap, R.st);
            if (d * 10 <= cap * 7 || cap >= want) {  // load factor <= 0.7 (or the table can no longer be too small)
                R.occ_slots.alloc(d, R.st);
                oc.write<u32>(nullptr, R.occ_slots.p);
                R.d = d;
                R.cap = cap;
                break;
            }
        }
        if (cap >= cap_max) throw Error(GRLGPU_ERR_LIMIT, "phrase table would exceed 2^31 slots");
        cap = std::min<u64>(cap * 8, std::min<u64>(want, cap_max));  // too full: regrow and redo the pass
    }
    ps_raw.release();
    R.ph_pos.alloc(R.d, R.st);
    R.ph_len.alloc(R.d, R.st);
    R.ph_freq.alloc(R.d, R.st);
    GRL_LAUNCH("dict_meta", 0, dict_meta_kernel, grid_for(R.d, 256), 256, 0, R.st, R.table.p, R.occ_slots.p, R.d, R.start_bits.p, R.end_bits, R.n, R.ph_pos.p, R.ph_len.p, R.ph_freq.p);
    dict_offsets(R);
    R.start_bits.release();
}

// phrase offsets inside the dictionary, number of entries, highest frequency, longest phrase
void dict_offsets(Round& R) {
    R.ph_off.alloc(R.d + 1, R.st);
    DevBuf<u64> tot64(1, R.st);
    {   // offsets as u64 first to detect overflow of the 32-bit entry index space
        DevBuf<u64> off64(R.d, R.st);
        exclusive_scan<u32, u64>(R.ph_len.p, off64.p, R.d, tot64.p, R.st);
        R.nE = d2h_scalar(tot64.p, R.st);
        if (R.nE >= 0xfffffff0ull) throw Error(GRLGPU_ERR_LIMIT, "dictionary of this round exceeds 2^32 symbols");
    }
    exclusive_scan<u32, u32>(R.ph_len.p, R.ph_off.p, R.d, R.ph_off.p + R.d, R.st);
    DevBuf<u64> mx(2, R.st);
    mx.zero();
    GRL_LAUNCH("reduce_max_u64", 0, reduce_max_u64_kernel, 296, 256, 0, R.st, R.ph_freq.p, R.d, mx.p);
    {
        DevBuf<u64> len64(R.d, R.st);
        GRL_LAUNCH("u32_to_u64", 0, u32_to_u64_kernel, grid_for(R.d, 256), 256, 0, R.st, R.ph_len.p, R.d, len64.p);
        GRL_LAUNCH("reduce_max_u64", 0, reduce_max_u64_kernel, 296, 256, 0, R.st, len64.p, R.d, mx.p + 1);
    }
    u64 hmx[2];
    d2h_small(hmx, mx.p, 16, R.st);
    R.max_freq = hmx[0];
    R.max_len = hmx[1];
}

template <class CellT, bool FIRST, class SymT>
void stage_gather(Round& R, bool force_ext = false) {
    grlgpu_ctx* c = R.c;
    R.sym_bits = bit_width64(c->alphabet + 1);
    R.K = std::max(1, 64 / R.sym_bits);
    // refinement by key extension needs ceil((longest phrase + 1) / K) passes at most; beyond 64 passes (or when a test
    // forces it) the groups are refined by prefix doubling on position-based ranks instead
    R.ext_mode = force_ext || (!(c->flags & GRLGPU_FLAG_FORCE_DOUBLING) && (R.max_len + 1 + (u64)R.K - 1) / (u64)R.K <= 64);
    R.spare = (R.ext_mode && R.sym_bits * R.K < 64) ? 64 - R.sym_bits * R.K : 0;
    const CellT* dtext = (const CellT*)(R.dict_text ? R.dict_text : c->text);
    DevBuf<u32> voff;
    R.nS = R.nE;
    if (R.ext_mode) {
        voff.alloc(R.d + 1, R.st);
        DevBuf<u32> tot(1, R.st);
        IsSuffix isuf0{c->is_suffix.p, c->sep, c->first};
        GRL_LAUNCH("phrase_vlen", R.d * 48, (phrase_vlen_kernel<CellT, FIRST>), grid_for(R.d, 256), 256, 0, R.st, dtext, R.ph_pos.p, R.ph_len.p, R.d, isuf0, voff.p);
        exclusive_scan<u32, u32>(voff.p, voff.p, R.d, voff.p + R.d, R.st);
        R.nS = d2h_scalar(voff.p + R.d, R.st);
        R.keys.alloc(R.nS, R.st);
        R.vals.alloc(R.nS, R.st);
    }
    R.D_raw.alloc((R.nE + 1) * sizeof(SymT), R.st);
    if (R.c->flags & GRLGPU_FLAG_KEEP_DICT) R.phr_of.alloc(R.nE, R.st);  // only the test hooks read it
    R.rem.alloc(R.nE, R.st);
    R.einfo.alloc(R.nE, R.st);
    IsSuffix isuf{R.c->is_suffix.p, R.c->sep, R.c->first};
    // metasymbols go to the phrase's table slot, or to the global per-phrase array in multi-GPU rounds
    GRL_LAUNCH("dict_gather", R.nE * (sizeof(CellT) + sizeof(SymT) + 20) + R.nS * 12 + R.d * 32, (dict_gather_kernel<CellT, FIRST, SymT>), grid_for(R.d, 256), 256, 0, R.st, dtext,
               R.ph_pos.p, R.ph_len.p, R.ph_off.p, R.ph_freq.p, R.ph_meta ? (const u32*)nullptr : (const u32*)R.occ_slots.p, R.d, isuf, (SymT*)R.D_raw.p, R.phr_of.p, R.rem.p,
               R.einfo.p, (const u32*)voff.p, c->alphabet + 1, R.sym_bits, R.K, R.spare, R.keys.p, R.vals.p);
}

// ---------------- dictionary stage: suffix order, groups, ranks, pre-BWT, rules, metasymbols ----------------
template <class SymT>
void stage_dict(Round& R) {
    grlgpu_ctx* c = R.c;
    cudaStream_t st = R.st;
    const u64 nE = R.nE, A = c->alphabet;
    const SymT* D = (const SymT*)R.D_raw.p;
    IsSuffix isuf{c->is_suffix.p, c->sep, c->first};

    // -- suffix order: one full sort on the packed first key, then refinement of the unresolved groups only --
    const bool ext_mode = R.ext_mode;
    const u64 nS = R.nS;  // sorted items: the valid entries (key extension) or every entry (prefix doubling)
    const u64 n_words = div_up(nS, 32);
    DevBuf<u32> order_buf, head_bits(n_words, st);
    u64 G = 0;
    {
        DevBuf<u64> keys = std::move(R.keys), keys_alt(nS, st);
        DevBuf<u32> vals = std::move(R.vals), vals_alt(nS, st);
        const int sym_bits = R.sym_bits, K = R.K;
        if (!ext_mode) {
            keys.alloc(nS, st);
            vals.alloc(nS, st);
            GRL_LAUNCH("sfx_first_key", nE * (sizeof(SymT) + 4 + 12), (sfx_first_key_kernel<SymT>), grid_for(nE, 256), 256, 0, st, D, R.rem.p, nE, A + 1, sym_bits, K, keys.p, vals.p);
        }
        u64 *kp = keys.p, *ka = keys_alt.p;
        u32 *vp = vals.p, *va = vals_alt.p;
        radix_sort_pairs(&kp, &vp, &ka, &va, nS, std::min(64, sym_bits * K + R.spare), st);
        if (vp != vals.p) std::swap(vals, vals_alt);
        order_buf = std::move(vals);
        vals_alt.release();
        u32* order_w = order_buf.p;
        u64 nA = 0;
        DevBuf<u32> apos;
        if (ext_mode) {
            {
                DevBuf<u32> flags(nS, st), active_bits(n_words, st);
                GRL_LAUNCH("first_heads", nS * 12, first_heads_kernel, grid_for(nS, 256), 256, 0, st, kp, nS, sym_bits, R.spare, A + 1, flags.p, head_bits.p, active_bits.p);
                keys.release(); keys_alt.release();
                BitmapCompactor ac;
                nA = ac.count(active_bits.p, nS, st);
                apos.alloc(nA, st);
                if (nA) ac.write<u32>(nullptr, apos.p);
            }
            u64 dpt = (u64)K;  // codes already compared
            while (nA > 0) {
                if (dpt > R.max_len + 1) throw Error(GRLGPU_ERR_STATE, "suffix refinement did not converge");
                DevBuf<u64> ak(nA, st), ak_alt(nA, st), nk(nA, st);
                DevBuf<u32> av(nA, st), av_alt(nA, st), ev(nA, st), gflag(nA, st), gexcl(nA, st), flags(nA, st), excl(nA, st), cnt(1, st);
                u64 *akp = ak.p, *aka = ak_alt.p;
                u32 *avp = av.p, *ava = av_alt.p;
                GRL_LAUNCH("ext_keys", nA * 48, (ext_keys_kernel<SymT>), grid_for(nA, 256), 256, 0, st, apos.p, order_w, D, R.rem.p, head_bits.p, nA, dpt, A + 1, sym_bits, K, akp,
                           avp, nk.p, ev.p, gflag.p);
                exclusive_scan<u32, u32>(gflag.p, gexcl.p, nA, cnt.p, st);
                const u64 n_groups = d2h_scalar(cnt.p, st);
                if (getenv("GRLGPU_TRACE")) fprintf(stderr, "[grlgpu] round %d refine: depth %llu, active %llu in %llu groups (of %llu sorted)\n", c->round + 1, dpt, nA, n_groups, nS);
                radix_sort_pairs(&akp, &avp, &aka, &ava, nA, std::min(64, sym_bits * K), st);  // by the extension key ...
                GRL_LAUNCH("ext_gid", nA * 12, ext_gid_kernel, grid_for(nA, 256), 256, 0, st, gflag.p, gexcl.p, nA);
                GRL_LAUNCH("ext_group_keys", nA * 16, ext_group_keys_kernel, grid_for(nA, 256), 256, 0, st, avp, gexcl.p, nA, akp);
                radix_sort_pairs(&akp, &avp, &aka, &ava, nA, std::max(1, bit_width64(n_groups)), st);  // ... then, stably, by group
                GRL_LAUNCH("ext_heads", nA * 24, ext_heads_kernel, grid_for(nA, 256), 256, 0, st, avp, akp, nk.p, nA, flags.p);
                GRL_LAUNCH("ext_writeback", nA * 16, ext_writeback_kernel, grid_for(nA, 256), 256, 0, st, apos.p, avp, ev.p, flags.p, nA, order_w, head_bits.p);
                dpt += (u64)K;
                GRL_LAUNCH("ext_next", nA * 16, ext_next_kernel, grid_for(nA, 256), 256, 0, st, apos.p, avp, ev.p, head_bits.p, R.rem.p, nA, nS, dpt, flags.p);
                exclusive_scan<u32, u32>(flags.p, excl.p, nA, cnt.p, st);
                const u64 nA2 = d2h_scalar(cnt.p, st);
                DevBuf<u32> apos2(nA2, st);
                if (nA2) GRL_LAUNCH("compact_apos", nA * 12, compact_apos_kernel, grid_for(nA, 256), 256, 0, st, flags.p, excl.p, apos.p, nA, apos2.p);
                apos = std::move(apos2);
                nA = nA2;
            }
        } else {
        R.rank.alloc(nE, st);
        {   // heads, position-based ranks and the first active set
            DevBuf<u32> flags(nE, st), excl(nE, st), active_bits(n_words, st), gcount(1, st);
            GRL_LAUNCH("first_heads", nE * 12, first_heads_kernel, grid_for(nE, 256), 256, 0, st, kp, nE, sym_bits, 0, A + 1, flags.p, head_bits.p, active_bits.p);
            keys.release(); keys_alt.release();
            exclusive_scan<u32, u32>(flags.p, excl.p, nE, gcount.p, st);
            const u64 G0 = d2h_scalar(gcount.p, st);
            DevBuf<u32> head_pos(G0, st);
            GRL_LAUNCH("head_pos", nE * 8, head_pos_kernel, grid_for(nE, 256), 256, 0, st, flags.p, excl.p, (const u32*)nullptr, nE, head_pos.p);
            if (nE <= (1ull << 24)) {  // the whole rank array fits the L2: scatter directly
                GRL_LAUNCH("assign_hrank", nE * 20, assign_hrank_kernel, grid_for(nE, 256), 256, 0, st, flags.p, excl.p, head_pos.p, order_w, nE, R.rank.p);
            } else {
                // a random 4-byte scatter over GBs costs ~78 B of DRAM traffic per entry (sector read-modify-write); partition the
                // (entry, rank) pairs by the entry's top byte first, so that every stretch of pairs lands in one 64 MB window
                DevBuf<u32> pk(nE, st), pv(nE, st), pk2(nE, st), pv2(nE, st);
                GRL_LAUNCH("hrank_pairs", nE * 20, hrank_pairs_kernel, grid_for(nE, 256), 256, 0, st, flags.p, excl.p, head_pos.p, order_w, nE, pk.p, pv.p);
                u32 *k1 = pk.p, *v1 = pv.p, *k2 = pk2.p, *v2 = pv2.p;
                radix_partition_u32(&k1, &v1, &k2, &v2, nE, 24, st);
                GRL_LAUNCH("scatter_pairs", nE * 12, scatter_pairs_kernel, grid_for(nE, 256), 256, 0, st, k1, v1, nE, R.rank.p);
            }
            BitmapCompactor ac;
            nA = ac.count(active_bits.p, nE, st);
            apos.alloc(nA, st);
            if (nA) ac.write<u32>(nullptr, apos.p);
        }
        const int rb = bit_width64(nE + 1);
        u64 h = (u64)K;
        while (nA > 0) {  // a key of h codes covers any suffix (<= max_len symbols + terminator) once h > max_len
            if (h > R.max_len + 1) throw Error(GRLGPU_ERR_STATE, "suffix refinement did not converge");
            DevBuf<u64> ak(nA, st), ak_alt(nA, st);
            DevBuf<u32> av(nA, st), av_alt(nA, st), flags(nA, st), excl(nA, st), cnt(1, st);
            u64 *akp = ak.p, *aka = ak_alt.p;
            u32 *avp = av.p, *ava = av_alt.p;
            GRL_LAUNCH("active_keys", nA * 36, active_keys_kernel, grid_for(nA, 256), 256, 0, st, apos.p, order_w, R.rank.p, R.rem.p, nA, h, (u32)(nE + 1), rb, akp, avp);
            radix_sort_pairs(&akp, &avp, &aka, &ava, nA, 2 * rb, st);
            GRL_LAUNCH("key_head_flags", nA * 12, key_head_flags_kernel, grid_for(nA, 256), 256, 0, st, akp, nA, flags.p);
            exclusive_scan<u32, u32>(flags.p, excl.p, nA, cnt.p, st);
            const u64 nH = d2h_scalar(cnt.p, st);
            DevBuf<u32> head_pos(nH, st);
            GRL_LAUNCH("head_pos", nA * 12, head_pos_kernel, grid_for(nA, 256), 256, 0, st, flags.p, excl.p, apos.p, nA, head_pos.p);
            GRL_LAUNCH("assign_hrank", nA * 20, assign_hrank_kernel, grid_for(nA, 256), 256, 0, st, flags.p, excl.p, head_pos.p, avp, nA, R.rank.p);
            GRL_LAUNCH("refine_writeback", nA * 16, refine_writeback_kernel, grid_for(nA, 256), 256, 0, st, apos.p, avp, flags.p, nA, order_w, head_bits.p);
            h *= 2;
            GRL_LAUNCH("active_next", nA * 16, active_next_kernel, grid_for(nA, 256), 256, 0, st, apos.p, avp, head_bits.p, R.rem.p, nA, nE, h, flags.p);
            exclusive_scan<u32, u32>(flags.p, excl.p, nA, cnt.p, st);
            const u64 nA2 = d2h_scalar(cnt.p, st);
            DevBuf<u32> apos2(nA2, st);
            if (nA2) GRL_LAUNCH("compact_apos", nA * 12, compact_apos_kernel, grid_for(nA, 256), 256, 0, st, flags.p, excl.p, apos.p, nA, apos2.p);
            apos = std::move(apos2);
            nA = nA2;
        }
            }  // doubling
    }
    const u32* order = order_buf.p;
    // dense group ids from the head bitmap: per-word prefix counts
    DevBuf<u32> head_pref(n_words, st);
    {
        DevBuf<u32> wc(n_words, st), gtot(1, st);
        GRL_LAUNCH("popc_words", n_words * 8, popc_words_kernel, grid_for(n_words, 256), 256, 0, st, head_bits.p, n_words, wc.p);
        exclusive_scan<u32, u32>(wc.p, head_pref.p, n_words, gtot.p, st);
        G = d2h_scalar(gtot.p, st);
    }
    R.G = G;

    // -- group aggregates --
    DevBuf<u32> gcnt(G, st), grep(G, st), ghead(G, st), rflag(G, st), vflag(G, st), rrank(G, st), vidx(G, st);
    DevBuf<u64> gacc(G, st), gmin(G, st), gmax(G, st), psym(G, st), gfull(G, st);
    gcnt.zero(); gacc.zero(); gmax.zero(); gmin.fill_ff();
    GRL_LAUNCH("group_reduce", nS * 24 + G * 32, group_reduce_kernel, grid_for(nS, 256), 256, 0, st, order, head_bits.p, head_pref.p, R.einfo.p, nS, gcnt.p, gacc.p, gmin.p, gmax.p, grep.p, ghead.p, gfull.p);
    R.einfo.release();
    const u64 bwt_dummy = A + 1, hocc_dummy = A + 2;  // exact_par_phase.hpp:113-115
    GRL_LAUNCH("group_finalize", 0, group_finalize_kernel, grid_for(G, 256), 256, 0, st, gcnt.p, gmin.p, gmax.p, G, bwt_dummy, hocc_dummy, rflag.p, vflag.p, psym.p);
    DevBuf<u32> cnt2(2, st);
    exclusive_scan<u32, u32>(rflag.p, rrank.p, G, cnt2.p, st);
    exclusive_scan<u32, u32>(vflag.p, vidx.p, G, cnt2.p + 1, st);
    u32 hc[2];
    d2h_small(hc, cnt2.p, 8, st);
    const u64 tot = hc[0], nV = hc[1];
    if (tot >= (1ull << 30)) throw Error(GRLGPU_ERR_LIMIT, "more than 2^30 ranks in one round");
    R.tot = tot;

    // -- preliminary BWT: maximal runs over the valid groups --
    c->lvl_sym_bytes = sizeof(SymT);
    {
        DevBuf<u64> csym(nV, st), clen(nV, st);
        GRL_LAUNCH("prebwt_compact", 0, prebwt_compact_kernel, grid_for(G, 256), 256, 0, st, vflag.p, vidx.p, psym.p, gacc.p, G, csym.p, clen.p);
        DevBuf<u32> hflag(nV, st), hexcl(nV, st), nrun(1, st);
        GRL_LAUNCH("key_head_flags", 0, key_head_flags_kernel, grid_for(nV, 256), 256, 0, st, csym.p, nV, hflag.p);
        exclusive_scan<u32, u32>(hflag.p, hexcl.p, nV, nrun.p, st);
        R.n_pre = d2h_scalar(nrun.p, st);
        c->pre_sym.alloc(R.n_pre * sizeof(SymT), st);
        c->pre_len.alloc(R.n_pre, st);
        c->pre_len.zero();
        GRL_LAUNCH("prebwt_runs", 0, (prebwt_runs_kernel<SymT>), grid_for(nV, 256), 256, 0, st, csym.p, clen.p, hflag.p, hexcl.p, nV, (SymT*)c->pre_sym.p, c->pre_len.p);
    }

    // -- metasymbols, next is_suffix, hocc marks, rules --
    DevBuf<u8> is_suffix_next(tot, st);
    is_suffix_next.zero();
    DevBuf<u32> erank(nE, st);
    erank.fill_ff();
    GRL_LAUNCH("full_apply", G * 20 + R.d * 40, full_apply_kernel, grid_for(G, 256), 256, 0, st, gcnt.p, rflag.p, rrank.p, gfull.p, G, (u64)0, R.table.p, R.ph_meta, is_suffix_next.p);
    gfull.release();
    if (ext_mode) {  // hocc marks in sorted order: group info is read sequentially, entries need no rank of their own
        DevBuf<u32> ginfo(G, st);
        GRL_LAUNCH("pack_ginfo_dense", G * 16, pack_ginfo_dense_kernel, grid_for(G, 256), 256, 0, st, gcnt.p, rflag.p, rrank.p, G, ginfo.p);
        GRL_LAUNCH("group_apply", nS * 12, group_apply_kernel, grid_for(nS, 256), 256, 0, st, order, head_bits.p, head_pref.p, ginfo.p, nS, (u64)0, 0u, erank.p);
    } else {
        DevBuf<u32> ginfo(nE, st);  // indexed by head position
        GRL_LAUNCH("pack_ginfo", G * 20, pack_ginfo_kernel, grid_for(G, 256), 256, 0, st, gcnt.p, rflag.p, rrank.p, ghead.p, G, ginfo.p);
        GRL_LAUNCH("entry_finalize", nE * 12, entry_finalize_kernel, grid_for(nE, 256), 256, 0, st, R.rank.p, nE, ginfo.p, erank.p);
    }
    c->rule_l.alloc(tot * sizeof(SymT), st);
    c->rule_r.alloc(tot * sizeof(SymT), st);
    c->has_hocc.alloc(tot, st);
    const u64 alph3 = A + 3, metasym_dummy = alph3 + tot + 1;  // exact_par_phase.cpp:19-20
    GRL_LAUNCH("rules", 0, (rules_kernel<SymT>), grid_for(G, 256), 256, 0, st, gcnt.p, rflag.p, rrank.p, grep.p, G, D, R.rem.p, erank.p, isuf, alph3, metasym_dummy, (SymT*)c->rule_l.p, (SymT*)c->rule_r.p, c->has_hocc.p);
    c->lvl_tot = tot;
    c->lvl_npre = R.n_pre;
    // every valid dictionary entry contributes its phrase's frequency to exactly one run: the lengths of a level sum to at most
    // n + p (phrases overlap by one cell). Multi-GPU: the frequencies are global, the local sizes bound nothing.
    c->lvl_n_in = R.ph_meta ? ~0ull : R.n + R.p;

    if (c->flags & GRLGPU_FLAG_KEEP_DICT) {
        c->kd_d = R.d; c->kd_nE = nE; c->kd_nS = nS;
        c->kd_order.alloc(nS, st);
        GRL_CUDA(cudaMemcpyAsync(c->kd_order.p, order, nS * 4, cudaMemcpyDeviceToDevice, st));
    }
    GRL_CUDA(cudaStreamSynchronize(st));  // temporaries above are released in stream order
    c->is_suffix = std::move(is_suffix_next);
}

template <class OutT>
void stage_rewrite(Round& R, DevBuf<u8>& new_text, DevBuf<u32>& new_end_bits) {
    new_text.alloc(std::max<u64>(R.p * sizeof(OutT), 16), R.st);
    new_end_bits.alloc(div_up(R.p, 32), R.st);
    GRL_LAUNCH("rewrite", R.p * (4 + 16 + sizeof(OutT)), (rewrite_kernel<OutT>), grid_for(R.p, 256), 256, 0, R.st, R.slot_of_phrase.p, R.p, R.table.p, (OutT*)new_text.p, new_end_bits.p);
}

// rewrite the local text with the metasymbols now stored in the local table, report, and make the parse
// the text of the next round. `tot` / `n_pre` / dictionary sizes describe the round's (global) dictionary.
struct RoundTimes { float all = 0, text = 0, dict = 0; };
void finish_round(grlgpu_ctx* c, Round& R, u64 tot, u64 n_pre, u64 dict_d, u64 dict_nE, u64 max_freq, const RoundTimes& tm, Timer* t_all, grlgpu_round_t* out) {
    Timer t_rw(c->st);
    t_rw.start();
    const int bps = bit_width64(tot) + 1;  // exact_par_phase.cpp:456-465
    const int w_out = bps <= 8 ? 1 : bps <= 16 ? 2 : bps <= 32 ? 4 : 8;
    DevBuf<u8> new_text;
    DevBuf<u32> new_end;
    if (w_out == 1) stage_rewrite<u8>(R, new_text, new_end);
    else if (w_out == 2) stage_rewrite<u16>(R, new_text, new_end);
    else if (w_out == 4) stage_rewrite<u32>(R, new_text, new_end);
    else stage_rewrite<u64>(R, new_text, new_end);
    t_rw.stop();
    if (t_all) t_all->stop();

    memset(out, 0, sizeof(*out));
    out->round = (u64)c->round + 1;
    out->n_in = R.n;
    out->n_strings = c->n_strings;
    out->parse_len = R.p;
    out->n_phrases = dict_d;
    out->dict_syms = dict_nE;
    out->max_freq = max_freq;
    out->alphabet = c->alphabet;
    out->tot_phrases = tot;
    out->n_pre_runs = n_pre;
    out->cell_bytes_in = (u32)c->w;
    out->cell_bytes_out = (u32)w_out;
    out->sym_bytes = (u32)c->lvl_sym_bytes;
    out->done = R.p == c->n_strings;
    out->algorithmic_bytes = R.n * (u64)c->w + R.p * (u64)w_out + dict_nE * (u64)c->w + 8 * dict_d;
    out->device_ms = t_all ? t_all->ms() : tm.all;
    out->text_pass_ms = tm.text;
    out->dict_ms = tm.dict;
    out->rewrite_ms = t_rw.ms();

    if ((c->flags & GRLGPU_FLAG_KEEP_DICT) && !R.ph_meta && c->kd_meta.p) {  // metasymbol per distinct phrase, for tests
        std::vector<u32> slots(R.d);
        GRL_CUDA(cudaMemcpy(slots.data(), R.occ_slots.p, R.d * 4, cudaMemcpyDeviceToHost));
        std::vector<ulonglong2> tab(R.cap);
        GRL_CUDA(cudaMemcpy(tab.data(), R.table.p, R.cap * sizeof(ulonglong2), cudaMemcpyDeviceToHost));
        std::vector<u64> metas(R.d);
        for (u64 i = 0; i < R.d; i++) metas[i] = tab[slots[i]].y;
        GRL_CUDA(cudaMemcpy(c->kd_meta.p, metas.data(), R.d * 8, cudaMemcpyHostToDevice));
    }

    // the parse becomes the text of the next round
    GRL_CUDA(cudaStreamSynchronize(c->st));
    c->prof.resolve();
    c->text_own = std::move(new_text);
    c->text = c->text_own.p;
    c->end_bits = std::move(new_end);
    c->n = R.p;
    c->w = w_out;
    c->first = false;
    c->alphabet = tot;
    c->round++;
    c->done = out->done != 0;
}

template <class CellT, bool FIRST>
void run_round_t(grlgpu_ctx* c, grlgpu_round_t* out) {
    Round R(c);
    Timer t_all(c->st), t_text(c->st), t_dict(c->st);
    t_all.start();
    t_text.start();
    stage_flags<CellT, FIRST>(R);
    stage_dedup<CellT>(R);
    t_text.stop();
    t_dict.start();
    const u64 A = c->alphabet;
    // rule values go up to A + 3 + tot + 1 with tot <= nE
    const bool wide = (A + R.nE + 8) >= (1ull << 32);
    if (wide) { stage_gather<CellT, FIRST, u64>(R); stage_dict<u64>(R); }
    else { stage_gather<CellT, FIRST, u32>(R); stage_dict<u32>(R); }
    t_dict.stop();
    if (c->flags & GRLGPU_FLAG_KEEP_DICT) {
        c->kd_D = std::move(R.D_raw);
        c->kd_off = std::move(R.ph_off);
        c->kd_len = std::move(R.ph_len);
        c->kd_freq = std::move(R.ph_freq);
        c->kd_phr_of = std::move(R.phr_of);
        c->kd_meta.alloc(R.d, c->st);
    }
    RoundTimes tm;
    tm.text = t_text.ms();
    tm.dict = t_dict.ms();
    finish_round(c, R, R.tot, R.n_pre, R.d, R.nE, R.max_freq, tm, &t_all, out);
}

// ---------------- multi-GPU round: the caller owns the exchange between the calls ----------------
struct MgRound {
    Round R;                        // this rank's shard
    DevBuf<u32> perm;               // local distinct phrases ordered by owner
    DevBuf<u64> offs;               // cell offset of each packed phrase (+ total)
    // owner side: dedup of what the other ranks sent
    DevBuf<ulonglong2> ptable;
    DevBuf<u32> pslots, p_len;
    DevBuf<u64> p_pos, p_freq, p_offs;
    DevBuf<u32> recv_dense;         // partition-local index of every phrase this rank received as an owner (for the metasymbol return)
    u64 m_recv = 0;
    u64 d_part = 0, cells_part = 0;
    const void* recv_cells = nullptr;
    Timer t_text, t_dict;
    float ms_text = 0;
    // distributed ranking: this rank's slice of the suffix order of the GLOBAL dictionary and its group results
    std::unique_ptr<Round> GR;      // the global dictionary
    DevBuf<u64> g_meta;
    const void* g_cells = nullptr;
    u64 nL = 0, G = 0, tot_local = 0, n_pre_local = 0;
    int sym_bytes = 4;
    DevBuf<u32> order, head_bits, head_pref, gcnt, grep, rflag, rrank;
    DevBuf<u64> gfull;
    DevBuf<u8> sl_rule_l, sl_rule_r, sl_has_hocc, sl_pre_sym;  // this rank's slice of the level artefacts
    DevBuf<u64> sl_pre_len;
    explicit MgRound(grlgpu_ctx* c) : R(c), t_text(c->st), t_dict(c->st) {}
};

template <class CellT, bool FIRST>
void mg_local_t(grlgpu_ctx* c, int G, grlgpu_part_t* per_owner, u64* parse_len_local) {
    delete c->mg;
    c->mg = nullptr;
    c->mg = new MgRound(c);
    c->mg_ranks = G;
    MgRound& M = *c->mg;
    Round& R = M.R;
    M.t_text.start();
    stage_flags<CellT, FIRST>(R);
    stage_dedup<CellT>(R);
    // owner of every local distinct phrase = content hash % G; order the phrases by owner
    DevBuf<u64> keys(R.d, R.st), keys_alt(R.d, R.st);
    DevBuf<u32> vals(R.d, R.st), vals_alt(R.d, R.st);
    GRL_LAUNCH("phrase_owner", 0, (phrase_owner_kernel<CellT>), grid_for(R.d, 256), 256, 0, R.st, (const CellT*)c->text, R.ph_pos.p, R.ph_len.p, R.d, (u32)G, keys.p, vals.p);
    u64 *kp = keys.p, *ka = keys_alt.p;
    u32 *vp = vals.p, *va = vals_alt.p;
    radix_sort_pairs(&kp, &vp, &ka, &va, R.d, std::max(1, bit_width64((u64)G - 1)), R.st);
    if (vp != vals.p) std::swap(vals, vals_alt);
    M.perm = std::move(vals);
    DevBuf<u32> lens_sorted(R.d, R.st);
    GRL_LAUNCH("gather_u32", 0, gather_u32_kernel, grid_for(R.d, 256), 256, 0, R.st, R.ph_len.p, M.perm.p, R.d, lens_sorted.p);
    M.offs.alloc(R.d + 1, R.st);
    exclusive_scan<u32, u64>(lens_sorted.p, M.offs.p, R.d, M.offs.p + R.d, R.st);
    DevBuf<u64> first(G + 1, R.st);
    GRL_LAUNCH("owner_bounds", 0, owner_bounds_kernel, 1, 32, 0, R.st, kp, R.d, (u32)G, first.p);
    std::vector<u64> hf(G + 1), ho(G + 1);
    GRL_CUDA(cudaMemcpyAsync(hf.data(), first.p, (G + 1) * 8, cudaMemcpyDeviceToHost, R.st));
    GRL_CUDA(cudaStreamSynchronize(R.st));
    for (int g = 0; g <= G; g++) ho[g] = d2h_scalar(M.offs.p + hf[g], R.st);
    for (int g = 0; g < G; g++) { per_owner[g].n_phrases = hf[g + 1] - hf[g]; per_owner[g].n_cells = ho[g + 1] - ho[g]; }
    *parse_len_local = R.p;
    M.t_text.stop();
    M.ms_text = M.t_text.ms();
}

template <class CellT>
void mg_pack_t(grlgpu_ctx* c, u32* d_lens, u64* d_counts, void* d_cells) {
    MgRound& M = *c->mg;
    Round& R = M.R;
    GRL_LAUNCH("pack_phrases", 0, (pack_phrases_kernel<CellT>), grid_for(R.d, 256), 256, 0, R.st, (const CellT*)c->text, R.ph_pos.p, R.ph_len.p, R.ph_freq.p, M.perm.p, M.offs.p, R.d, d_lens, d_counts, (CellT*)d_cells);
    GRL_CUDA(cudaStreamSynchronize(R.st));
}

template <class CellT>
void mg_merge_t(grlgpu_ctx* c, const u32* lens, const u64* counts, const void* cells, u64 m, u64 n_cells, grlgpu_part_t* part) {
    MgRound& M = *c->mg;
    cudaStream_t st = c->st;
    M.recv_cells = cells;
    DevBuf<u64> offs(m + 1, st);
    exclusive_scan<u32, u64>(lens, offs.p, m, offs.p + m, st);
    {   // the pack tables cannot tell apart lengths >= 2^24-1 (no bitmap to consult)
        DevBuf<u64> len64(m, st), mx(1, st);
        mx.zero();
        GRL_LAUNCH("u32_to_u64", 0, u32_to_u64_kernel, grid_for(m, 256), 256, 0, st, lens, m, len64.p);
        GRL_LAUNCH("reduce_max_u64", 0, reduce_max_u64_kernel, 296, 256, 0, st, len64.p, m, mx.p);
        if (d2h_scalar(mx.p, st) >= HT_LEN_SAT) throw Error(GRLGPU_ERR_LIMIT, "multi-GPU rounds support phrases shorter than 2^24-1 cells");
        if (d2h_scalar(offs.p + m, st) != n_cells) throw Error(GRLGPU_ERR_ARG, "received cell count does not match the received lengths");
    }
    const u64 cap = std::max<u64>(1024, (m + m / 2 + m / 10 + 255) / 256 * 256);
    if (cap > (1ull << 31) - 256) throw Error(GRLGPU_ERR_LIMIT, "phrase table would exceed 2^31 slots");
    M.ptable.alloc(cap, st);
    GRL_LAUNCH("table_init", cap * 16, table_init_kernel, grid_for(cap, 256), 256, 0, st, M.ptable.p, cap);
    DevBuf<u32> overflow(1, st);
    overflow.zero();
    DevBuf<u32> recv_slot(m, st);
    GRL_LAUNCH("pack_insert", 0, (pack_insert_kernel<CellT>), grid_for(m, 256), 256, 0, st, (const CellT*)cells, offs.p, lens, counts, m, M.ptable.p, cap, overflow.p, recv_slot.p);
    if (d2h_scalar(overflow.p, st)) throw Error(GRLGPU_ERR_STATE, "partition table overflow");
    DevBuf<u32> occ_bits(cap / 32, st);
    GRL_LAUNCH("table_occupancy", cap * 16, table_occupancy_kernel, (unsigned)(cap / 256), 256, 0, st, M.ptable.p, cap, occ_bits.p);
    BitmapCompactor oc;
    M.d_part = oc.count(occ_bits.p, cap, st);
    M.pslots.alloc(M.d_part, st);
    oc.write<u32>(nullptr, M.pslots.p);
    M.p_pos.alloc(M.d_part, st);
    M.p_len.alloc(M.d_part, st);
    M.p_freq.alloc(M.d_part, st);
    GRL_LAUNCH("dict_meta", 0, dict_meta_kernel, grid_for(M.d_part, 256), 256, 0, st, M.ptable.p, M.pslots.p, M.d_part, (const u32*)nullptr, (const u32*)nullptr, (u64)0, M.p_pos.p, M.p_len.p, M.p_freq.p);
    M.p_offs.alloc(M.d_part + 1, st);
    exclusive_scan<u32, u64>(M.p_len.p, M.p_offs.p, M.d_part, M.p_offs.p + M.d_part, st);
    M.cells_part = d2h_scalar(M.p_offs.p + M.d_part, st);
    // the frequencies have been read: the count field now holds the slot's dense index, which every received phrase inherits
    M.m_recv = m;
    M.recv_dense.alloc(m, st);
    GRL_LAUNCH("slot_dense", 0, slot_dense_kernel, grid_for(M.d_part, 256), 256, 0, st, M.pslots.p, M.d_part, M.ptable.p);
    GRL_LAUNCH("recv_dense", 0, recv_dense_kernel, grid_for(m, 256), 256, 0, st, recv_slot.p, m, M.ptable.p, M.recv_dense.p);
    part->n_phrases = M.d_part;
    part->n_cells = M.cells_part;
}

template <class CellT>
void mg_pack_part_t(grlgpu_ctx* c, u32* d_lens, u64* d_freqs, void* d_cells) {
    MgRound& M = *c->mg;
    GRL_LAUNCH("pack_phrases", 0, (pack_phrases_kernel<CellT>), grid_for(M.d_part, 256), 256, 0, c->st, (const CellT*)M.recv_cells, M.p_pos.p, M.p_len.p, M.p_freq.p, (const u32*)nullptr, M.p_offs.p, M.d_part, d_lens, d_freqs, (CellT*)d_cells);
    GRL_CUDA(cudaStreamSynchronize(c->st));
    M.ptable.release(); M.pslots.release(); M.p_pos.release(); M.p_len.release(); M.p_freq.release(); M.p_offs.release();
}

template <class CellT, bool FIRST>
void mg_global_t(grlgpu_ctx* c, const u32* lens, const u64* freqs, const void* cells, u64 d, u64 n_cells, int done_global, grlgpu_round_t* out);

// global dictionary as a Round: "text" = the gathered cells (shared by the replicated and the distributed ranking)
void mg_setup_global(grlgpu_ctx* c, Round& GR, const u32* lens, const u64* freqs, const void* cells, u64 d, u64 n_cells) {
    cudaStream_t st = c->st;
    GR.dict_text = cells;
    GR.d = d;
    GR.ph_len.alloc(d, st);
    GR.ph_freq.alloc(d, st);
    GR.ph_pos.alloc(d + 1, st);
    GRL_CUDA(cudaMemcpyAsync(GR.ph_len.p, lens, d * 4, cudaMemcpyDeviceToDevice, st));
    GRL_CUDA(cudaMemcpyAsync(GR.ph_freq.p, freqs, d * 8, cudaMemcpyDeviceToDevice, st));
    exclusive_scan<u32, u64>(GR.ph_len.p, GR.ph_pos.p, d, GR.ph_pos.p + d, st);
    if (d2h_scalar(GR.ph_pos.p + d, st) != n_cells) throw Error(GRLGPU_ERR_ARG, "gathered cell count does not match the gathered lengths");
    dict_offsets(GR);
}

// content -> global phrase index table, metasymbol of every local distinct phrase, rewrite of the shard
template <class CellT>
void mg_map_and_rewrite(grlgpu_ctx* c, MgRound& M, Round& GR, const void* cells, const u64* g_meta, u64 tot, u64 n_pre, int done_global, grlgpu_round_t* out,
                        const u64* local_meta = nullptr) {
    Round& R = M.R;
    cudaStream_t st = c->st;
    const u64 d = GR.d;
    if (local_meta) {  // the owners returned the metasymbols in pack order: no global table, no content lookups
        GRL_LAUNCH("apply_reply", R.d * 24, apply_reply_kernel, grid_for(R.d, 256), 256, 0, st, M.perm.p, R.occ_slots.p, local_meta, R.d, R.table.p);
        M.t_dict.stop();
        RoundTimes tm;
        tm.text = M.ms_text;
        tm.dict = M.t_dict.ms();
        tm.all = tm.text + tm.dict;
        finish_round(c, R, tot, n_pre, GR.d, GR.nE, GR.max_freq, tm, nullptr, out);
        out->done = done_global ? 1u : 0u;
        c->done = done_global != 0;
        return;
    }
    const u64 gcap = std::max<u64>(1024, (d + d / 2 + d / 10 + 255) / 256 * 256);
    if (gcap > (1ull << 31) - 256) throw Error(GRLGPU_ERR_LIMIT, "phrase table would exceed 2^31 slots");
    DevBuf<ulonglong2> gtable(gcap, st);
    GRL_LAUNCH("table_init", gcap * 16, table_init_kernel, grid_for(gcap, 256), 256, 0, st, gtable.p, gcap);
    DevBuf<u32> flag(2, st);
    flag.zero();
    GRL_LAUNCH("pack_insert", 0, (pack_insert_kernel<CellT>), grid_for(d, 256), 256, 0, st, (const CellT*)cells, GR.ph_pos.p, GR.ph_len.p, (const u64*)nullptr, d, gtable.p, gcap, flag.p, (u32*)nullptr);
    GRL_LAUNCH("map_local", 0, (map_local_kernel<CellT>), grid_for(R.d, 256), 256, 0, st, (const CellT*)c->text, R.ph_pos.p, R.ph_len.p, R.occ_slots.p, R.d, (const CellT*)cells, gtable.p, gcap, g_meta, R.table.p, flag.p + 1);
    u32 hflag[2];
    d2h_small(hflag, flag.p, 8, st);
    if (hflag[0]) throw Error(GRLGPU_ERR_STATE, "global table overflow");
    if (hflag[1]) throw Error(GRLGPU_ERR_STATE, "a local phrase is missing from the global dictionary");
    M.t_dict.stop();
    RoundTimes tm;
    tm.text = M.ms_text;
    tm.dict = M.t_dict.ms();
    tm.all = tm.text + tm.dict;
    finish_round(c, R, tot, n_pre, GR.d, GR.nE, GR.max_freq, tm, nullptr, out);
    out->done = done_global ? 1u : 0u;  // the phase ends when EVERY rank's strings are single cells
    c->done = done_global != 0;
}

// Distributed ranking, step 1: this rank sorts and groups the suffix entries whose first key falls in its range.
// info[0] = 1 if the distributed path was taken (0: nothing was done, call grlgpu_mg_global), info[1] = ranked groups
// of this rank, info[2] = preliminary-BWT runs of this rank, info[3] = dictionary entries nE, info[4] = symbol bytes.
template <class CellT, bool FIRST, class SymT>
void mg_rank_sort_sym(grlgpu_ctx* c, const u32* lens, const u64* freqs, const void* cells, u64 d, u64 n_cells, int rank_id, int n_ranks, u64* info) {
    MgRound& M = *c->mg;
    cudaStream_t st = c->st;
    Round& GR = *M.GR;
    const u64 nE = GR.nE, A = c->alphabet;
    stage_gather<CellT, FIRST, SymT>(GR, true);  // first keys + entry ids of the valid entries come with it
    const SymT* D = (const SymT*)GR.D_raw.p;
    const int sym_bits = GR.sym_bits, K = GR.K;
    const int key_bits = std::min(64, sym_bits * K), first_bits = std::min(64, sym_bits * K + GR.spare);
    const u64 nS = GR.nS;
    M.sym_bytes = sizeof(SymT);
    // splitters from a regular sample of the first keys, identical on every rank
    DevBuf<u64> keys = std::move(GR.keys);
    DevBuf<u32> ids = std::move(GR.vals);
    u64 lo = 0, hi = 0;
    int hi_open = 1;
    if (nS) {
        const u64 ns = std::min<u64>(nS, 1ull << 16), stride = std::max<u64>(1, nS / ns);
        DevBuf<u64> sk(ns, st), sk2(ns, st);
        DevBuf<u32> sv(ns, st), sv2(ns, st);
        GRL_LAUNCH("key_sample", 0, key_sample_kernel, grid_for(ns, 256), 256, 0, st, keys.p, nS, stride, ns, sk.p, sv.p);
        u64 *a = sk.p, *b = sk2.p;
        u32 *av = sv.p, *bv = sv2.p;
        radix_sort_pairs(&a, &av, &b, &bv, ns, first_bits, st);
        std::vector<u64> hs(ns);
        GRL_CUDA(cudaMemcpyAsync(hs.data(), a, ns * 8, cudaMemcpyDeviceToHost, st));
        GRL_CUDA(cudaStreamSynchronize(st));
        auto splitter = [&](int r) { return hs[(size_t)((u64)r * ns / (u64)n_ranks)]; };
        lo = rank_id == 0 ? 0 : splitter(rank_id);
        if (rank_id + 1 < n_ranks) { hi = splitter(rank_id + 1); hi_open = 0; }
    }
    // my entries
    DevBuf<u64> mk, mk_alt;
    DevBuf<u32> mv, mv_alt;
    u64 nL = 0;
    {
        DevBuf<u32> flags(nS, st), excl(nS, st), cnt(1, st);
        GRL_LAUNCH("key_range_flags", nS * 12, key_range_flags_kernel, grid_for(nS, 256), 256, 0, st, keys.p, nS, lo, hi, hi_open, flags.p);
        exclusive_scan<u32, u32>(flags.p, excl.p, nS, cnt.p, st);
        nL = nS ? d2h_scalar(cnt.p, st) : 0;
        mk.alloc(nL, st); mk_alt.alloc(nL, st); mv.alloc(nL, st); mv_alt.alloc(nL, st);
        GRL_LAUNCH("key_range_compact", nS * 20, key_range_compact_kernel, grid_for(nS, 256), 256, 0, st, keys.p, ids.p, flags.p, excl.p, nS, mk.p, mv.p);
    }
    keys.release(); ids.release();
    M.nL = nL;
    const u64 n_words = div_up(std::max<u64>(nL, 1), 32);
    M.head_bits.alloc(n_words, st);
    M.head_bits.zero();
    u64 *kp = mk.p, *ka = mk_alt.p;
    u32 *vp = mv.p, *va = mv_alt.p;
    radix_sort_pairs(&kp, &vp, &ka, &va, nL, first_bits, st);
    if (vp != mv.p) std::swap(mv, mv_alt);
    M.order = std::move(mv);
    mv_alt.release();
    u32* order_w = M.order.p;
    u64 nA = 0;
    DevBuf<u32> apos;
    if (nL) {
        DevBuf<u32> flags(nL, st), active_bits(n_words, st);
        GRL_LAUNCH("first_heads", nL * 12, first_heads_kernel, grid_for(nL, 256), 256, 0, st, kp, nL, sym_bits, GR.spare, A + 1, flags.p, M.head_bits.p, active_bits.p);
        BitmapCompactor ac;
        nA = ac.count(active_bits.p, nL, st);
        apos.alloc(nA, st);
        if (nA) ac.write<u32>(nullptr, apos.p);
    }
    mk.release(); mk_alt.release();
    u64 dpt = (u64)K;
    while (nA > 0) {  // refinement by key extension: local to the dictionary text, no exchange needed
        if (dpt > GR.max_len + 1) throw Error(GRLGPU_ERR_STATE, "suffix refinement did not converge");
        DevBuf<u64> ak(nA, st), ak_alt(nA, st), nk(nA, st);
        DevBuf<u32> av(nA, st), av_alt(nA, st), ev(nA, st), gflag(nA, st), gexcl(nA, st), flags(nA, st), excl(nA, st), cnt(1, st);
        u64 *akp = ak.p, *aka = ak_alt.p;
        u32 *avp = av.p, *ava = av_alt.p;
        GRL_LAUNCH("ext_keys", nA * 48, (ext_keys_kernel<SymT>), grid_for(nA, 256), 256, 0, st, apos.p, order_w, D, GR.rem.p, M.head_bits.p, nA, dpt, A + 1, sym_bits, K, akp, avp,
                   nk.p, ev.p, gflag.p);
        exclusive_scan<u32, u32>(gflag.p, gexcl.p, nA, cnt.p, st);
        const u64 n_groups = d2h_scalar(cnt.p, st);
        radix_sort_pairs(&akp, &avp, &aka, &ava, nA, key_bits, st);
        GRL_LAUNCH("ext_gid", nA * 12, ext_gid_kernel, grid_for(nA, 256), 256, 0, st, gflag.p, gexcl.p, nA);
        GRL_LAUNCH("ext_group_keys", nA * 16, ext_group_keys_kernel, grid_for(nA, 256), 256, 0, st, avp, gexcl.p, nA, akp);
        radix_sort_pairs(&akp, &avp, &aka, &ava, nA, std::max(1, bit_width64(n_groups)), st);
        GRL_LAUNCH("ext_heads", nA * 24, ext_heads_kernel, grid_for(nA, 256), 256, 0, st, avp, akp, nk.p, nA, flags.p);
        GRL_LAUNCH("ext_writeback", nA * 16, ext_writeback_kernel, grid_for(nA, 256), 256, 0, st, apos.p, avp, ev.p, flags.p, nA, order_w, M.head_bits.p);
        dpt += (u64)K;
        GRL_LAUNCH("ext_next", nA * 16, ext_next_kernel, grid_for(nA, 256), 256, 0, st, apos.p, avp, ev.p, M.head_bits.p, GR.rem.p, nA, nL, dpt, flags.p);
        exclusive_scan<u32, u32>(flags.p, excl.p, nA, cnt.p, st);
        const u64 nA2 = d2h_scalar(cnt.p, st);
        DevBuf<u32> apos2(nA2, st);
        if (nA2) GRL_LAUNCH("compact_apos", nA * 12, compact_apos_kernel, grid_for(nA, 256), 256, 0, st, flags.p, excl.p, apos.p, nA, apos2.p);
        apos = std::move(apos2);
        nA = nA2;
    }
    // groups of my slice
    M.head_pref.alloc(n_words, st);
    u64 G = 0;
    {
        DevBuf<u32> wc(n_words, st), gtot(1, st);
        GRL_LAUNCH("popc_words", n_words * 8, popc_words_kernel, grid_for(n_words, 256), 256, 0, st, M.head_bits.p, n_words, wc.p);
        exclusive_scan<u32, u32>(wc.p, M.head_pref.p, n_words, gtot.p, st);
        G = nL ? d2h_scalar(gtot.p, st) : 0;
    }
    M.G = G;
    M.gcnt.alloc(G, st); M.grep.alloc(G, st); M.rflag.alloc(G, st); M.rrank.alloc(G, st);
    M.gfull.alloc(G, st);
    DevBuf<u32> ghead(G, st), vflag(G, st), vidx(G, st);
    DevBuf<u64> gacc(G, st), gmin(G, st), gmax(G, st), psym(G, st);
    M.gcnt.zero(); gacc.zero(); gmax.zero(); gmin.fill_ff();
    GRL_LAUNCH("group_reduce", nL * 24 + G * 32, group_reduce_kernel, grid_for(nL, 256), 256, 0, st, M.order.p, M.head_bits.p, M.head_pref.p, GR.einfo.p, nL, M.gcnt.p, gacc.p, gmin.p,
               gmax.p, M.grep.p, ghead.p, M.gfull.p);
    GR.einfo.release();
    const u64 bwt_dummy = A + 1, hocc_dummy = A + 2;
    GRL_LAUNCH("group_finalize", 0, group_finalize_kernel, grid_for(G, 256), 256, 0, st, M.gcnt.p, gmin.p, gmax.p, G, bwt_dummy, hocc_dummy, M.rflag.p, vflag.p, psym.p);
    DevBuf<u32> cnt2(2, st);
    exclusive_scan<u32, u32>(M.rflag.p, M.rrank.p, G, cnt2.p, st);
    exclusive_scan<u32, u32>(vflag.p, vidx.p, G, cnt2.p + 1, st);
    u32 hc[2];
    d2h_small(hc, cnt2.p, 8, st);
    M.tot_local = hc[0];
    const u64 nV = hc[1];
    {   // preliminary BWT of my slice: maximal runs over my valid groups (the caller merges across rank boundaries)
        DevBuf<u64> csym(nV, st), clen(nV, st);
        GRL_LAUNCH("prebwt_compact", 0, prebwt_compact_kernel, grid_for(G, 256), 256, 0, st, vflag.p, vidx.p, psym.p, gacc.p, G, csym.p, clen.p);
        DevBuf<u32> hflag(nV, st), hexcl(nV, st), nrun(1, st);
        GRL_LAUNCH("key_head_flags", 0, key_head_flags_kernel, grid_for(nV, 256), 256, 0, st, csym.p, nV, hflag.p);
        exclusive_scan<u32, u32>(hflag.p, hexcl.p, nV, nrun.p, st);
        M.n_pre_local = nV ? d2h_scalar(nrun.p, st) : 0;
        M.sl_pre_sym.alloc(M.n_pre_local * sizeof(SymT), st);
        M.sl_pre_len.alloc(M.n_pre_local, st);
        M.sl_pre_len.zero();
        GRL_LAUNCH("prebwt_runs", 0, (prebwt_runs_kernel<SymT>), grid_for(nV, 256), 256, 0, st, csym.p, clen.p, hflag.p, hexcl.p, nV, (SymT*)M.sl_pre_sym.p, M.sl_pre_len.p);
    }
    GRL_CUDA(cudaStreamSynchronize(st));
    info[0] = 1; info[1] = M.tot_local; info[2] = M.n_pre_local; info[3] = nE; info[4] = sizeof(SymT);
}

template <class CellT, bool FIRST>
void mg_rank_sort_t(grlgpu_ctx* c, const u32* lens, const u64* freqs, const void* cells, u64 d, u64 n_cells, int rank_id, int n_ranks, u64* info) {
    MgRound& M = *c->mg;
    M.t_dict.start();
    M.GR.reset(new Round(c));
    Round& GR = *M.GR;
    mg_setup_global(c, GR, lens, freqs, cells, d, n_cells);
    M.g_cells = cells;
    const int sym_bits = bit_width64(c->alphabet + 1);
    const u64 K = (u64)std::max(1, 64 / sym_bits);
    const bool ext_ok = (GR.max_len + 1 + K - 1) / K <= 64;
    const bool big = GR.nE >= (1ull << 22) || (c->flags & GRLGPU_FLAG_FORCE_DIST_RANK);
    info[0] = 0; info[1] = info[2] = 0; info[3] = GR.nE; info[4] = 4;
    if (!ext_ok || !big || n_ranks < 2) { M.GR.reset(); return; }  // small or long-phrase dictionaries: replicated ranking
    const bool wide = (c->alphabet + GR.nE + 8) >= (1ull << 32);
    if (wide) mg_rank_sort_sym<CellT, FIRST, u64>(c, lens, freqs, cells, d, n_cells, rank_id, n_ranks, info);
    else mg_rank_sort_sym<CellT, FIRST, u32>(c, lens, freqs, cells, d, n_cells, rank_id, n_ranks, info);
}

// step 2: with the global rank offset of this rank known, write my share of the global per-phrase metasymbols, of the
// next round's is_suffix and of the hocc marks (rank + 1) into caller-owned, zero-initialised device arrays
template <class SymT>
void mg_rank_apply_sym(grlgpu_ctx* c, u64 rank_base, u64* g_meta, u8* is_suffix_next, u32* erank1) {
    MgRound& M = *c->mg;
    Round& GR = *M.GR;
    cudaStream_t st = c->st;
    DevBuf<u32> ginfo(M.G, st);
    GRL_LAUNCH("full_apply", M.G * 20, full_apply_kernel, grid_for(M.G, 256), 256, 0, st, M.gcnt.p, M.rflag.p, M.rrank.p, M.gfull.p, M.G, rank_base, (ulonglong2*)nullptr, g_meta, is_suffix_next);
    M.gfull.release();
    GRL_LAUNCH("pack_ginfo_dense", M.G * 16, pack_ginfo_dense_kernel, grid_for(M.G, 256), 256, 0, st, M.gcnt.p, M.rflag.p, M.rrank.p, M.G, ginfo.p);
    GRL_LAUNCH("group_apply", M.nL * 12, group_apply_kernel, grid_for(M.nL, 256), 256, 0, st, M.order.p, M.head_bits.p, M.head_pref.p, ginfo.p, M.nL, rank_base, 1u, erank1);
    GRL_CUDA(cudaStreamSynchronize(st));
}

// step 3 (after the caller all-reduced the three arrays with MAX): rules of my ranked groups, local rewrite
template <class CellT, class SymT>
void mg_rank_finish_sym(grlgpu_ctx* c, u64 rank_base, u64 tot, u64 n_pre_global, const u64* g_meta, const u8* is_suffix_next, u32* erank1, int done_global, grlgpu_round_t* out,
                        const u64* local_meta) {
    MgRound& M = *c->mg;
    Round& GR = *M.GR;
    cudaStream_t st = c->st;
    const u64 A = c->alphabet;
    IsSuffix isuf{c->is_suffix.p, c->sep, c->first};
    GRL_LAUNCH("erank_decode", GR.nE * 8, erank_decode_kernel, grid_for(GR.nE, 256), 256, 0, st, erank1, GR.nE);
    M.sl_rule_l.alloc(M.tot_local * sizeof(SymT), st);
    M.sl_rule_r.alloc(M.tot_local * sizeof(SymT), st);
    M.sl_has_hocc.alloc(M.tot_local, st);
    const u64 alph3 = A + 3, metasym_dummy = alph3 + tot + 1;
    GRL_LAUNCH("rules", 0, (rules_kernel<SymT>), grid_for(M.G, 256), 256, 0, st, M.gcnt.p, M.rflag.p, M.rrank.p, M.grep.p, M.G, (const SymT*)GR.D_raw.p, GR.rem.p, erank1, isuf, alph3,
               metasym_dummy, (SymT*)M.sl_rule_l.p, (SymT*)M.sl_rule_r.p, M.sl_has_hocc.p);
    (void)rank_base;
    // the next round's is_suffix is global
    DevBuf<u8> isn(tot, st);
    GRL_CUDA(cudaMemcpyAsync(isn.p, is_suffix_next, tot, cudaMemcpyDeviceToDevice, st));
    c->lvl_sym_bytes = sizeof(SymT);
    mg_map_and_rewrite<CellT>(c, M, GR, M.g_cells, g_meta, tot, n_pre_global, done_global, out, local_meta);
    c->is_suffix = std::move(isn);
    c->rule_l.release(); c->rule_r.release(); c->has_hocc.release(); c->pre_sym.release(); c->pre_len.release();  // level artefacts live in slices this round
    c->lvl_tot = 0; c->lvl_npre = 0; c->lvl_n_in = ~0ull;
    // keep only the slices (until grlgpu_mg_level_slice / the next round)
    M.GR.reset();
    M.order.release(); M.head_bits.release(); M.head_pref.release(); M.gfull.release();
    M.gcnt.release(); M.grep.release(); M.rflag.release(); M.rrank.release();
}

template <class CellT, bool FIRST>
void mg_global_t(grlgpu_ctx* c, const u32* lens, const u64* freqs, const void* cells, u64 d, u64 n_cells, int done_global, grlgpu_round_t* out) {
    MgRound& M = *c->mg;
    cudaStream_t st = c->st;
    if (!M.GR) M.t_dict.start();
    // the global dictionary (identical on every rank), ranked here in full (replicated ranking)
    Round GR(c);
    mg_setup_global(c, GR, lens, freqs, cells, d, n_cells);
    DevBuf<u64> g_meta(d, st);
    GR.ph_meta = g_meta.p;
    const bool wide = (c->alphabet + GR.nE + 8) >= (1ull << 32);
    if (wide) { stage_gather<CellT, FIRST, u64>(GR); stage_dict<u64>(GR); }
    else { stage_gather<CellT, FIRST, u32>(GR); stage_dict<u32>(GR); }
    mg_map_and_rewrite<CellT>(c, M, GR, cells, g_meta.p, GR.tot, GR.n_pre, done_global, out);
    delete c->mg;
    c->mg = nullptr;
}

void run_round(grlgpu_ctx* c, grlgpu_round_t* out) {
    if (c->first) {
        switch (c->w) {
            case 1: run_round_t<u8, true>(c, out); break;
            case 2: run_round_t<u16, true>(c, out); break;
            case 4: run_round_t<u32, true>(c, out); break;
            default: run_round_t<u64, true>(c, out); break;
        }
    } else {
        switch (c->w) {
            case 1: run_round_t<u8, false>(c, out); break;
            case 2: run_round_t<u16, false>(c, out); break;
            case 4: run_round_t<u32, false>(c, out); break;
            default: run_round_t<u64, false>(c, out); break;
        }
    }
}

template <class CellT>
void compute_stats(grlgpu_ctx* c) {
    const CellT* text = (const CellT*)c->text;
    CellT sep_c;
    GRL_CUDA(cudaMemcpyAsync(&sep_c, text + (c->n - 1), sizeof(CellT), cudaMemcpyDeviceToHost, c->st));
    GRL_CUDA(cudaStreamSynchronize(c->st));
    DevBuf<StatsAcc> acc(1, c->st);
    StatsAcc h;
    memset(&h, 0, sizeof(h));
    h.min_sym = ~0ULL;
    GRL_CUDA(cudaMemcpyAsync(acc.p, &h, sizeof(h), cudaMemcpyHostToDevice, c->st));
    GRL_LAUNCH("stats", 0, (stats_kernel<CellT>), 148 * 8, 256, 0, c->st, text, c->n, sep_c, acc.p);
    GRL_CUDA(cudaMemcpyAsync(&h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, c->st));
    GRL_CUDA(cudaStreamSynchronize(c->st));
    grlgpu_stats_t& s = c->stats;
    s.n_syms = c->n;
    s.sep_sym = (u64)sep_c;
    s.n_strings = h.n_sep;
    s.max_sym_freq = c->n;  // utils.cpp:117
    if (sizeof(CellT) == 1) {  // utils.cpp:161-175
        memcpy(c->hist, h.hist, sizeof(c->hist));
        int lo = 0, hi = 255;
        while (h.hist[lo] == 0) lo++;
        while (h.hist[hi] == 0) hi--;
        s.min_sym = lo; s.max_sym = hi;
        u64 m = 0;
        for (int i = 0; i < 256; i++) m = std::max<u64>(m, h.hist[i]);
        s.max_sym_freq = m;
    } else { s.min_sym = h.min_sym; s.max_sym = h.max_sym; }
    if (s.sep_sym != s.min_sym) throw Error(GRLGPU_ERR_ILL_FORMED, "the collection is ill formed: the last symbol is not the smallest symbol");
    // longest string: positions of the separators -> adjacent differences
    {
        const u64 n_words = div_up(c->n, 32);
        DevBuf<u32> eb(n_words, c->st), sb(n_words, c->st);
        GRL_LAUNCH("first_round_end_bits", c->n * sizeof(CellT), (first_round_end_bits_kernel<CellT>), grid_for(n_words, 256), 256, 0, c->st, text, c->n, sep_c, eb.p);
        GRL_LAUNCH("next_start_bits", 0, next_start_bits_kernel, grid_for(n_words, 256), 256, 0, c->st, eb.p, c->n, sb.p);
        BitmapCompactor bc;
        const u64 ns = bc.count(sb.p, c->n, c->st);
        DevBuf<u64> ptrs(ns + 1, c->st), mx(1, c->st);
        bc.write<u64>(nullptr, ptrs.p);
        GRL_CUDA(cudaMemcpyAsync(ptrs.p + ns, &c->n, 8, cudaMemcpyHostToDevice, c->st));
        mx.zero();
        GRL_LAUNCH("adjacent_max_diff", 0, adjacent_max_diff_kernel, 296, 256, 0, c->st, ptrs.p, ns, mx.p);
        s.longest_string = d2h_scalar(mx.p, c->st);
    }
    c->sep = s.sep_sym;
    c->n_strings = s.n_strings;
    c->alphabet = s.max_sym + 1;  // exact_par_phase.cpp:316
    c->have_stats = true;
}

template <class F>
int guarded(grlgpu_ctx* c, F&& f) {
    struct CtxGuard {  // per-call binding of the context's launch accounting and memory pool
        explicit CtxGuard(grlgpu_ctx* x) { g_prof = x ? &x->prof : nullptr; g_pool = x ? &x->pool : nullptr; }
        ~CtxGuard() { g_prof = nullptr; g_pool = nullptr; }
    } cg(c);
    try {
        if (c) GRL_CUDA(cudaSetDevice(c->device));
        f();
        return GRLGPU_OK;
    } catch (const Error& e) {
        if (c) c->last_error = e.what();
        cudaGetLastError();
        if (e.code == GRLGPU_ERR_CUDA && std::string(e.what()).find("out of memory") != std::string::npos) return GRLGPU_ERR_NOMEM;
        return e.code;
    } catch (const std::exception& e) {
        if (c) c->last_error = e.what();
        return GRLGPU_ERR_CUDA;
    }
}

}  // namespace

extern "C" {

int grlgpu_create_on_stream(grlgpu_ctx** ctx, int device, uint64_t flags, void* cuda_stream);
int grlgpu_create(grlgpu_ctx** ctx, int device, uint64_t flags) { return grlgpu_create_on_stream(ctx, device, flags, nullptr); }

int grlgpu_create_on_stream(grlgpu_ctx** ctx, int device, uint64_t flags, void* cuda_stream) {
    if (!ctx) return GRLGPU_ERR_ARG;
    *ctx = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0 || device < 0 || device >= n_dev) return GRLGPU_ERR_CUDA;
    std::unique_ptr<grlgpu_ctx> c(new grlgpu_ctx());
    c->device = device;
    c->flags = flags;
    int rc = guarded(c.get(), [&] {
        c->pool.init(device);
        if (cuda_stream) { c->st = (cudaStream_t)cuda_stream; c->own_stream = false; }
        else GRL_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    });
    if (rc != GRLGPU_OK) return rc;
    *ctx = c.release();
    return GRLGPU_OK;
}

int grlgpu_destroy(grlgpu_ctx* ctx) {
    if (!ctx) return GRLGPU_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    cudaStream_t st = ctx->st;
    const bool own = ctx->own_stream;
    ctx->prof.resolve();
    delete ctx->mg;
    ctx->mg = nullptr;
    if (ctx->copy_st) { cudaStreamSynchronize(ctx->copy_st); cudaStreamDestroy(ctx->copy_st); cudaEventDestroy(ctx->copy_ev); }
    delete ctx;  // the stream was synchronised above: the pool's slabs are idle
    if (own) cudaStreamDestroy(st);
    return GRLGPU_OK;
}

static int set_text_common(grlgpu_ctx* ctx, const void* text, uint64_t n_syms, int sym_bytes, bool on_device) {
    if (!ctx || !text || n_syms == 0 || !(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) return GRLGPU_ERR_ARG;
    if (on_device && ((uintptr_t)text & 15)) return GRLGPU_ERR_ARG;
    return guarded(ctx, [&] {
        ctx->n = n_syms; ctx->w = sym_bytes; ctx->first = true; ctx->round = 0; ctx->done = false; ctx->have_stats = false;
        if (on_device) { ctx->text_own.release(); ctx->text = text; }
        else {
            ctx->text_own.alloc(n_syms * (u64)sym_bytes + 16, ctx->st);
            GRL_CUDA(cudaMemcpyAsync(ctx->text_own.p, text, n_syms * (u64)sym_bytes, cudaMemcpyHostToDevice, ctx->st));
            GRL_CUDA(cudaStreamSynchronize(ctx->st));
            ctx->text = ctx->text_own.p;
        }
    });
}
int grlgpu_set_text(grlgpu_ctx* ctx, const void* text, uint64_t n_syms, int sym_bytes) { return set_text_common(ctx, text, n_syms, sym_bytes, false); }
int grlgpu_set_text_device(grlgpu_ctx* ctx, const void* dev_text, uint64_t n_syms, int sym_bytes) { return set_text_common(ctx, dev_text, n_syms, sym_bytes, true); }

int grlgpu_stats(grlgpu_ctx* ctx, grlgpu_stats_t* out) {
    if (!ctx || !out) return GRLGPU_ERR_ARG;
    if (!ctx->text || !ctx->first) return GRLGPU_ERR_STATE;
    int rc = guarded(ctx, [&] {
        if (!ctx->have_stats) {
            switch (ctx->w) {
                case 1: compute_stats<u8>(ctx); break;
                case 2: compute_stats<u16>(ctx); break;
                case 4: compute_stats<u32>(ctx); break;
                default: compute_stats<u64>(ctx); break;
            }
        }
    });
    if (rc == GRLGPU_OK) *out = ctx->stats;
    return rc;
}

int grlgpu_round(grlgpu_ctx* ctx, grlgpu_round_t* out) {
    if (!ctx || !out) return GRLGPU_ERR_ARG;
    if (!ctx->text || ctx->done) return GRLGPU_ERR_STATE;
    if (!ctx->have_stats) {
        grlgpu_stats_t s;
        int rc = grlgpu_stats(ctx, &s);
        if (rc != GRLGPU_OK) return rc;
    }
    return guarded(ctx, [&] { run_round(ctx, out); });
}

int grlgpu_fetch_level(grlgpu_ctx* ctx, void* rule_l, void* rule_r, uint8_t* has_hocc, void* pre_sym, uint64_t* pre_len) {
    if (!ctx) return GRLGPU_ERR_ARG;
    if (ctx->round == 0 || !ctx->rule_l.p) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        const u64 sb = (u64)ctx->lvl_sym_bytes;
        if (rule_l) GRL_CUDA(cudaMemcpyAsync(rule_l, ctx->rule_l.p, ctx->lvl_tot * sb, cudaMemcpyDeviceToHost, ctx->st));
        if (rule_r) GRL_CUDA(cudaMemcpyAsync(rule_r, ctx->rule_r.p, ctx->lvl_tot * sb, cudaMemcpyDeviceToHost, ctx->st));
        if (has_hocc) GRL_CUDA(cudaMemcpyAsync(has_hocc, ctx->has_hocc.p, ctx->lvl_tot, cudaMemcpyDeviceToHost, ctx->st));
        if (pre_sym) GRL_CUDA(cudaMemcpyAsync(pre_sym, ctx->pre_sym.p, ctx->lvl_npre * sb, cudaMemcpyDeviceToHost, ctx->st));
        if (pre_len) GRL_CUDA(cudaMemcpyAsync(pre_len, ctx->pre_len.p, ctx->lvl_npre * 8, cudaMemcpyDeviceToHost, ctx->st));
        GRL_CUDA(cudaStreamSynchronize(ctx->st));
    });
}

static __global__ void __launch_bounds__(256) narrow_u64_kernel(const u64* __restrict__ in, u64 n, u32* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (u32)in[i];
}
// Same as grlgpu_fetch_level / _async with 32-bit run lengths (the run lengths of a level sum to at most n_in + parse_len,
// so this applies whenever that is < 2^32): 4 bytes less per preliminary-BWT run over PCIe.
int grlgpu_fetch_level32(grlgpu_ctx* ctx, void* rule_l, void* rule_r, uint8_t* has_hocc, void* pre_sym, uint32_t* pre_len32, int async) {
    if (!ctx) return GRLGPU_ERR_ARG;
    if (ctx->round == 0 || !ctx->rule_l.p) return GRLGPU_ERR_STATE;
    if (ctx->lvl_n_in >= (1ull << 32)) return GRLGPU_ERR_LIMIT;
    return guarded(ctx, [&] {
        cudaStream_t cs = ctx->st;
        DevBuf<u32> len32(pre_len32 ? ctx->lvl_npre : 0, ctx->st);
        if (pre_len32 && ctx->lvl_npre)
            GRL_LAUNCH("narrow_u64", ctx->lvl_npre * 12, narrow_u64_kernel, grid_for(ctx->lvl_npre, 256), 256, 0, ctx->st, ctx->pre_len.p, ctx->lvl_npre, len32.p);
        if (async) {
            if (!ctx->copy_st) {
                GRL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
                GRL_CUDA(cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming));
            }
            GRL_CUDA(cudaEventRecord(ctx->copy_ev, ctx->st));
            GRL_CUDA(cudaStreamWaitEvent(ctx->copy_st, ctx->copy_ev, 0));
            cs = ctx->copy_st;
        }
        const u64 sb = (u64)ctx->lvl_sym_bytes;
        if (pre_sym) GRL_CUDA(cudaMemcpyAsync(pre_sym, ctx->pre_sym.p, ctx->lvl_npre * sb, cudaMemcpyDeviceToHost, cs));
        if (pre_len32) GRL_CUDA(cudaMemcpyAsync(pre_len32, len32.p, ctx->lvl_npre * 4, cudaMemcpyDeviceToHost, cs));
        if (rule_l) GRL_CUDA(cudaMemcpyAsync(rule_l, ctx->rule_l.p, ctx->lvl_tot * sb, cudaMemcpyDeviceToHost, cs));
        if (rule_r) GRL_CUDA(cudaMemcpyAsync(rule_r, ctx->rule_r.p, ctx->lvl_tot * sb, cudaMemcpyDeviceToHost, cs));
        if (has_hocc) GRL_CUDA(cudaMemcpyAsync(has_hocc, ctx->has_hocc.p, ctx->lvl_tot, cudaMemcpyDeviceToHost, cs));
        if (async) {
            ctx->parked_u8.push_back(std::move(ctx->rule_l));
            ctx->parked_u8.push_back(std::move(ctx->rule_r));
            ctx->parked_u8.push_back(std::move(ctx->has_hocc));
            ctx->parked_u8.push_back(std::move(ctx->pre_sym));
            ctx->parked_u32.push_back(std::move(len32));
            ctx->pre_len.release();
        } else GRL_CUDA(cudaStreamSynchronize(ctx->st));
    });
}

int grlgpu_fetch_level_async(grlgpu_ctx* ctx, void* rule_l, void* rule_r, uint8_t* has_hocc, void* pre_sym, uint64_t* pre_len) {
    if (!ctx) return GRLGPU_ERR_ARG;
    if (ctx->round == 0 || !ctx->rule_l.p) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        if (!ctx->copy_st) {
            GRL_CUDA(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
            GRL_CUDA(cudaEventCreateWithFlags(&ctx->copy_ev, cudaEventDisableTiming));
        }
        // the round that produced the level has been synchronised already; order the copy stream after it anyway
        GRL_CUDA(cudaEventRecord(ctx->copy_ev, ctx->st));
        GRL_CUDA(cudaStreamWaitEvent(ctx->copy_st, ctx->copy_ev, 0));
        const u64 sb = (u64)ctx->lvl_sym_bytes;
        if (rule_l) GRL_CUDA(cudaMemcpyAsync(rule_l, ctx->rule_l.p, ctx->lvl_tot * sb, cudaMemcpyDeviceToHost, ctx->copy_st));
        if (rule_r) GRL_CUDA(cudaMemcpyAsync(rule_r, ctx->rule_r.p, ctx->lvl_tot * sb, cudaMemcpyDeviceToHost, ctx->copy_st));
        if (has_hocc) GRL_CUDA(cudaMemcpyAsync(has_hocc, ctx->has_hocc.p, ctx->lvl_tot, cudaMemcpyDeviceToHost, ctx->copy_st));
        if (pre_sym) GRL_CUDA(cudaMemcpyAsync(pre_sym, ctx->pre_sym.p, ctx->lvl_npre * sb, cudaMemcpyDeviceToHost, ctx->copy_st));
        if (pre_len) GRL_CUDA(cudaMemcpyAsync(pre_len, ctx->pre_len.p, ctx->lvl_npre * 8, cudaMemcpyDeviceToHost, ctx->copy_st));
        ctx->parked_u8.push_back(std::move(ctx->rule_l));
        ctx->parked_u8.push_back(std::move(ctx->rule_r));
        ctx->parked_u8.push_back(std::move(ctx->has_hocc));
        ctx->parked_u8.push_back(std::move(ctx->pre_sym));
        ctx->parked_u64.push_back(std::move(ctx->pre_len));
    });
}

int grlgpu_fetch_wait(grlgpu_ctx* ctx) {
    if (!ctx) return GRLGPU_ERR_ARG;
    return guarded(ctx, [&] {
        if (ctx->copy_st) GRL_CUDA(cudaStreamSynchronize(ctx->copy_st));
        ctx->parked_u8.clear();
        ctx->parked_u64.clear();
        ctx->parked_u32.clear();
    });
}

int grlgpu_fetch_parse(grlgpu_ctx* ctx, void* dst) {
    if (!ctx || !dst) return GRLGPU_ERR_ARG;
    if (ctx->round == 0) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        GRL_CUDA(cudaMemcpyAsync(dst, ctx->text, ctx->n * (u64)ctx->w, cudaMemcpyDeviceToHost, ctx->st));
        GRL_CUDA(cudaStreamSynchronize(ctx->st));
    });
}

int grlgpu_fetch_str_ptrs(grlgpu_ctx* ctx, uint64_t* dst) {
    if (!ctx || !dst) return GRLGPU_ERR_ARG;
    if (ctx->round == 0) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        const u64 n_words = div_up(ctx->n, 32);
        DevBuf<u32> sb(n_words, ctx->st);
        GRL_LAUNCH("next_start_bits", 0, next_start_bits_kernel, grid_for(n_words, 256), 256, 0, ctx->st, ctx->end_bits.p, ctx->n, sb.p);
        BitmapCompactor bc;
        const u64 ns = bc.count(sb.p, ctx->n, ctx->st);
        if (ns != ctx->n_strings) throw Error(GRLGPU_ERR_STATE, "string count changed between rounds");
        DevBuf<u64> ptrs(ns + 1, ctx->st);
        bc.write<u64>(nullptr, ptrs.p);
        GRL_CUDA(cudaMemcpyAsync(dst, ptrs.p, ns * 8, cudaMemcpyDeviceToHost, ctx->st));
        GRL_CUDA(cudaStreamSynchronize(ctx->st));
        dst[ns] = ctx->n;
    });
}

int grlgpu_fetch_dictionary(grlgpu_ctx* ctx, uint64_t* syms, uint64_t* lens, uint64_t* freqs, uint64_t* metas) {
    if (!ctx) return GRLGPU_ERR_ARG;
    if (ctx->round == 0 || !(ctx->flags & GRLGPU_FLAG_KEEP_DICT) || !ctx->kd_order.p) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        const u64 d = ctx->kd_d, nE = ctx->kd_nE, nS = ctx->kd_nS;
        std::vector<u32> order(nS), phr_of(nE), off(d + 1), len(d);
        std::vector<u64> freq(d), meta(d);
        std::vector<u8> Draw(nE * (u64)ctx->lvl_sym_bytes);
        GRL_CUDA(cudaMemcpy(order.data(), ctx->kd_order.p, nS * 4, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(phr_of.data(), ctx->kd_phr_of.p, nE * 4, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(off.data(), ctx->kd_off.p, (d + 1) * 4, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(len.data(), ctx->kd_len.p, d * 4, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(freq.data(), ctx->kd_freq.p, d * 8, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(meta.data(), ctx->kd_meta.p, d * 8, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(Draw.data(), ctx->kd_D.p, Draw.size(), cudaMemcpyDeviceToHost));
        u64 k = 0, so = 0;
        for (u64 i = 0; i < nS; i++) {  // full-phrase entries in suffix order = phrases in A.2 order
            const u32 e = order[i], ph = phr_of[e];
            if (e != off[ph]) continue;
            if (lens) lens[k] = len[ph];
            if (freqs) freqs[k] = freq[ph];
            if (metas) metas[k] = meta[ph];
            if (syms)
                for (u32 t = 0; t < len[ph]; t++)
                    syms[so + t] = ctx->lvl_sym_bytes == 4 ? ((const u32*)Draw.data())[e + t] : ((const u64*)Draw.data())[e + t];
            so += len[ph];
            k++;
        }
    });
}

int grlgpu_histogram(grlgpu_ctx* ctx, uint64_t* hist256) {
    if (!ctx || !hist256) return GRLGPU_ERR_ARG;
    if (!ctx->have_stats) return GRLGPU_ERR_STATE;
    memcpy(hist256, ctx->hist, sizeof(ctx->hist));
    return GRLGPU_OK;
}

// ---- multi-GPU rounds ----
#define MG_DISPATCH_FIRST(fn, ...)                                                     \
    do {                                                                               \
        if (ctx->first) switch (ctx->w) {                                              \
            case 1: fn<u8, true>(__VA_ARGS__); break;                                  \
            case 2: fn<u16, true>(__VA_ARGS__); break;                                 \
            case 4: fn<u32, true>(__VA_ARGS__); break;                                 \
            default: fn<u64, true>(__VA_ARGS__); break;                                \
        } else switch (ctx->w) {                                                       \
            case 1: fn<u8, false>(__VA_ARGS__); break;                                 \
            case 2: fn<u16, false>(__VA_ARGS__); break;                                \
            case 4: fn<u32, false>(__VA_ARGS__); break;                                \
            default: fn<u64, false>(__VA_ARGS__); break;                               \
        }                                                                              \
    } while (0)
#define MG_DISPATCH(fn, ...)                                                           \
    do {                                                                               \
        switch (ctx->w) {                                                              \
            case 1: fn<u8>(__VA_ARGS__); break;                                        \
            case 2: fn<u16>(__VA_ARGS__); break;                                       \
            case 4: fn<u32>(__VA_ARGS__); break;                                       \
            default: fn<u64>(__VA_ARGS__); break;                                      \
        }                                                                              \
    } while (0)

int grlgpu_mg_set_alphabet(grlgpu_ctx* ctx, uint64_t global_max_sym) {
    if (!ctx) return GRLGPU_ERR_ARG;
    if (!ctx->have_stats || !ctx->first) return GRLGPU_ERR_STATE;
    if (global_max_sym < ctx->stats.max_sym) return GRLGPU_ERR_ARG;
    ctx->alphabet = global_max_sym + 1;
    return GRLGPU_OK;
}
int grlgpu_mg_local(grlgpu_ctx* ctx, int n_ranks, grlgpu_part_t* per_owner, uint64_t* parse_len_local) {
    if (!ctx || !per_owner || !parse_len_local || n_ranks < 1 || n_ranks > 31) return GRLGPU_ERR_ARG;
    if (!ctx->text || ctx->done || !ctx->have_stats) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        u64 pl = 0;
        MG_DISPATCH_FIRST(mg_local_t, ctx, n_ranks, per_owner, &pl);
        *parse_len_local = pl;
    });
}
int grlgpu_mg_pack(grlgpu_ctx* ctx, uint32_t* d_lens, uint64_t* d_counts, void* d_cells) {
    if (!ctx || !d_lens || !d_counts || !d_cells) return GRLGPU_ERR_ARG;
    if (!ctx->mg) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] { MG_DISPATCH(mg_pack_t, ctx, d_lens, (u64*)d_counts, d_cells); });
}
int grlgpu_mg_merge(grlgpu_ctx* ctx, const uint32_t* d_lens, const uint64_t* d_counts, const void* d_cells, uint64_t m, uint64_t n_cells, grlgpu_part_t* part) {
    if (!ctx || !part || !d_lens || !d_counts || !d_cells) return GRLGPU_ERR_ARG;
    if (!ctx->mg) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] { MG_DISPATCH(mg_merge_t, ctx, d_lens, (const u64*)d_counts, d_cells, (u64)m, (u64)n_cells, part); });
}
int grlgpu_mg_pack_part(grlgpu_ctx* ctx, uint32_t* d_lens, uint64_t* d_freqs, void* d_cells) {
    if (!ctx || !d_lens || !d_freqs || !d_cells) return GRLGPU_ERR_ARG;
    if (!ctx->mg) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] { MG_DISPATCH(mg_pack_part_t, ctx, d_lens, (u64*)d_freqs, d_cells); });
}
int grlgpu_mg_global(grlgpu_ctx* ctx, const uint32_t* d_lens, const uint64_t* d_freqs, const void* d_cells, uint64_t d, uint64_t n_cells, int done_global,
                     grlgpu_round_t* out) {
    if (!ctx || !out || !d_lens || !d_freqs || !d_cells || d == 0) return GRLGPU_ERR_ARG;
    if (!ctx->mg) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] { MG_DISPATCH_FIRST(mg_global_t, ctx, d_lens, (const u64*)d_freqs, d_cells, (u64)d, (u64)n_cells, done_global, out); });
}

int grlgpu_mg_rank_sort(grlgpu_ctx* ctx, const uint32_t* d_lens, const uint64_t* d_freqs, const void* d_cells, uint64_t d, uint64_t n_cells, int rank_id, int n_ranks,
                        uint64_t* info5) {
    if (!ctx || !info5 || !d_lens || !d_freqs || !d_cells || d == 0 || rank_id < 0 || rank_id >= n_ranks) return GRLGPU_ERR_ARG;
    if (!ctx->mg) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] { MG_DISPATCH_FIRST(mg_rank_sort_t, ctx, d_lens, (const u64*)d_freqs, d_cells, (u64)d, (u64)n_cells, rank_id, n_ranks, (u64*)info5); });
}
int grlgpu_mg_rank_apply(grlgpu_ctx* ctx, uint64_t rank_base, uint64_t* d_ph_meta, uint8_t* d_is_suffix_next, uint32_t* d_erank1) {
    if (!ctx || !d_ph_meta || !d_is_suffix_next || !d_erank1) return GRLGPU_ERR_ARG;
    if (!ctx->mg || !ctx->mg->GR) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        if (ctx->mg->sym_bytes == 8) mg_rank_apply_sym<u64>(ctx, rank_base, (u64*)d_ph_meta, d_is_suffix_next, d_erank1);
        else mg_rank_apply_sym<u32>(ctx, rank_base, (u64*)d_ph_meta, d_is_suffix_next, d_erank1);
    });
}
extern "C++" {
template <class CellT>
static void mg_rank_finish_cell(grlgpu_ctx* ctx, u64 rank_base, u64 tot, u64 n_pre, const u64* m, const u8* s, u32* e, int done, grlgpu_round_t* out, const u64* lm) {
    if (ctx->mg->sym_bytes == 8) mg_rank_finish_sym<CellT, u64>(ctx, rank_base, tot, n_pre, m, s, e, done, out, lm);
    else mg_rank_finish_sym<CellT, u32>(ctx, rank_base, tot, n_pre, m, s, e, done, out, lm);
}
}
int grlgpu_mg_reply(grlgpu_ctx* ctx, uint64_t part_base, const uint64_t* d_ph_meta, uint64_t* d_reply) {
    if (!ctx || !d_ph_meta || !d_reply) return GRLGPU_ERR_ARG;
    if (!ctx->mg || !ctx->mg->recv_dense.p) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        MgRound& M = *ctx->mg;
        GRL_LAUNCH("reply_meta", M.m_recv * 20, reply_meta_kernel, grid_for(M.m_recv, 256), 256, 0, ctx->st, M.recv_dense.p, M.m_recv, (u64)part_base, (const u64*)d_ph_meta,
                   (u64*)d_reply);
        GRL_CUDA(cudaStreamSynchronize(ctx->st));
    });
}
int grlgpu_mg_rank_finish(grlgpu_ctx* ctx, uint64_t rank_base, uint64_t tot, uint64_t n_pre_runs, const uint64_t* d_ph_meta, const uint8_t* d_is_suffix_next,
                          uint32_t* d_erank1, const uint64_t* d_local_meta, int done_global, grlgpu_round_t* out) {
    if (!ctx || !out || !d_ph_meta || !d_is_suffix_next || !d_erank1) return GRLGPU_ERR_ARG;
    if (!ctx->mg || !ctx->mg->GR) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        MG_DISPATCH(mg_rank_finish_cell, ctx, (u64)rank_base, (u64)tot, (u64)n_pre_runs, (const u64*)d_ph_meta, d_is_suffix_next, d_erank1, done_global, out, (const u64*)d_local_meta);
    });
}
int grlgpu_mg_level_slice(grlgpu_ctx* ctx, void* d_rule_l, void* d_rule_r, uint8_t* d_has_hocc, void* d_pre_sym, uint64_t* d_pre_len) {
    if (!ctx) return GRLGPU_ERR_ARG;
    if (!ctx->mg) return GRLGPU_ERR_STATE;
    return guarded(ctx, [&] {
        MgRound& M = *ctx->mg;
        const u64 sb = (u64)M.sym_bytes;
        if (d_rule_l) GRL_CUDA(cudaMemcpyAsync(d_rule_l, M.sl_rule_l.p, M.tot_local * sb, cudaMemcpyDeviceToDevice, ctx->st));
        if (d_rule_r) GRL_CUDA(cudaMemcpyAsync(d_rule_r, M.sl_rule_r.p, M.tot_local * sb, cudaMemcpyDeviceToDevice, ctx->st));
        if (d_has_hocc) GRL_CUDA(cudaMemcpyAsync(d_has_hocc, M.sl_has_hocc.p, M.tot_local, cudaMemcpyDeviceToDevice, ctx->st));
        if (d_pre_sym) GRL_CUDA(cudaMemcpyAsync(d_pre_sym, M.sl_pre_sym.p, M.n_pre_local * sb, cudaMemcpyDeviceToDevice, ctx->st));
        if (d_pre_len) GRL_CUDA(cudaMemcpyAsync(d_pre_len, M.sl_pre_len.p, M.n_pre_local * 8, cudaMemcpyDeviceToDevice, ctx->st));
        GRL_CUDA(cudaStreamSynchronize(ctx->st));
    });
}

int grlgpu_profile_enable(grlgpu_ctx* ctx, int on) {
    if (!ctx) return GRLGPU_ERR_ARG;
    ctx->prof.resolve();
    ctx->prof.timing = on != 0;
    return GRLGPU_OK;
}
int grlgpu_profile_reset(grlgpu_ctx* ctx) {
    if (!ctx) return GRLGPU_ERR_ARG;
    ctx->prof.reset();
    return GRLGPU_OK;
}
uint64_t grlgpu_launch_count(const grlgpu_ctx* ctx) { return ctx ? ctx->prof.launches : 0; }
int grlgpu_profile_entry(grlgpu_ctx* ctx, int index, char* name, int name_cap, uint64_t* launches, double* total_ms, uint64_t* model_bytes) {
    if (!ctx || !name || name_cap <= 0) return GRLGPU_ERR_ARG;
    ctx->prof.resolve();
    if (index < 0 || (size_t)index >= ctx->prof.acc.size()) return 1;
    auto it = ctx->prof.acc.begin();
    std::advance(it, index);
    snprintf(name, (size_t)name_cap, "%s", it->first.c_str());
    if (launches) *launches = it->second.launches;
    if (total_ms) *total_ms = it->second.ms;
    if (model_bytes) *model_bytes = it->second.bytes;
    return GRLGPU_OK;
}

const char* grlgpu_strerror(int status) {
    switch (status) {
        case GRLGPU_OK: return "ok";
        case GRLGPU_ERR_ARG: return "invalid argument";
        case GRLGPU_ERR_ILL_FORMED: return "the collection is ill formed";
        case GRLGPU_ERR_CUDA: return "CUDA error or no usable device";
        case GRLGPU_ERR_NOMEM: return "out of device memory";
        case GRLGPU_ERR_STATE: return "call out of order";
        case GRLGPU_ERR_LIMIT: return "size limit of the device path exceeded";
    }
    return "unknown status";
}
const char* grlgpu_last_error(const grlgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

// ---- self-test hooks ----
int grlgpu_selftest_scan(const uint32_t* in, uint64_t n, uint64_t* out_exclusive, uint64_t* total) {
    return guarded(nullptr, [&] {
        cudaStream_t st = nullptr;
        DevBuf<u32> din(n, st);
        DevBuf<u64> dout(n, st), tot(1, st);
        GRL_CUDA(cudaMemcpy(din.p, in, n * 4, cudaMemcpyHostToDevice));
        exclusive_scan<u32, u64>(din.p, dout.p, n, tot.p, st);
        GRL_CUDA(cudaMemcpy(out_exclusive, dout.p, n * 8, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(total, tot.p, 8, cudaMemcpyDeviceToHost));
    });
}
int grlgpu_selftest_sort(uint64_t* keys, uint32_t* vals, uint64_t n, int n_bits) {
    return guarded(nullptr, [&] {
        cudaStream_t st = nullptr;
        DevBuf<u64> k(n, st), ka(n, st);
        DevBuf<u32> v(n, st), va(n, st);
        GRL_CUDA(cudaMemcpy(k.p, keys, n * 8, cudaMemcpyHostToDevice));
        GRL_CUDA(cudaMemcpy(v.p, vals, n * 4, cudaMemcpyHostToDevice));
        u64 *kp = k.p, *kap = ka.p;
        u32 *vp = v.p, *vap = va.p;
        radix_sort_pairs(&kp, &vp, &kap, &vap, n, n_bits, st);
        GRL_CUDA(cudaStreamSynchronize(st));
        GRL_CUDA(cudaMemcpy(keys, kp, n * 8, cudaMemcpyDeviceToHost));
        GRL_CUDA(cudaMemcpy(vals, vp, n * 4, cudaMemcpyDeviceToHost));
    });
}
int grlgpu_selftest_compact(const uint32_t* bits, const uint32_t* prev_bits, uint64_t n_bits, uint64_t* out, uint64_t* count) {
    return guarded(nullptr, [&] {
        cudaStream_t st = nullptr;
        const u64 n_words = div_up(n_bits, 32);
        DevBuf<u32> b(n_words, st), pb(n_words, st);
        GRL_CUDA(cudaMemcpy(b.p, bits, n_words * 4, cudaMemcpyHostToDevice));
        if (prev_bits) GRL_CUDA(cudaMemcpy(pb.p, prev_bits, n_words * 4, cudaMemcpyHostToDevice));
        BitmapCompactor bc;
        const u64 cnt = bc.count(b.p, n_bits, st);
        DevBuf<u64> o(cnt, st);
        bc.write<u64>(prev_bits ? pb.p : nullptr, o.p);
        GRL_CUDA(cudaMemcpy(out, o.p, cnt * 8, cudaMemcpyDeviceToHost));
        *count = cnt;
    });
}

}  // extern "C"
