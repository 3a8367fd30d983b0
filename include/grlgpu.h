/*
 * grlgpu.h -- C ABI of the B200 (sm_100a) parse phase of grlBWT.
 *
 * The reference (ddiazdom/grlBWT) has no FFI: the path sits behind C++ templates. This header is the
 * boundary a maintainer binds instead of those templates; every entry point cites the reference
 * interface it replaces (paths relative to the reference tree). Conventions: plain pointers and
 * sizes, caller-owned buffers, int status (0 = ok, < 0 = grlgpu_status), no exceptions and no
 * exit() across the boundary, one context driven by one host thread, one context per GPU.
 * There is no CPU fallback: every call fails with GRLGPU_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef GRLGPU_H
#define GRLGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct grlgpu_ctx grlgpu_ctx;

enum grlgpu_status {
    GRLGPU_OK = 0,
    GRLGPU_ERR_ARG = -1,        /* bad argument (null pointer, symbol width not in {1,2,4,8}, empty text) */
    GRLGPU_ERR_ILL_FORMED = -2, /* "Error: the file is ill formed" (external/cdt/lib/utils.cpp:177-180) */
    GRLGPU_ERR_CUDA = -3,       /* CUDA runtime error / no device */
    GRLGPU_ERR_NOMEM = -4,      /* device memory exhausted */
    GRLGPU_ERR_STATE = -5,      /* call out of order (no text, phase already finished, ...) */
    GRLGPU_ERR_LIMIT = -6       /* dictionary of one round exceeds 2^32 symbols or the table 2^31 slots */
};

/* flags of grlgpu_create */
#define GRLGPU_FLAG_SMALL_TABLE 1ull /* tests: start the phrase table tiny so the regrow path runs */
#define GRLGPU_FLAG_FORCE_SLOW_SCAN 2ull /* tests: always take the long-run (summary/resolve) LMS path */
#define GRLGPU_FLAG_FORCE_UNCACHED 8ull /* tests: dedup every phrase with the thread-per-phrase kernel (no cached tiles) */
#define GRLGPU_FLAG_SMALL_PILOT 16ull /* tests: one pilot tile, so small inputs exercise pilot + remainder */
#define GRLGPU_FLAG_FORCE_DOUBLING 32ull /* tests: refine suffix groups by prefix doubling even when key extension would do */
#define GRLGPU_FLAG_FORCE_DIST_RANK 64ull /* tests: distribute the dictionary ranking over the ranks even for tiny dictionaries */
#define GRLGPU_FLAG_KEEP_DICT 4ull /* tests: keep the last round's dictionary for grlgpu_fetch_dictionary */

/* = str_collection (external/cdt/include/utils.h:20-28) as filled by collection_stats (utils.cpp:100-189) */
typedef struct {
    uint64_t n_syms;
    uint64_t n_strings;
    uint64_t longest_string;
    uint64_t min_sym;
    uint64_t max_sym;
    uint64_t max_sym_freq; /* byte alphabet: highest histogram count; wider alphabets: n_syms (utils.cpp:117) */
    uint64_t sep_sym;
} grlgpu_stats_t;

/* what one parse round reports (the "Stats:" block of par_round, lib/exact_algo/exact_par_phase.cpp:484-488,
 * plus the sizes the caller needs to fetch the level's artefacts) */
typedef struct {
    uint64_t round;          /* 1-based */
    uint64_t n_in;           /* cells of the round's input text */
    uint64_t n_strings;
    uint64_t parse_len;      /* p: phrase occurrences = cells of the output parse ("Parse size") */
    uint64_t n_phrases;      /* d: distinct phrases ("Parsing phrases") */
    uint64_t dict_syms;      /* sum of phrase lengths ("Number of symbols in the phrases") */
    uint64_t max_freq;       /* highest phrase frequency */
    uint64_t alphabet;       /* A: alphabet size of the input text (= tot_phrases of the previous round) */
    uint64_t tot_phrases;    /* ranks handed out = "Number of unsolved BWT blocks" */
    uint64_t n_pre_runs;     /* runs of the preliminary BWT of this level */
    uint64_t algorithmic_bytes; /* B_r = n*w + p*w' + dict_syms*w + 8*d */
    uint32_t cell_bytes_in;  /* w  */
    uint32_t cell_bytes_out; /* w' by sym_width(tot_phrases)+1 (exact_par_phase.cpp:456-465) */
    uint32_t sym_bytes;      /* element width (4 or 8) of rule_l / rule_r / pre_sym in grlgpu_fetch_level */
    uint32_t done;           /* 1 when parse_len == n_strings (exact_par_phase.cpp:496) */
    float device_ms;         /* CUDA-event time of the whole round on the context's stream */
    float text_pass_ms;      /* of which: boundary scan + compaction + dedup (text read) */
    float dict_ms;           /* of which: dictionary gather + suffix ordering + ranks + rules */
    float rewrite_ms;        /* of which: rewrite */
} grlgpu_round_t;

/* context on one device. replaces: construction of the parse strategy (exact_par_phase.cpp:265-283) */
int grlgpu_create(grlgpu_ctx** ctx, int device, uint64_t flags);
/* same, but every kernel and copy is issued on the caller's CUDA stream (a cudaStream_t), so the
 * caller can bracket calls with its own events (bench.py passes torch's current stream) */
int grlgpu_create_on_stream(grlgpu_ctx** ctx, int device, uint64_t flags, void* cuda_stream);
int grlgpu_destroy(grlgpu_ctx* ctx);

/* round-1 input, caller-owned HOST memory, copied to the device. replaces: i_file_stream over the
 * input file (external/cdt/include/file_streams.hpp:93-105) */
int grlgpu_set_text(grlgpu_ctx* ctx, const void* text, uint64_t n_syms, int sym_bytes);
/* same, text already resident in DEVICE memory (16-byte aligned); borrowed, not copied, never written */
int grlgpu_set_text_device(grlgpu_ctx* ctx, const void* dev_text, uint64_t n_syms, int sym_bytes);

/* replaces: collection_stats<sym_type>() (utils.cpp:100-189); validates sep == min && last == sep */
int grlgpu_stats(grlgpu_ctx* ctx, grlgpu_stats_t* out);

/* one parse round on the device. replaces: par_round<strategy>() (exact_par_phase.cpp:374-497):
 * get_phrases (parsing_strategies.h:244-275 / :618-642), dictionary + process_dictionary
 * (exact_par_phase.hpp:106-183, exact_par_phase.cpp:97-263), metasymbol assignment (:427-450),
 * parse_text (parsing_strategies.h:388-497 / :644-676). The new parse stays on the device. */
int grlgpu_round(grlgpu_ctx* ctx, grlgpu_round_t* out);

/* artefacts of the last round for the induction phase, copied to caller-owned host buffers.
 * replaces: files dict_lev_k (exact_par_phase.hpp:189-203) and pre_bwt_lev_k (exact_par_phase.cpp:150-155).
 *   rule_l, rule_r : tot_phrases elements of sym_bytes each (dict[2u], dict[2u+1] of produce_grammar,
 *                    numeric conventions of exact_par_phase.cpp:19-20: alphabet A+3, dummy A+3+tot+1)
 *   has_hocc       : tot_phrases bytes (phrases_has_hocc)
 *   pre_sym/len    : n_pre_runs runs, maximal; dummies A+1 (from BWT i+1) and A+2 (from hocc buffer)
 * any pointer may be NULL to skip that array. */
int grlgpu_fetch_level(grlgpu_ctx* ctx, void* rule_l, void* rule_r, uint8_t* has_hocc, void* pre_sym, uint64_t* pre_len);

/* same copies, issued on a second stream: the call returns at once and the NEXT round may run while they are in
 * flight (give pinned host buffers, or the driver stages and the overlap is lost). The level's device buffers
 * stay parked until grlgpu_fetch_wait, which must be called before the host buffers are read. After this call the
 * level can no longer be fetched again. */
int grlgpu_fetch_level_async(grlgpu_ctx* ctx, void* rule_l, void* rule_r, uint8_t* has_hocc, void* pre_sym, uint64_t* pre_len);
/* grlgpu_fetch_level with 32-bit run lengths (async != 0: as grlgpu_fetch_level_async, completed by grlgpu_fetch_wait).
 * The run lengths of a level sum to at most n_in + parse_len: GRLGPU_ERR_LIMIT when that is >= 2^32 (use the 64-bit call).
 * 4 bytes less per preliminary-BWT run over PCIe; replaces the same files as grlgpu_fetch_level. */
int grlgpu_fetch_level32(grlgpu_ctx* ctx, void* rule_l, void* rule_r, uint8_t* has_hocc, void* pre_sym, uint32_t* pre_len32, int async);
int grlgpu_fetch_wait(grlgpu_ctx* ctx);

/* current parse (output of the last round): parse_len cells of cell_bytes_out bytes, cells = rank<<1|rep.
 * replaces: file tmp_input (exact_par_phase.cpp:307-308); after the last round this is the final parse
 * consumed by parse2bwt (exact_ind_phase.cpp:603-672) */
int grlgpu_fetch_parse(grlgpu_ctx* ctx, void* dst);
/* start offset of every string in the current parse, n_strings+1 entries (parsing_info::str_ptrs) */
int grlgpu_fetch_str_ptrs(grlgpu_ctx* ctx, uint64_t* dst);
/* dictionary of the last round in insertion-independent form, for tests: phrases listed in the A.2
 * order; syms: dict_syms u64 values, lens/freqs/metas: n_phrases entries (NULL to skip) */
int grlgpu_fetch_dictionary(grlgpu_ctx* ctx, uint64_t* syms, uint64_t* lens, uint64_t* freqs, uint64_t* metas);

const char* grlgpu_strerror(int status);
const char* grlgpu_last_error(const grlgpu_ctx* ctx);

/* ---- multi-GPU rounds ---------------------------------------------------------------------
 * One context per GPU, each holding a shard of WHOLE strings (the reference's own split,
 * include/parsing_strategies.h:208-214, so no phrase crosses a shard). The caller owns the exchange (NCCL
 * through torch.distributed in bench.py / grlbwt_b200/multigpu.py); the library does every device step
 * before, between and after. Per round, on every rank, in this order:
 *   grlgpu_mg_local      boundary scan + local dedup; per owner rank (content hash % n_ranks) how many distinct
 *                        local phrases / cells go there; local parse length (caller all-reduces termination)
 *   grlgpu_mg_pack       fill caller-allocated DEVICE send buffers, ordered by owner   -> all-to-all-v
 *   grlgpu_mg_merge      owner side: dedup what arrived, sum the counts                -> partition sizes
 *   grlgpu_mg_pack_part  fill DEVICE buffers with this rank's partition               -> all-gather-v
 *   grlgpu_mg_global     the gathered dictionary (identical on every rank, rank order) is ranked exactly as in
 *                        grlgpu_round; local phrases get their metasymbols; the shard is rewritten.
 * Level artefacts are identical on every rank (fetch them on one). Before the first round every rank calls
 * grlgpu_stats and then grlgpu_mg_set_alphabet with the maximum symbol over all ranks.
 * replaces: mt_parse_strat_t's thread fan-out + serial join_thread_phrases (parsing_strategies.h:244-386). */
typedef struct {
    uint64_t n_phrases;
    uint64_t n_cells;
} grlgpu_part_t;
/* byte alphabets: the 256-bin symbol histogram behind grlgpu_stats (ranks sum theirs to get the global max_sym_freq) */
int grlgpu_histogram(grlgpu_ctx* ctx, uint64_t* hist256);
int grlgpu_mg_set_alphabet(grlgpu_ctx* ctx, uint64_t global_max_sym);
int grlgpu_mg_local(grlgpu_ctx* ctx, int n_ranks, grlgpu_part_t* per_owner, uint64_t* parse_len_local);
int grlgpu_mg_pack(grlgpu_ctx* ctx, uint32_t* d_lens, uint64_t* d_counts, void* d_cells);
int grlgpu_mg_merge(grlgpu_ctx* ctx, const uint32_t* d_lens, const uint64_t* d_counts, const void* d_cells, uint64_t m, uint64_t n_cells,
                    grlgpu_part_t* part);
int grlgpu_mg_pack_part(grlgpu_ctx* ctx, uint32_t* d_lens, uint64_t* d_freqs, void* d_cells);
int grlgpu_mg_global(grlgpu_ctx* ctx, const uint32_t* d_lens, const uint64_t* d_freqs, const void* d_cells, uint64_t d, uint64_t n_cells,
                     int done_global, grlgpu_round_t* out);

/* Distributed ranking of the gathered dictionary (optional; replaces grlgpu_mg_global when info5[0] comes back 1).
 * Every rank sorts, refines and groups only the suffix entries whose first key falls in its range (splitters from
 * a regular sample, identical on every rank), so the dominant cost of unique-heavy rounds shrinks with the ranks:
 *   grlgpu_mg_rank_sort    -> info5 = {distributed?, ranked groups here, pre-BWT runs here, dictionary entries nE, symbol bytes}
 *                             (0 in info5[0]: nothing was done -- small or long-phrase dictionary -- call grlgpu_mg_global)
 *   caller: rank_base = exclusive prefix of the ranked-group counts over the ranks, tot = their sum; allocates
 *           zero-filled device arrays ph_meta[d] (u64), is_suffix_next[tot] (u8), erank1[nE] (u32)
 *   grlgpu_mg_rank_apply   writes this rank's share into them (hocc marks as rank + 1)   -> all-reduce(MAX) of the three
 *   grlgpu_mg_reply        owner side of the metasymbol return (optional): reply[k] = metasymbol of the k-th phrase this rank
 *                          RECEIVED in grlgpu_mg_merge; part_base = global index of this rank's first partition phrase
 *                          -> reverse all-to-all-v: every rank gets the metasymbols of the pack it sent, in pack order
 *   grlgpu_mg_rank_finish  rules of this rank's groups, metasymbols of the local phrases (from d_local_meta when the owners
 *                          returned them, else -- NULL -- by content lookup in a table of the whole dictionary), rewrite
 *   grlgpu_mg_level_slice  this rank's slice of the level artefacts into caller DEVICE buffers (rules / has_hocc: info5[1]
 *                          entries, positions [rank_base, rank_base + info5[1]); pre-BWT: info5[2] runs, to be
 *                          concatenated in rank order, merging equal symbols where two ranks meet) */
int grlgpu_mg_rank_sort(grlgpu_ctx* ctx, const uint32_t* d_lens, const uint64_t* d_freqs, const void* d_cells, uint64_t d, uint64_t n_cells, int rank_id,
                        int n_ranks, uint64_t* info5);
int grlgpu_mg_rank_apply(grlgpu_ctx* ctx, uint64_t rank_base, uint64_t* d_ph_meta, uint8_t* d_is_suffix_next, uint32_t* d_erank1);
int grlgpu_mg_reply(grlgpu_ctx* ctx, uint64_t part_base, const uint64_t* d_ph_meta, uint64_t* d_reply);
int grlgpu_mg_rank_finish(grlgpu_ctx* ctx, uint64_t rank_base, uint64_t tot, uint64_t n_pre_runs, const uint64_t* d_ph_meta, const uint8_t* d_is_suffix_next,
                          uint32_t* d_erank1, const uint64_t* d_local_meta, int done_global, grlgpu_round_t* out);
int grlgpu_mg_level_slice(grlgpu_ctx* ctx, void* d_rule_l, void* d_rule_r, uint8_t* d_has_hocc, void* d_pre_sym, uint64_t* d_pre_len);

/* launch accounting: number of kernel launches issued by this context so far, and (after
 * grlgpu_profile_enable(ctx, 1)) per-kernel CUDA-event durations measured live on the launch stream.
 * grlgpu_profile_entry returns 1 past the last entry; model_bytes = expected DRAM bytes of the launches. */
int grlgpu_profile_enable(grlgpu_ctx* ctx, int on);
int grlgpu_profile_reset(grlgpu_ctx* ctx);
uint64_t grlgpu_launch_count(const grlgpu_ctx* ctx);
int grlgpu_profile_entry(grlgpu_ctx* ctx, int index, char* name, int name_cap, uint64_t* launches, double* total_ms, uint64_t* model_bytes);

/* self-test hooks for the device primitives the path is built from (host buffers in / out) */
int grlgpu_selftest_scan(const uint32_t* in, uint64_t n, uint64_t* out_exclusive, uint64_t* total);
int grlgpu_selftest_sort(uint64_t* keys, uint32_t* vals, uint64_t n, int n_bits);
int grlgpu_selftest_compact(const uint32_t* bits, const uint32_t* prev_bits, uint64_t n_bits, uint64_t* out, uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* GRLGPU_H */
