/*
 * TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * CPU restatement (plain C, single thread, O(n log n) comparison sorts) of grlBWT's parse phase
 * and induction phase, written from the behaviour of the reference sources, each function citing
 * the reference file:line it follows (paths relative to /root/reference).  It is the checker the
 * CUDA path is compared with in tests/, in __graft_entry__.smoke() and in bench.py's cpu_baseline.
 *
 * Parity pin: the reference ships no tests for this path ("parity unpinned" by the reference's
 * own tests).  This restatement is pinned instead against outputs of the reference
 * itself run in the build container (oracle/_ref, recipe in oracle/Makefile): per-round parse
 * dumps and .rl_bwt sha256s committed under tests/golden/ (generator: tests/golden/make_golden.py).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef uint64_t u64;

enum { /* level scalars */
    OR_N_IN = 0, OR_P = 1, OR_D = 2, OR_SUM_LEN = 3, OR_MAX_FREQ = 4, OR_TOT_PHRASES = 5, OR_ALPHABET = 6,
    OR_CELL_BYTES = 7, OR_N_PRE = 8, OR_PARSE_LEN = 9, OR_N_STR = 10, OR_LONGEST = 11
};
enum { /* level arrays (all u64) */
    OR_A_PARSE = 0, OR_A_STR_PTRS = 1, OR_A_DICT_SYMS = 2, OR_A_DICT_LEN = 3, OR_A_DICT_FREQ = 4, OR_A_DICT_META = 5,
    OR_A_PRE_SYM = 6, OR_A_PRE_LEN = 7, OR_A_RULE_L = 8, OR_A_RULE_R = 9, OR_A_HAS_HOCC = 10, OR_A_IS_SUFFIX = 11
};

typedef struct {
    u64 n_in, p, d, sum_len, max_freq, tot_phrases, alphabet, cell_bytes, n_pre, parse_len, n_str, longest;
    u64 *parse, *str_ptrs, *dict_syms, *dict_len, *dict_freq, *dict_meta, *pre_sym, *pre_len, *rule_l, *rule_r,
        *has_hocc, *is_suffix;
} level_t;

typedef struct oracle {
    /* collection stats, utils.cpp:100-189 */
    u64 n_syms, n_strings, longest, min_sym, max_sym, max_sym_freq, sep;
    int sym_bytes;
    /* current round text (parsing_info, parsing_strategies.h:11-20) */
    u64 n;        /* cells */
    u64 *v;       /* symbol values (cell>>1 from round 2 on) */
    uint8_t *rep; /* rep bit per cell (all 1 in round 1, parsing_strategies.h:102-103) */
    u64 *str_ptrs; /* n_strings+1 */
    u64 alphabet;  /* tot_phrases of the text = alphabet size */
    uint8_t *is_suffix; /* phrase_desc bit-vector */
    u64 prev_alph;
    int n_levels, cap_levels;
    level_t *lev;
    /* final BWT */
    u64 n_runs, *run_sym, *run_len, sb, fb;
} oracle_t;

static int sym_width(u64 v) { return v == 0 ? 0 : 64 - __builtin_clzll(v); } /* cdt_common.cpp:6-9 */
static u64 int_ceil(u64 a, u64 b) { return (a + b - 1) / b; }

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
static void *xcalloc(size_t n, size_t s) {
    void *p = calloc(n ? n : 1, s);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

/* ---- a1: collection_stats, external/cdt/lib/utils.cpp:100-189 ---- */
oracle_t *oracle_create(const void *text, u64 n, int sym_bytes, int *err) {
    *err = 0;
    if (n == 0 || !(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) { *err = -1; return NULL; }
    oracle_t *o = (oracle_t *)xcalloc(1, sizeof(oracle_t));
    o->sym_bytes = sym_bytes;
    o->n_syms = o->n = n;
    o->v = (u64 *)xmalloc(n * sizeof(u64));
    o->rep = (uint8_t *)xmalloc(n);
    memset(o->rep, 1, n);
    for (u64 i = 0; i < n; i++) {
        u64 s = 0;
        memcpy(&s, (const char *)text + i * sym_bytes, sym_bytes); /* little endian cells */
        o->v[i] = s;
    }
    o->sep = o->v[n - 1]; /* utils.cpp:111-114: separator = last symbol of the file */
    u64 mn = ~0ULL, mx = 0, n_str = 0, pos = 0, longest = 0;
    for (u64 i = 0; i < n; i++) {
        if (o->v[i] < mn) mn = o->v[i];
        if (o->v[i] > mx) mx = o->v[i];
        if (o->v[i] == o->sep) { n_str++; if (i + 1 - pos > longest) longest = i + 1 - pos; pos = i + 1; }
    }
    if (o->sep != mn) { *err = -2; free(o->v); free(o->rep); free(o); return NULL; } /* utils.cpp:177-180 "ill formed" */
    o->min_sym = mn; o->max_sym = mx; o->n_strings = n_str; o->longest = longest;
    o->max_sym_freq = n; /* utils.cpp:117: wide alphabets keep n_syms */
    if (sym_bytes == 1) { /* utils.cpp:161-175 */
        u64 h[256] = {0}, m = 0;
        for (u64 i = 0; i < n; i++) h[o->v[i]]++;
        for (int c = 0; c < 256; c++) if (h[c] > m) m = h[c];
        o->max_sym_freq = m;
    }
    o->str_ptrs = (u64 *)xmalloc((n_str + 1) * sizeof(u64));
    u64 k = 0; pos = 0;
    for (u64 i = 0; i < n; i++) if (o->v[i] == o->sep) { o->str_ptrs[k++] = pos; pos = i + 1; }
    o->str_ptrs[n_str] = n; /* exact_par_phase.cpp:318 */
    o->alphabet = mx + 1;   /* exact_par_phase.cpp:316 */
    o->is_suffix = (uint8_t *)xcalloc(o->alphabet, 1);
    o->is_suffix[mn] = 1;   /* exact_par_phase.cpp:311-312 */
    o->prev_alph = 0;
    return o;
}

/* ---- A.2 order: symbol by symbol; a proper prefix is GREATER (common.h:43-61, exact_LMS_induction.h:124-126) ---- */
static const u64 *g_v; /* text of the round, for the qsort comparators */
static int cmp_pg(const u64 *a, u64 la, const u64 *b, u64 lb) {
    u64 m = la < lb ? la : lb;
    for (u64 i = 0; i < m; i++) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    if (la == lb) return 0;
    return la < lb ? 1 : -1;
}
typedef struct { u64 start, len; } occ_t;
static int cmp_occ(const void *x, const void *y) {
    const occ_t *a = (const occ_t *)x, *b = (const occ_t *)y;
    int c = cmp_pg(g_v + a->start, a->len, g_v + b->start, b->len);
    if (c) return c;
    return a->start < b->start ? -1 : (a->start > b->start); /* deterministic: first occurrence first */
}
typedef struct { u64 phr, k; } sfx_t;
static const u64 *g_dsyms, *g_doff, *g_dlen;
static int cmp_sfx_content(const sfx_t *a, const sfx_t *b) {
    return cmp_pg(g_dsyms + g_doff[a->phr] + a->k, g_dlen[a->phr] - a->k, g_dsyms + g_doff[b->phr] + b->k,
                  g_dlen[b->phr] - b->k);
}
static int cmp_sfx(const void *x, const void *y) {
    const sfx_t *a = (const sfx_t *)x, *b = (const sfx_t *)y;
    int c = cmp_sfx_content(a, b);
    if (c) return c;
    if (a->phr != b->phr) return a->phr < b->phr ? -1 : 1;
    return a->k < b->k ? -1 : (a->k > b->k);
}

/* ---- one parse round: par_round, exact_par_phase.cpp:374-497 ---- */
static int oracle_round(oracle_t *o) {
    if (o->n_levels == o->cap_levels) {
        o->cap_levels = o->cap_levels ? 2 * o->cap_levels : 16;
        o->lev = (level_t *)realloc(o->lev, o->cap_levels * sizeof(level_t));
    }
    level_t *L = &o->lev[o->n_levels];
    memset(L, 0, sizeof(*L));
    const u64 n = o->n, N = o->n_strings, A = o->alphabet;
    const u64 *v = o->v;
    L->n_in = n; L->alphabet = A; L->n_str = N;

    /* A.1 phrase boundaries: lms_parsing::operator(), parsing_strategies.h:82-145.
       is_start[j]=1 for string starts and LMS breaks. */
    uint8_t *is_start = (uint8_t *)xcalloc(n + 1, 1);
    for (u64 s = 0; s < N; s++) {
        u64 st = o->str_ptrs[s], en = o->str_ptrs[s + 1] - 1;
        is_start[st] = 1;
        int type_next = 0; /* type[en] = L (0) ; S = 1 */
        for (u64 i = en; i-- > st;) {
            int t;
            if (v[i] != v[i + 1]) t = v[i] < v[i + 1]; else t = type_next;
            /* parsing_strategies.h:121-123: (type&3)==2 and (rep&3)==3 */
            if (v[i] != v[i + 1] && type_next == 1 && t == 0 && o->rep[i] && o->rep[i + 1]) is_start[i + 1] = 1;
            type_next = t;
        }
    }
    u64 p = 0;
    for (u64 i = 0; i < n; i++) p += is_start[i];
    occ_t *occ = (occ_t *)xmalloc(p * sizeof(occ_t));
    u64 *new_ptrs = (u64 *)xmalloc((N + 1) * sizeof(u64));
    {
        u64 j = 0;
        for (u64 s = 0; s < N; s++) {
            u64 st = o->str_ptrs[s], en = o->str_ptrs[s + 1] - 1;
            new_ptrs[s] = j;
            u64 b = st;
            for (u64 i = st + 1; i <= en; i++)
                if (is_start[i]) { occ[j].start = b; occ[j].len = i - b + 1; j++; b = i; } /* closed interval: shares cell i */
            occ[j].start = b; occ[j].len = en - b + 1; j++;
        }
        new_ptrs[N] = j;
    }
    free(is_start);
    L->p = p;

    /* A.2 dictionary: distinct phrases + frequencies (ext_hash_functor, exact_par_phase.hpp:33-39),
       kept here directly in A.2 order. */
    occ_t *srt = (occ_t *)xmalloc(p * sizeof(occ_t));
    memcpy(srt, occ, p * sizeof(occ_t));
    g_v = v;
    qsort(srt, p, sizeof(occ_t), cmp_occ);
    u64 d = 0, sum_len = 0;
    for (u64 i = 0; i < p; i++)
        if (i == 0 || cmp_pg(v + srt[i].start, srt[i].len, v + srt[i - 1].start, srt[i - 1].len) != 0) { d++; sum_len += srt[i].len; }
    u64 *doff = (u64 *)xmalloc((d + 1) * sizeof(u64)), *dlen = (u64 *)xmalloc(d * sizeof(u64)),
        *dfreq = (u64 *)xcalloc(d, sizeof(u64)), *dsyms = (u64 *)xmalloc(sum_len * sizeof(u64));
    {
        u64 k = 0, off = 0;
        for (u64 i = 0; i < p; i++) {
            if (i == 0 || cmp_pg(v + srt[i].start, srt[i].len, v + srt[i - 1].start, srt[i - 1].len) != 0) {
                doff[k] = off; dlen[k] = srt[i].len;
                memcpy(dsyms + off, v + srt[i].start, srt[i].len * sizeof(u64));
                off += srt[i].len; k++;
            }
            dfreq[k - 1]++;
        }
        doff[d] = off;
    }
    u64 max_freq = 0;
    for (u64 i = 0; i < d; i++) if (dfreq[i] > max_freq) max_freq = dfreq[i];
    L->d = d; L->sum_len = sum_len; L->max_freq = max_freq;

    /* A.3 ranks among unsolved blocks + preliminary BWT: produce_pre_bwt, exact_par_phase.cpp:136-242 */
    sfx_t *sf = (sfx_t *)xmalloc(sum_len * sizeof(sfx_t));
    u64 ns = 0;
    for (u64 i = 0; i < d; i++)
        for (u64 k = 0; k < dlen[i]; k++) {
            if (dlen[i] - k == 1 && !o->is_suffix[dsyms[doff[i] + k]]) continue; /* :163 invalid suffix */
            sf[ns].phr = i; sf[ns].k = k; ns++;
        }
    g_dsyms = dsyms; g_doff = doff; g_dlen = dlen;
    qsort(sf, ns, sizeof(sfx_t), cmp_sfx);

    const u64 bwt_dummy = A + 1, hocc_dummy = A + 2; /* exact_par_phase.hpp:113-115 */
    u64 *meta = (u64 *)xcalloc(d, sizeof(u64));
    u64 *hocc_rank_at = (u64 *)xmalloc(sum_len * sizeof(u64)); /* phr_marks + new_phrases_ht, :190-207 */
    memset(hocc_rank_at, 0xff, sum_len * sizeof(u64));
    u64 *rep_phr = (u64 *)xmalloc((ns + 1) * sizeof(u64)), *rep_k = (u64 *)xmalloc((ns + 1) * sizeof(u64));
    uint8_t *hh = (uint8_t *)xcalloc(ns + 1, 1);
    u64 *pre_sym = (u64 *)xmalloc((ns + 1) * sizeof(u64)), *pre_len = (u64 *)xmalloc((ns + 1) * sizeof(u64));
    u64 n_pre = 0, rank = 0;
    for (u64 g0 = 0; g0 < ns;) {
        u64 g1 = g0 + 1;
        while (g1 < ns && cmp_sfx_content(&sf[g1], &sf[g0]) == 0) g1++;
        int full = 0, multi = 0;
        u64 acc = 0, first_left = 0;
        for (u64 e = g0; e < g1; e++) {
            u64 left = sf[e].k == 0 ? bwt_dummy : dsyms[doff[sf[e].phr] + sf[e].k - 1];
            if (sf[e].k == 0) full = 1;
            if (e == g0) first_left = left; else if (left != first_left) multi = 1;
            acc += dfreq[sf[e].phr];
        }
        u64 sym;
        if (full || multi) { /* :187 */
            if (g1 - g0 > 1) {
                hh[rank] = 1; sym = hocc_dummy;
                for (u64 e = g0; e < g1; e++) hocc_rank_at[doff[sf[e].phr] + sf[e].k] = rank;
            } else sym = bwt_dummy;
            for (u64 e = g0; e < g1; e++)
                if (sf[e].k == 0) meta[sf[e].phr] = (rank << 1) | (dfreq[sf[e].phr] > 1); /* :174-176 */
            rep_phr[rank] = sf[g0].phr; rep_k[rank] = sf[g0].k;
            rank++;
        } else sym = first_left;
        /* canonical form here: maximal runs (the reference never merges its first two runs, :212-216;
           that only changes the private pre_bwt file, not the induced BWT) */
        if (n_pre > 0 && pre_sym[n_pre - 1] == sym) pre_len[n_pre - 1] += acc;
        else { pre_sym[n_pre] = sym; pre_len[n_pre] = acc; n_pre++; }
        g0 = g1;
    }
    free(sf);
    const u64 tot = rank;
    L->tot_phrases = tot; L->n_pre = n_pre; L->pre_sym = pre_sym; L->pre_len = pre_len;

    /* A.5 grammar rules: produce_grammar, exact_par_phase.cpp:14-95 */
    const u64 alph3 = A + 3, metasym_dummy = alph3 + tot + 1; /* :19-20 */
    u64 *rule_l = (u64 *)xmalloc((tot + 1) * sizeof(u64)), *rule_r = (u64 *)xmalloc((tot + 1) * sizeof(u64)),
        *has_hocc = (u64 *)xmalloc((tot + 1) * sizeof(u64));
    for (u64 u = 0; u < tot; u++) {
        has_hocc[u] = hh[u];
        const u64 *P = dsyms + doff[rep_phr[u]];
        u64 len = dlen[rep_phr[u]], pos = rep_k[u];
        if (pos == len - 1) { rule_l[u] = metasym_dummy; rule_r[u] = P[pos]; continue; } /* :38-41 */
        pos++;
        while (hocc_rank_at[doff[rep_phr[u]] + pos] == ~0ULL && pos != len - 1) pos++; /* :43-44 */
        u64 l_sym = P[pos - 1];
        u64 hr = hocc_rank_at[doff[rep_phr[u]] + pos];
        if (hr != ~0ULL) { rule_l[u] = l_sym; rule_r[u] = alph3 + hr; } /* :49-80 */
        else { u64 r_sym = P[pos]; rule_l[u] = metasym_dummy; rule_r[u] = o->is_suffix[r_sym] ? r_sym : l_sym; } /* :81-85 */
    }
    free(hh); free(rep_phr); free(rep_k); free(hocc_rank_at);
    L->rule_l = rule_l; L->rule_r = rule_r; L->has_hocc = has_hocc;

    /* a9 metasymbol assignment: exact_par_phase.cpp:427-450 */
    uint8_t *new_is_suffix = (uint8_t *)xcalloc(tot + 1, 1);
    for (u64 i = 0; i < d; i++) new_is_suffix[meta[i] >> 1] = o->is_suffix[dsyms[doff[i] + dlen[i] - 1]];
    L->is_suffix = (u64 *)xmalloc((tot + 1) * sizeof(u64));
    for (u64 i = 0; i < tot; i++) L->is_suffix[i] = new_is_suffix[i];

    /* A.4 rewrite: ext_parse_functor exact_par_phase.hpp:53-59, parse_text parsing_strategies.h:644-676 */
    u64 *occ_meta = (u64 *)xmalloc(p * sizeof(u64));
    {   /* occurrence -> distinct phrase by binary search in the sorted dictionary */
        for (u64 j = 0; j < p; j++) {
            u64 lo = 0, hi = d;
            while (lo < hi) {
                u64 mid = (lo + hi) / 2;
                int c = cmp_pg(dsyms + doff[mid], dlen[mid], v + occ[j].start, occ[j].len);
                if (c < 0) lo = mid + 1; else hi = mid;
            }
            occ_meta[j] = meta[lo];
        }
    }
    u64 bps = (u64)sym_width(tot) + 1; /* exact_par_phase.cpp:456-465 */
    L->cell_bytes = bps <= 8 ? 1 : bps <= 16 ? 2 : bps <= 32 ? 4 : 8;
    L->parse_len = p;
    L->parse = occ_meta;
    L->str_ptrs = new_ptrs;
    L->dict_syms = dsyms; L->dict_len = dlen; L->dict_freq = dfreq; L->dict_meta = meta;
    u64 longest = 0;
    for (u64 s = 0; s < N; s++) if (new_ptrs[s + 1] - new_ptrs[s] > longest) longest = new_ptrs[s + 1] - new_ptrs[s];
    L->longest = longest;
    free(doff); free(occ); free(srt);

    /* next round's text */
    free(o->v); free(o->rep); free(o->is_suffix); free(o->str_ptrs);
    o->n = p;
    o->v = (u64 *)xmalloc(p * sizeof(u64));
    o->rep = (uint8_t *)xmalloc(p);
    for (u64 j = 0; j < p; j++) { o->v[j] = occ_meta[j] >> 1; o->rep[j] = (uint8_t)(occ_meta[j] & 1); }
    o->str_ptrs = (u64 *)xmalloc((N + 1) * sizeof(u64));
    memcpy(o->str_ptrs, new_ptrs, (N + 1) * sizeof(u64));
    o->is_suffix = new_is_suffix;
    o->prev_alph = alph3;   /* exact_par_phase.cpp:420 */
    o->alphabet = tot;
    o->n_levels++;
    return p == N; /* exact_par_phase.cpp:496 */
}

/* par_phase, exact_par_phase.cpp:285-372: rounds until every string is one cell. Returns number of rounds. */
int oracle_par_phase(oracle_t *o) {
    while (!oracle_round(o)) {}
    return o->n_levels;
}

u64 oracle_level_scalar(const oracle_t *o, int level, int what) {
    if (level < 0 || level >= o->n_levels) return ~0ULL;
    const level_t *L = &o->lev[level];
    const u64 tab[] = {L->n_in, L->p, L->d, L->sum_len, L->max_freq, L->tot_phrases, L->alphabet, L->cell_bytes,
                       L->n_pre, L->parse_len, L->n_str, L->longest};
    return (what >= 0 && what < 12) ? tab[what] : ~0ULL;
}
const u64 *oracle_level_array(const oracle_t *o, int level, int what, u64 *count) {
    if (level < 0 || level >= o->n_levels) return NULL;
    const level_t *L = &o->lev[level];
    switch (what) {
        case OR_A_PARSE: *count = L->parse_len; return L->parse;
        case OR_A_STR_PTRS: *count = L->n_str + 1; return L->str_ptrs;
        case OR_A_DICT_SYMS: *count = L->sum_len; return L->dict_syms;
        case OR_A_DICT_LEN: *count = L->d; return L->dict_len;
        case OR_A_DICT_FREQ: *count = L->d; return L->dict_freq;
        case OR_A_DICT_META: *count = L->d; return L->dict_meta;
        case OR_A_PRE_SYM: *count = L->n_pre; return L->pre_sym;
        case OR_A_PRE_LEN: *count = L->n_pre; return L->pre_len;
        case OR_A_RULE_L: *count = L->tot_phrases; return L->rule_l;
        case OR_A_RULE_R: *count = L->tot_phrases; return L->rule_r;
        case OR_A_HAS_HOCC: *count = L->tot_phrases; return L->has_hocc;
        case OR_A_IS_SUFFIX: *count = L->tot_phrases; return L->is_suffix;
    }
    return NULL;
}
u64 oracle_stat(const oracle_t *o, int what) {
    const u64 tab[] = {o->n_syms, o->n_strings, o->longest, o->min_sym, o->max_sym, o->max_sym_freq, o->sep};
    return (what >= 0 && what < 7) ? tab[what] : ~0ULL;
}

/* ---- run list helper ---- */
typedef struct { u64 *sym, *len, n, cap; } runs_t;
static void runs_push(runs_t *r, u64 s, u64 l) { /* merges equal neighbours: bwt_io.h:448-498 users */
    if (l == 0) return;
    if (r->n && r->sym[r->n - 1] == s) { r->len[r->n - 1] += l; return; }
    if (r->n == r->cap) {
        r->cap = r->cap ? 2 * r->cap : 1024;
        r->sym = (u64 *)realloc(r->sym, r->cap * sizeof(u64));
        r->len = (u64 *)realloc(r->len, r->cap * sizeof(u64));
    }
    r->sym[r->n] = s; r->len[r->n] = l; r->n++;
}
static void runs_push_raw(runs_t *r, u64 s, u64 l) { /* no merging (hocc buckets keep FROM_BWT marks apart) */
    if (r->n == r->cap) {
        r->cap = r->cap ? 2 * r->cap : 1024;
        r->sym = (u64 *)realloc(r->sym, r->cap * sizeof(u64));
        r->len = (u64 *)realloc(r->len, r->cap * sizeof(u64));
    }
    r->sym[r->n] = s; r->len[r->n] = l; r->n++;
}

/* ---- induction phase: ind_phase, exact_ind_phase.cpp:674-697 (parse2bwt :603-672, infer_lvl_bwt :111-386/:388-601) ---- */
int oracle_ind_phase(oracle_t *o) {
    if (o->n_levels == 0) return -1;
    const int R = o->n_levels;
    runs_t bwt = {0};
    { /* deepest level: final parse (>>1) in string order, run-length encoded (parse2bwt_int :621-635) */
        const level_t *L = &o->lev[R - 1];
        for (u64 i = 0; i < L->parse_len; i++) runs_push(&bwt, L->parse[i] >> 1, 1);
    }
    for (int lv = R - 1; lv >= 0; lv--) {
        const level_t *L = &o->lev[lv];
        const u64 A = L->alphabet, alph3 = A + 3, bwt_dummy = A + 1, hocc_dummy = A + 2, tot = L->tot_phrases;
        const u64 FROM_BWT = ~0ULL;
        /* step 1 (:143-258): per run of BWT_{i+1}, feed the hocc buckets along the grammar chain and
           replace the run's symbol by the terminal of the chain */
        runs_t *bucket = (runs_t *)xcalloc(tot, sizeof(runs_t));
        for (u64 i = 0; i < bwt.n; i++) {
            u64 P = bwt.sym[i], f = bwt.len[i];
            if (L->has_hocc[P]) {
                runs_t *b = &bucket[P];
                if (b->n && b->sym[b->n - 1] == FROM_BWT) b->len[b->n - 1] += f; else runs_push_raw(b, FROM_BWT, f);
            }
            u64 l = L->rule_l[P], r = L->rule_r[P];
            while (r >= alph3) {
                u64 g = r - alph3;
                runs_t *b = &bucket[g];
                if (b->n && b->sym[b->n - 1] == l) b->len[b->n - 1] += f; else runs_push_raw(b, l, f);
                l = L->rule_l[g]; r = L->rule_r[g];
            }
            bwt.sym[i] = r;
        }
        /* step 2 (:287-361): assemble along the preliminary BWT */
        runs_t out = {0};
        u64 sp = 0;  /* stream pointer in the rewritten BWT_{i+1} */
        u64 hb = 0, hi = 0; /* bucket index, index inside bucket */
#define TAKE(F) do { u64 _f = (F); while (_f) { u64 _t = bwt.len[sp] < _f ? bwt.len[sp] : _f; \
            runs_push(&out, bwt.sym[sp], _t); _f -= _t; bwt.len[sp] -= _t; if (bwt.len[sp] == 0) sp++; } } while (0)
        for (u64 i = 0; i < L->n_pre; i++) {
            u64 s = L->pre_sym[i], f = L->pre_len[i];
            if (s == bwt_dummy) { TAKE(f); }
            else if (s == hocc_dummy) {
                while (f) {
                    while (hb < tot && hi == bucket[hb].n) { hb++; hi = 0; }
                    runs_t *b = &bucket[hb];
                    u64 t = b->len[hi] < f ? b->len[hi] : f;
                    if (b->sym[hi] == FROM_BWT) { TAKE(t); } else runs_push(&out, b->sym[hi], t);
                    b->len[hi] -= t; f -= t;
                    if (b->len[hi] == 0) hi++;
                }
            } else runs_push(&out, s, f);
        }
#undef TAKE
        for (u64 g = 0; g < tot; g++) { free(bucket[g].sym); free(bucket[g].len); }
        free(bucket);
        free(bwt.sym); free(bwt.len);
        bwt = out;
    }
    o->n_runs = bwt.n; o->run_sym = bwt.sym; o->run_len = bwt.len;
    /* header widths of bwt_lev_0 (:274-276 with level-0 dictionary values) */
    o->sb = int_ceil((u64)sym_width(o->max_sym + 1 + 3), 8);
    o->fb = int_ceil((u64)sym_width(o->max_sym_freq), 8);
    return 0;
}

u64 oracle_bwt_runs(const oracle_t *o, const u64 **syms, const u64 **lens, u64 *sb, u64 *fb) {
    *syms = o->run_sym; *lens = o->run_len; *sb = o->sb; *fb = o->fb;
    return o->n_runs;
}

/* .rl_bwt bytes: bwt_io.h:377-382 (header), :448-490 (records); SURVEY.md App. C. Returns bytes written. */
u64 oracle_write_rl_bwt(const oracle_t *o, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return 0;
    fwrite(&o->sb, 8, 1, f); fwrite(&o->fb, 8, 1, f);
    for (u64 i = 0; i < o->n_runs; i++) { fwrite(&o->run_sym[i], o->sb, 1, f); fwrite(&o->run_len[i], o->fb, 1, f); }
    fclose(f);
    return 16 + o->n_runs * (o->sb + o->fb);
}

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    for (int i = 0; i < o->n_levels; i++) {
        level_t *L = &o->lev[i];
        free(L->parse); free(L->str_ptrs); free(L->dict_syms); free(L->dict_len); free(L->dict_freq); free(L->dict_meta);
        free(L->pre_sym); free(L->pre_len); free(L->rule_l); free(L->rule_r); free(L->has_hocc); free(L->is_suffix);
    }
    free(o->lev); free(o->v); free(o->rep); free(o->str_ptrs); free(o->is_suffix); free(o->run_sym); free(o->run_len);
    free(o);
}
