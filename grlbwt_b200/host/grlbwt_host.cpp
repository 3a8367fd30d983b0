// C entry points of the host side (include/grlbwt.h): whole BCR BWT construction for callers that
// cannot include the C++ templates of grl_bwt.hpp (tests, bench.py through ctypes).
#include "../../include/grlbwt.h"

#include <cstdlib>
#include <cstring>
#include <string>

#include "grl_bwt.hpp"

static thread_local std::string g_last_error;
static thread_local std::vector<grlbwt::RoundDigest> g_last_digests;
static thread_local uint64_t g_last_exchange_bytes = 0;
static thread_local std::string g_last_comm;

static int fill_result(grlbwt::BwtResult& r, grlbwt_result_t* out) {
    const size_t nr = r.n_runs();
    out->n_runs = nr;
    out->sb = r.sb;
    out->fb = r.fb;
    out->syms = (uint64_t*)malloc((nr ? nr : 1) * sizeof(uint64_t));
    out->lens = (uint64_t*)malloc((nr ? nr : 1) * sizeof(uint64_t));
    if (!out->syms || !out->lens) { free(out->syms); free(out->lens); out->syms = out->lens = nullptr; g_last_error = "out of host memory"; return -100; }
    if (r.narrow) {
        for (size_t k = 0; k < nr; k++) out->syms[k] = r.runs32.sym[k];
        if (r.runs32.len32.empty()) memcpy(out->lens, r.runs32.len.data(), nr * sizeof(uint64_t));
        else for (size_t k = 0; k < nr; k++) out->lens[k] = r.runs32.len32[k];
    } else {
        memcpy(out->syms, r.runs.sym.data(), nr * sizeof(uint64_t));
        memcpy(out->lens, r.runs.len.data(), nr * sizeof(uint64_t));
    }
    out->n_rounds = r.parse.rounds.size();
    out->h2d_ms = r.parse.h2d_ms;
    out->par_phase_ms = r.parse.par_ms;
    out->ind_phase_ms = r.ind_ms;
    out->induced_on_device = r.parse.induced_on_device ? 1 : 0;
    for (const auto& rd : r.parse.rounds) { out->device_ms += rd.device_ms; out->algorithmic_bytes += rd.algorithmic_bytes; }
    g_last_digests = r.parse.digests;
    g_last_exchange_bytes = r.parse.exchange_bytes;
    g_last_comm = r.parse.comm_kind;
    return GRLGPU_OK;
}

extern "C" {

int grlbwt_build(const void* text, uint64_t n_syms, int sym_bytes, int device, int n_threads, int verbose, grlbwt_result_t* out) {
    if (!text || !out || n_syms == 0) return GRLGPU_ERR_ARG;
    if (!(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) return GRLGPU_ERR_ARG;
    memset(out, 0, sizeof(*out));
    try {
        grlbwt::BwtResult r = grlbwt::build_bwt(text, n_syms, sym_bytes, device, (size_t)(n_threads > 0 ? n_threads : 1), verbose != 0);
        return fill_result(r, out);
    } catch (const grlbwt::GpuError& e) {
        g_last_error = e.what();
        return e.status;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

int grlbwt_build_mg(const void* text, uint64_t n_syms, int sym_bytes, const int* devices, int n_ranks, int n_threads, int comm_kind, int verbose,
                    grlbwt_result_t* out) {
    if (!text || !out || n_syms == 0 || !devices || n_ranks < 1 || n_ranks > 31) return GRLGPU_ERR_ARG;
    if (!(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) return GRLGPU_ERR_ARG;
    memset(out, 0, sizeof(*out));
    try {
        grlbwt::TextSource src;
        src.mem = (const unsigned char*)text;
        src.bytes = n_syms * (uint64_t)sym_bytes;
        grlbwt::BwtResult r = grlbwt::build_bwt(src, sym_bytes, std::vector<int>(devices, devices + n_ranks), comm_kind, (size_t)(n_threads > 0 ? n_threads : 1), verbose != 0);
        return fill_result(r, out);
    } catch (const grlbwt::GpuError& e) {
        g_last_error = e.what();
        return e.status;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

// whole construction with the level-0 BWT delivered into caller-owned 32-bit arrays (see include/grlbwt.h)
int grlbwt_build_to(const void* text, uint64_t n_syms, int sym_bytes, const int* devices, int n_ranks, int n_threads, int comm_kind, uint32_t* out_syms,
                    uint32_t* out_lens, uint64_t cap_runs, grlbwt_result_t* out) {
    if (!text || !out || n_syms == 0 || !devices || n_ranks < 1 || n_ranks > 31 || !out_syms || !out_lens) return GRLGPU_ERR_ARG;
    if (!(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) return GRLGPU_ERR_ARG;
    memset(out, 0, sizeof(*out));
    try {
        grlbwt::TextSource src;
        src.mem = (const unsigned char*)text;
        src.bytes = n_syms * (uint64_t)sym_bytes;
        grlbwt::OutputBuffers ob;
        ob.sym = out_syms; ob.len = out_lens; ob.cap = cap_runs;
        grlbwt::BwtResult r = grlbwt::build_bwt(src, sym_bytes, std::vector<int>(devices, devices + n_ranks), comm_kind, (size_t)(n_threads > 0 ? n_threads : 1), false, &ob);
        const size_t nr = r.n_runs();
        out->n_runs = nr; out->sb = r.sb; out->fb = r.fb;
        out->n_rounds = r.parse.rounds.size();
        out->h2d_ms = r.parse.h2d_ms; out->par_phase_ms = r.parse.par_ms; out->ind_phase_ms = r.ind_ms;
        out->induced_on_device = r.parse.induced_on_device ? 1 : 0;
        for (const auto& rd : r.parse.rounds) { out->device_ms += rd.device_ms; out->algorithmic_bytes += rd.algorithmic_bytes; }
        g_last_digests = r.parse.digests;
        g_last_exchange_bytes = r.parse.exchange_bytes;
        g_last_comm = r.parse.comm_kind;
        if (!r.parse.bwt_in_caller_buffers) {  // host induction (or too many runs for the buffers): copy what fits the 32-bit arrays
            if (nr > cap_runs) { g_last_error = "the output buffers are too small: " + std::to_string(nr) + " runs"; return GRLGPU_ERR_LIMIT; }
            for (size_t k = 0; k < nr; k++) {
                const uint64_t sy = r.narrow ? (uint64_t)r.runs32.sym[k] : r.runs.sym[k], ln = r.narrow ? r.runs32.length(k) : r.runs.len[k];
                if (sy >> 32 || ln >> 32) { g_last_error = "the BWT does not fit 32-bit symbols and lengths"; return GRLGPU_ERR_LIMIT; }
                out_syms[k] = (uint32_t)sy;
                out_lens[k] = (uint32_t)ln;
            }
        }
        return GRLGPU_OK;
    } catch (const grlbwt::GpuError& e) {
        g_last_error = e.what();
        return e.status;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

// whole construction with the level-0 BWT delivered as the image of the .rl_bwt file (see include/grlbwt.h)
int grlbwt_build_packed(const void* text, uint64_t n_syms, int sym_bytes, const int* devices, int n_ranks, int n_threads, int comm_kind, void* out_image,
                        uint64_t cap_bytes, uint64_t* image_bytes, grlbwt_result_t* out) {
    if (!text || !out || n_syms == 0 || !devices || n_ranks < 1 || n_ranks > 31 || !out_image || cap_bytes < 16) return GRLGPU_ERR_ARG;
    if (!(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) return GRLGPU_ERR_ARG;
    memset(out, 0, sizeof(*out));
    if (image_bytes) *image_bytes = 0;
    try {
        grlbwt::TextSource src;
        src.mem = (const unsigned char*)text;
        src.bytes = n_syms * (uint64_t)sym_bytes;
        grlbwt::OutputBuffers ob;
        ob.packed = (unsigned char*)out_image; ob.packed_cap = cap_bytes;
        grlbwt::BwtResult r = grlbwt::build_bwt(src, sym_bytes, std::vector<int>(devices, devices + n_ranks), comm_kind, (size_t)(n_threads > 0 ? n_threads : 1), false, &ob);
        const size_t nr = r.n_runs();
        out->n_runs = nr; out->sb = r.sb; out->fb = r.fb;
        out->n_rounds = r.parse.rounds.size();
        out->h2d_ms = r.parse.h2d_ms; out->par_phase_ms = r.parse.par_ms; out->ind_phase_ms = r.ind_ms;
        out->induced_on_device = r.parse.induced_on_device ? 1 : 0;
        for (const auto& rd : r.parse.rounds) { out->device_ms += rd.device_ms; out->algorithmic_bytes += rd.algorithmic_bytes; }
        g_last_digests = r.parse.digests;
        g_last_exchange_bytes = r.parse.exchange_bytes;
        g_last_comm = r.parse.comm_kind;
        const uint64_t need = 16 + (uint64_t)nr * (r.sb + r.fb);
        if (image_bytes) *image_bytes = need;
        if (!r.parse.bwt_in_caller_buffers) {  // host induction: pack here
            if (need > cap_bytes) { g_last_error = "the output buffer is too small: the .rl_bwt image has " + std::to_string(need) + " bytes"; return GRLGPU_ERR_LIMIT; }
            if (r.narrow && !r.runs32.len32.empty()) grlbwt::pack_rl_bwt((unsigned char*)out_image, r.runs32.sym.data(), r.runs32.len32.data(), nr, r.sb, r.fb);
            else if (r.narrow) grlbwt::pack_rl_bwt((unsigned char*)out_image, r.runs32.sym.data(), r.runs32.len.data(), nr, r.sb, r.fb);
            else grlbwt::pack_rl_bwt((unsigned char*)out_image, r.runs.sym.data(), r.runs.len.data(), nr, r.sb, r.fb);
        }
        return GRLGPU_OK;
    } catch (const grlbwt::GpuError& e) {
        g_last_error = e.what();
        return e.status;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

uint64_t grlbwt_last_digests(uint64_t* out, uint64_t cap_rounds) {
    const uint64_t n = g_last_digests.size();
    for (uint64_t i = 0; i < n && i < cap_rounds && out; i++) {
        const grlbwt::RoundDigest& d = g_last_digests[i];
        uint64_t* o = out + i * 9;
        o[0] = d.tot; o[1] = d.n_pre; o[2] = d.parse_len; o[3] = d.n_phrases; o[4] = d.dict_syms;
        for (int k = 0; k < 4; k++) o[5 + k] = d.cs[k];
    }
    return n;
}
uint64_t grlbwt_last_exchange_bytes(void) { return g_last_exchange_bytes; }
const char* grlbwt_last_comm(void) { return g_last_comm.c_str(); }

void grlbwt_free_result(grlbwt_result_t* r) {
    if (!r) return;
    free(r->syms);
    free(r->lens);
    r->syms = r->lens = nullptr;
    r->n_runs = 0;
}

int grlbwt_build_file(const char* input_file, const char* output_file, int sym_bytes, int device, int n_threads, int verbose) {
    if (!input_file || !output_file) return GRLGPU_ERR_ARG;
    if (!(sym_bytes == 1 || sym_bytes == 2 || sym_bytes == 4 || sym_bytes == 8)) return GRLGPU_ERR_ARG;
    try {
        grlbwt::TextSource src;
        src.file = input_file;
        src.bytes = grlbwt::file_size_of(input_file);
        if (src.bytes == 0 || src.bytes % (uint64_t)sym_bytes) return GRLGPU_ERR_ILL_FORMED;
        grlbwt::BwtResult r = grlbwt::build_bwt(src, sym_bytes, std::vector<int>{device}, grlbwt::COMM_AUTO, (size_t)(n_threads > 0 ? n_threads : 1), verbose != 0);
        r.write(output_file);
        return GRLGPU_OK;
    } catch (const grlbwt::GpuError& e) {
        g_last_error = e.what();
        return e.status;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

int grlbwt_selftest_write(const char* path, const uint64_t* syms, const uint64_t* lens, uint64_t n_runs, uint64_t sb, uint64_t fb, int narrow) {
    try {
        if (!path || sb == 0 || sb > 8 || fb == 0 || fb > 8) throw std::runtime_error("bad arguments");
        if (narrow) {
            std::vector<uint32_t> s32(syms, syms + n_runs);
            grlbwt::write_rl_bwt(path, s32.data(), lens, n_runs, sb, fb);
        } else grlbwt::write_rl_bwt(path, syms, lens, n_runs, sb, fb);
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

// the in-memory packer grlbwt_build_packed uses when the induction ran on the host (CPU-only self test)
int grlbwt_selftest_pack(void* out_image, uint64_t cap_bytes, const uint64_t* syms, const uint64_t* lens, uint64_t n_runs, uint64_t sb, uint64_t fb, int narrow) {
    try {
        if (!out_image || sb == 0 || sb > 8 || fb == 0 || fb > 8 || cap_bytes < 16 + n_runs * (sb + fb)) throw std::runtime_error("bad arguments");
        if (narrow) {
            std::vector<uint32_t> s32(syms, syms + n_runs), l32(lens, lens + n_runs);
            grlbwt::pack_rl_bwt((unsigned char*)out_image, s32.data(), l32.data(), n_runs, sb, fb);
        } else grlbwt::pack_rl_bwt((unsigned char*)out_image, syms, lens, n_runs, sb, fb);
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

const char* grlbwt_last_error(void) { return g_last_error.c_str(); }

// the shard boundaries a multi-GPU run would use (CPU-only self test of gpu_par_phase.hpp: shard_bounds)
int grlbwt_selftest_shard_bounds(const void* text, uint64_t n_syms, int sym_bytes, int n_ranks, uint64_t* bounds_out, int* n_shards) {
    try {
        if (!text || !bounds_out || !n_shards || n_syms == 0 || n_ranks < 1) throw std::runtime_error("bad arguments");
        grlbwt::TextSource src;
        src.mem = (const unsigned char*)text;
        src.bytes = n_syms * (uint64_t)sym_bytes;
        const std::vector<uint64_t> b = grlbwt::shard_bounds(src, sym_bytes, n_ranks);
        for (size_t i = 0; i < b.size(); i++) bounds_out[i] = b[i];
        *n_shards = (int)b.size() - 1;
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

// host induction alone, from caller-provided level artefacts (CPU-only self test of ind_phase.hpp)
int grlbwt_selftest_induce(int n_levels, const uint64_t* alphabet, const uint64_t* tot, const uint64_t* const* rule_l, const uint64_t* const* rule_r,
                           const uint8_t* const* has_hocc, const uint64_t* n_pre, const uint64_t* const* pre_sym, const uint64_t* const* pre_len,
                           const uint64_t* final_parse, uint64_t n_strings, int n_threads, grlbwt_result_t* out) {
    try {
        if (n_threads > 0) {  // multi-threaded 32-bit path (ind_phase_mt.hpp)
            std::vector<grlbwt::Level32> lv((size_t)n_levels);
            for (int i = 0; i < n_levels; i++) {
                grlbwt::Level32& L = lv[(size_t)i];
                L.alphabet = alphabet[i];
                L.tot_phrases = tot[i];
                L.rule_l.assign(rule_l[i], rule_l[i] + tot[i]);
                L.rule_r.assign(rule_r[i], rule_r[i] + tot[i]);
                L.has_hocc.assign(has_hocc[i], has_hocc[i] + tot[i]);
                L.pre_sym.assign(pre_sym[i], pre_sym[i] + n_pre[i]);
                // both storage widths of the run lengths get exercised: odd levels keep them in 32 bits when they fit (what
                // gpu_par_phase does for every level that came through grlgpu_fetch_level32)
                bool fits = (i & 1) != 0;
                for (uint64_t k = 0; fits && k < n_pre[i]; k++) fits = pre_len[i][k] < (1ull << 32);
                if (fits) L.pre_len32.assign(pre_len[i], pre_len[i] + n_pre[i]);
                else L.pre_len.assign(pre_len[i], pre_len[i] + n_pre[i]);
            }
            grlbwt::RunArr r = grlbwt::ind_phase_mt(lv, final_parse, n_strings, (size_t)n_threads);
            memset(out, 0, sizeof(*out));
            out->n_runs = r.size();
            out->syms = (uint64_t*)malloc((r.size() ? r.size() : 1) * sizeof(uint64_t));
            out->lens = (uint64_t*)malloc((r.size() ? r.size() : 1) * sizeof(uint64_t));
            for (size_t k = 0; k < r.size(); k++) { out->syms[k] = r.sym[k]; out->lens[k] = r.len[k]; }
            return 0;
        }
        std::vector<grlbwt::Level> levels((size_t)n_levels);
        for (int i = 0; i < n_levels; i++) {
            grlbwt::Level& L = levels[(size_t)i];
            L.alphabet = alphabet[i];
            L.tot_phrases = tot[i];
            L.rule_l.assign(rule_l[i], rule_l[i] + tot[i]);
            L.rule_r.assign(rule_r[i], rule_r[i] + tot[i]);
            L.has_hocc.assign(has_hocc[i], has_hocc[i] + tot[i]);
            L.pre_sym.assign(pre_sym[i], pre_sym[i] + n_pre[i]);
            L.pre_len.assign(pre_len[i], pre_len[i] + n_pre[i]);
        }
        grlbwt::RunList r = grlbwt::ind_phase<uint64_t>(levels, final_parse, n_strings);
        memset(out, 0, sizeof(*out));
        out->n_runs = r.size();
        out->syms = (uint64_t*)malloc((r.size() ? r.size() : 1) * sizeof(uint64_t));
        out->lens = (uint64_t*)malloc((r.size() ? r.size() : 1) * sizeof(uint64_t));
        memcpy(out->syms, r.sym.data(), r.size() * sizeof(uint64_t));
        memcpy(out->lens, r.len.data(), r.size() * sizeof(uint64_t));
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return -100;
    }
}

}  // extern "C"
