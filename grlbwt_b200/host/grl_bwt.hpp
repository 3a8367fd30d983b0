// Entry point of the tool, same shape as the reference's include/grl_bwt.hpp:23-79:
//   grl_bwt_algo<sym_type, false>(i_file, o_file, tmp_ws, n_threads, hbuff_frac, b_p_r)
// = parse phase (here: on the B200 through the C ABI of include/grlgpu.h, replacing
// exact_algo::par_phase, lib/exact_algo/exact_par_phase.cpp:285-372) + induction phase
// (host, ind_phase.hpp) + .rl_bwt output (rl_bwt_io.hpp). No CPU fallback for the parse phase:
// if the device library reports an error the call throws.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/grlgpu.h"
#include "ind_phase.hpp"
#include "ind_phase_mt.hpp"
#include "rl_bwt_io.hpp"

// same role as the reference's tmp_workspace (external/cdt/include/utils.h:52-104): a private folder
// under -T that is removed on destruction. The device path keeps the per-level artefacts in memory,
// so the folder only hosts the output file while it is being written.
struct tmp_workspace {
    std::string tmp_folder;
    bool remove_all;
    explicit tmp_workspace(const std::string& base_folder = std::filesystem::temp_directory_path().string(), bool rem_all = true,
                           const std::string& prefix = "tmp")
        : remove_all(rem_all) {
        std::string tmpl = (std::filesystem::canonical(std::filesystem::path(base_folder)) / (prefix + ".XXXXXX")).string();
        std::vector<char> buf(tmpl.begin(), tmpl.end());
        buf.push_back('\0');
        if (mkdtemp(buf.data()) == nullptr) throw std::runtime_error("Error trying to create a temporal folder");
        tmp_folder = buf.data();
    }
    tmp_workspace(const tmp_workspace&) = delete;
    std::string get_file(const std::string& name) const { return (std::filesystem::path(tmp_folder) / name).string(); }
    std::string folder() const { return tmp_folder; }
    ~tmp_workspace() {
        if (remove_all) {
            std::error_code ec;
            std::filesystem::remove_all(tmp_folder, ec);
        }
    }
};

namespace grlbwt {

struct ParseResult {
    grlgpu_stats_t stats{};
    std::vector<Level> levels;        // 64-bit symbols: only filled when some level needs them
    std::vector<Level32> levels32;    // 32-bit symbols (the usual case), consumed by the multi-threaded induction
    bool wide = false;
    std::vector<grlgpu_round_t> rounds;
    std::vector<uint64_t> final_parse;  // one cell per string, cells = rank<<1|rep
    double h2d_ms = 0, par_ms = 0;
};

struct GpuError : std::runtime_error {
    int status;
    GpuError(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};

inline double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// The parse phase on the device: rounds until every string is one metasymbol
// (replaces exact_algo::par_phase<sym_type>, exact_par_phase.cpp:285-372).
inline ParseResult gpu_par_phase(const void* text, uint64_t n_syms, int sym_bytes, int device, bool verbose) {
    ParseResult res;
    grlgpu_ctx* ctx = nullptr;
    int rc = grlgpu_create(&ctx, device, 0);
    if (rc != GRLGPU_OK) throw GpuError(rc, std::string("grlgpu_create: ") + grlgpu_strerror(rc));
    auto fail = [&](const char* what, int st) {
        std::string m = std::string(what) + ": " + grlgpu_strerror(st) + " (" + grlgpu_last_error(ctx) + ")";
        grlgpu_destroy(ctx);
        throw GpuError(st, m);
    };
    auto t0 = std::chrono::steady_clock::now();
    if ((rc = grlgpu_set_text(ctx, text, n_syms, sym_bytes)) != GRLGPU_OK) fail("grlgpu_set_text", rc);
    res.h2d_ms = ms_since(t0);
    t0 = std::chrono::steady_clock::now();
    if ((rc = grlgpu_stats(ctx, &res.stats)) != GRLGPU_OK) fail("grlgpu_stats", rc);
    if (verbose) {
        std::cout << "Stats: " << std::endl;
        std::cout << "  Smallest symbol               : " << res.stats.min_sym << std::endl;
        std::cout << "  Greatest symbol               : " << res.stats.max_sym << std::endl;
        std::cout << "  Number of symbols in the file : " << res.stats.n_syms << std::endl;
        std::cout << "  Number of strings             : " << res.stats.n_strings << std::endl;
        std::cout << "Parsing the text:    " << std::endl;
    }
    for (;;) {
        grlgpu_round_t r;
        if ((rc = grlgpu_round(ctx, &r)) != GRLGPU_OK) fail("grlgpu_round", rc);
        if (r.sym_bytes == 8 && !res.wide) {  // first wide level: move what was collected so far to 64-bit symbols
            res.wide = true;
            for (const Level32& s : res.levels32) {
                Level L;
                L.alphabet = s.alphabet; L.tot_phrases = s.tot_phrases; L.has_hocc = s.has_hocc; L.pre_len = s.pre_len;
                if (!s.pre_len32.empty()) L.pre_len.assign(s.pre_len32.begin(), s.pre_len32.end());
                L.rule_l.assign(s.rule_l.begin(), s.rule_l.end()); L.rule_r.assign(s.rule_r.begin(), s.rule_r.end());
                L.pre_sym.assign(s.pre_sym.begin(), s.pre_sym.end());
                res.levels.push_back(std::move(L));
            }
            res.levels32.clear();
        }
        if (!res.wide) {
            Level32 L;
            L.alphabet = r.alphabet;
            L.tot_phrases = r.tot_phrases;
            L.has_hocc.resize(r.tot_phrases);
            L.rule_l.resize(r.tot_phrases);
            L.rule_r.resize(r.tot_phrases);
            L.pre_sym.resize(r.n_pre_runs);
            if (r.n_in + r.parse_len < (1ull << 32)) {  // run lengths fit 32 bits: 4 bytes less per run over PCIe and in memory
                L.pre_len32.resize(r.n_pre_runs);
                rc = grlgpu_fetch_level32(ctx, L.rule_l.data(), L.rule_r.data(), L.has_hocc.data(), L.pre_sym.data(), L.pre_len32.data(), 0);
            } else {
                L.pre_len.resize(r.n_pre_runs);
                rc = grlgpu_fetch_level(ctx, L.rule_l.data(), L.rule_r.data(), L.has_hocc.data(), L.pre_sym.data(), L.pre_len.data());
            }
            if (rc != GRLGPU_OK) fail("grlgpu_fetch_level", rc);
            res.levels32.push_back(std::move(L));
        } else {
            Level L;
            L.alphabet = r.alphabet;
            L.tot_phrases = r.tot_phrases;
            L.has_hocc.resize(r.tot_phrases);
            L.pre_len.resize(r.n_pre_runs);
            L.rule_l.resize(r.tot_phrases);
            L.rule_r.resize(r.tot_phrases);
            L.pre_sym.resize(r.n_pre_runs);
            if (r.sym_bytes == 8) {
                rc = grlgpu_fetch_level(ctx, L.rule_l.data(), L.rule_r.data(), L.has_hocc.data(), L.pre_sym.data(), L.pre_len.data());
                if (rc != GRLGPU_OK) fail("grlgpu_fetch_level", rc);
            } else {
                std::vector<uint32_t> l32(r.tot_phrases), r32(r.tot_phrases), p32(r.n_pre_runs);
                rc = grlgpu_fetch_level(ctx, l32.data(), r32.data(), L.has_hocc.data(), p32.data(), L.pre_len.data());
                if (rc != GRLGPU_OK) fail("grlgpu_fetch_level", rc);
                for (uint64_t i = 0; i < r.tot_phrases; i++) { L.rule_l[i] = l32[i]; L.rule_r[i] = r32[i]; }
                for (uint64_t i = 0; i < r.n_pre_runs; i++) L.pre_sym[i] = p32[i];
            }
            res.levels.push_back(std::move(L));
        }
        if (verbose) {
            std::cout << "  Parsing round " << r.round << std::endl;
            std::cout << "    Stats:" << std::endl;
            std::cout << "      Parsing phrases:                  " << r.n_phrases << std::endl;
            std::cout << "      Number of symbols in the phrases: " << r.dict_syms << std::endl;
            std::cout << "      Number of unsolved BWT blocks:    " << r.tot_phrases << std::endl;
            std::cout << "      Parse size:                       " << r.parse_len << std::endl;
            std::cout << "      Device time (ms):                 " << r.device_ms << " (text " << r.text_pass_ms << ", dictionary " << r.dict_ms
                      << ", rewrite " << r.rewrite_ms << ")" << std::endl;
        }
        res.rounds.push_back(r);
        if (r.done) {
            res.final_parse.resize(r.parse_len);
            std::vector<unsigned char> raw(r.parse_len * (uint64_t)r.cell_bytes_out);
            if ((rc = grlgpu_fetch_parse(ctx, raw.data())) != GRLGPU_OK) fail("grlgpu_fetch_parse", rc);
            for (uint64_t i = 0; i < r.parse_len; i++) {
                uint64_t v = 0;
                memcpy(&v, raw.data() + i * r.cell_bytes_out, r.cell_bytes_out);
                res.final_parse[i] = v;
            }
            break;
        }
    }
    res.par_ms = ms_since(t0);
    grlgpu_destroy(ctx);
    return res;
}

struct BwtResult {
    RunList runs;        // 64-bit symbols (wide alphabets)
    RunArr runs32;       // 32-bit symbols (the usual case), filled when narrow
    bool narrow = false;
    size_t n_runs() const { return narrow ? runs32.size() : runs.size(); }
    void write(const std::string& path) const {
        if (narrow) write_rl_bwt(path, runs32.sym.data(), runs32.len.data(), runs32.size(), sb, fb);
        else write_rl_bwt(path, runs.sym.data(), runs.len.data(), runs.size(), sb, fb);
    }
    uint64_t sb = 0, fb = 0;
    ParseResult parse;
    double ind_ms = 0;
};

inline BwtResult build_bwt(const void* text, uint64_t n_syms, int sym_bytes, int device, size_t n_threads, bool verbose) {
    BwtResult out;
    out.parse = gpu_par_phase(text, n_syms, sym_bytes, device, verbose);
    auto t0 = std::chrono::steady_clock::now();
    if (verbose) std::cout << "Inferring the BWT" << std::endl;
    if (out.parse.wide) out.runs = ind_phase<uint64_t>(out.parse.levels, out.parse.final_parse.data(), out.parse.final_parse.size());
    else {  // -t host threads drive the induction
        out.runs32 = ind_phase_mt(out.parse.levels32, out.parse.final_parse.data(), out.parse.final_parse.size(), std::max<size_t>(1, n_threads));
        out.narrow = true;
    }
    out.ind_ms = ms_since(t0);
    // header widths of the level-0 BWT (exact_ind_phase.cpp:274-276 with the level-0 dictionary: alphabet =
    // max_sym+1+3, prev_alphabet = 0, max_sym_freq from collection_stats)
    out.sb = int_ceil((uint64_t)sym_width(out.parse.stats.max_sym + 1 + 3), 8);
    out.fb = int_ceil((uint64_t)sym_width(out.parse.stats.max_sym_freq), 8);
    return out;
}

inline std::vector<unsigned char> read_whole_file(const std::string& path) {
    std::ifstream ifs(path, std::ios::binary | std::ios::ate);
    if (!ifs) throw std::runtime_error("cannot open " + path);
    const std::streamsize sz = ifs.tellg();
    ifs.seekg(0);
    std::vector<unsigned char> buf((size_t)sz);
    if (sz && !ifs.read((char*)buf.data(), sz)) throw std::runtime_error("cannot read " + path);
    return buf;
}

}  // namespace grlbwt

template <class sym_type, bool opt_bwt>
void grl_bwt_algo(std::string& i_file, std::string& o_file, tmp_workspace& tmp_ws, size_t n_threads, float hbuff_frac, uint8_t b_p_r, int device = 0) {
    (void)hbuff_frac;  // -f bounded the CPU hash buffers of the reference; the device table needs no such cap
    (void)b_p_r;       // -b only caps in-RAM run-length bytes in the reference; the output does not depend on it
    if constexpr (opt_bwt) {
        std::cout << "This option is broken" << std::endl;  // same as the reference (grl_bwt.hpp:29-31)
        exit(0);
    }
    std::cout << "Reading the file" << std::endl;
    std::vector<unsigned char> buf = grlbwt::read_whole_file(i_file);
    if (buf.empty() || buf.size() % sizeof(sym_type)) {
        std::cout << "Error: the file is ill formed" << std::endl;
        exit(1);
    }
    grlbwt::BwtResult res;
    try {
        res = grlbwt::build_bwt(buf.data(), buf.size() / sizeof(sym_type), (int)sizeof(sym_type), device, n_threads, true);
    } catch (const grlbwt::GpuError& e) {
        if (e.status == GRLGPU_ERR_ILL_FORMED) {
            std::cout << "Error: the file is ill formed" << std::endl;  // utils.cpp:177-180
            exit(1);
        }
        throw;
    }
    std::string tmp_out = tmp_ws.get_file("bwt_lev_0");
    res.write(tmp_out);
    std::error_code ec;
    std::filesystem::rename(tmp_out, o_file, ec);
    if (ec) {  // the reference fails across filesystems (grl_bwt.hpp:77); copy instead
        std::filesystem::copy_file(tmp_out, o_file, std::filesystem::copy_options::overwrite_existing);
        std::filesystem::remove(tmp_out);
    }
    std::cout << "The resulting BCR BWT was stored in " << o_file << std::endl;
}
