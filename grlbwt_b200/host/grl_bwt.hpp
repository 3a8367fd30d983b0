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
#include "gpu_par_phase.hpp"
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

struct BwtResult {
    RunList runs;        // 64-bit symbols (wide alphabets)
    RunArr runs32;       // 32-bit symbols (the usual case), filled when narrow
    bool narrow = false;
    size_t n_runs() const { return narrow ? runs32.size() : runs.size(); }
    void write(const std::string& path) const {
        if (narrow && !runs32.len32.empty()) write_rl_bwt(path, runs32.sym.data(), runs32.len32.data(), runs32.size(), sb, fb);
        else if (narrow) write_rl_bwt(path, runs32.sym.data(), runs32.len.data(), runs32.size(), sb, fb);
        else write_rl_bwt(path, runs.sym.data(), runs.len.data(), runs.size(), sb, fb);
    }
    uint64_t sb = 0, fb = 0;
    ParseResult parse;
    double ind_ms = 0;
};

// devices: one entry per rank; one entry = the single-GPU path. comm_kind: CommKind of gpu_par_phase.hpp
inline BwtResult build_bwt(const TextSource& src, int sym_bytes, const std::vector<int>& devices, int comm_kind, size_t n_threads, bool verbose,
                           const OutputBuffers* ob = nullptr) {
    BwtResult out;
    // the induction runs on the device when it can (fewer than 2^32 symbols, 32-bit level symbols, memory permitting): the levels
    // then never leave the GPU and the parse phase hands back the level-0 BWT; GRLBWT_HOST_INDUCTION=1 forces the host code
    const bool dev_ind = getenv("GRLBWT_HOST_INDUCTION") == nullptr && src.bytes / (uint64_t)sym_bytes < 0xfffffff0ull;
    out.parse = devices.size() > 1 ? gpu_par_phase_mg(src, sym_bytes, devices, comm_kind, verbose, dev_ind, ob)
                                   : gpu_par_phase(src, sym_bytes, devices.empty() ? 0 : devices[0], verbose, dev_ind, ob);
    auto t0 = std::chrono::steady_clock::now();
    if (out.parse.induced_on_device) {
        if (verbose) { printf("Inferring the BWT (on the device): %.1f ms\n", out.parse.dev_ind_ms); fflush(stdout); }
        out.runs32 = std::move(out.parse.bwt_dev);
        out.narrow = true;
        out.ind_ms = out.parse.dev_ind_ms;
    } else {
        if (verbose) { printf("Inferring the BWT\n"); fflush(stdout); }
        if (out.parse.wide) out.runs = ind_phase<uint64_t>(out.parse.levels, out.parse.final_parse.data(), out.parse.final_parse.size());
        else {  // -t host threads drive the induction
            out.runs32 = ind_phase_mt(out.parse.levels32, out.parse.final_parse.data(), out.parse.final_parse.size(), std::max<size_t>(1, n_threads));
            out.narrow = true;
        }
        out.ind_ms = ms_since(t0);
    }
    header_widths(out.parse.stats, out.sb, out.fb);
    return out;
}
inline BwtResult build_bwt(const void* text, uint64_t n_syms, int sym_bytes, int device, size_t n_threads, bool verbose) {
    TextSource src;
    src.mem = (const unsigned char*)text;
    src.bytes = n_syms * (uint64_t)sym_bytes;
    return build_bwt(src, sym_bytes, std::vector<int>{device}, COMM_AUTO, n_threads, verbose);
}

inline uint64_t file_size_of(const std::string& path) {
    std::error_code ec;
    const auto sz = std::filesystem::file_size(path, ec);
    if (ec) throw std::runtime_error("cannot open " + path);
    return (uint64_t)sz;
}

}  // namespace grlbwt

template <class sym_type, bool opt_bwt>
void grl_bwt_algo(std::string& i_file, std::string& o_file, tmp_workspace& tmp_ws, size_t n_threads, float hbuff_frac, uint8_t b_p_r,
                  const std::vector<int>& devices = std::vector<int>{0}, int comm_kind = 0) {
    (void)hbuff_frac;  // -f bounded the CPU hash buffers of the reference; the device table needs no such cap
    (void)b_p_r;       // -b only caps in-RAM run-length bytes in the reference; the output does not depend on it
    if constexpr (opt_bwt) {
        std::cout << "This option is broken" << std::endl;  // same as the reference (grl_bwt.hpp:29-31)
        exit(0);
    }
    std::cout << "Reading the file" << std::endl;
    grlbwt::TextSource src;
    src.file = i_file;
    src.bytes = grlbwt::file_size_of(i_file);  // streamed to the device(s) through pinned staging buffers, never held whole in host memory
    if (src.bytes == 0 || src.bytes % sizeof(sym_type)) {
        std::cout << "Error: the file is ill formed" << std::endl;
        exit(1);
    }
    grlbwt::BwtResult res;
    try {
        res = grlbwt::build_bwt(src, (int)sizeof(sym_type), devices, comm_kind, n_threads, true);
    } catch (const grlbwt::GpuError& e) {
        if (e.status == GRLGPU_ERR_ILL_FORMED) {
            std::cout << "Error: the file is ill formed" << std::endl;  // utils.cpp:177-180
            exit(1);
        }
        throw;
    }
    std::string tmp_out = tmp_ws.get_file("bwt_lev_0");
    const auto t_w = std::chrono::steady_clock::now();
    res.write(tmp_out);
    printf("Timing (ms): text to device %.1f, parse phase %.1f (%d rank%s, %s), induction %.1f (%s), writing %zu runs %.1f\n", res.parse.h2d_ms, res.parse.par_ms,
           res.parse.n_ranks, res.parse.n_ranks > 1 ? "s" : "", res.parse.comm_kind.c_str(), res.ind_ms, res.parse.induced_on_device ? "device" : "host", res.n_runs(),
           grlbwt::ms_since(t_w));
    fflush(stdout);
    std::error_code ec;
    std::filesystem::rename(tmp_out, o_file, ec);
    if (ec) {  // the reference fails across filesystems (grl_bwt.hpp:77); copy instead
        std::filesystem::copy_file(tmp_out, o_file, std::filesystem::copy_options::overwrite_existing);
        std::filesystem::remove(tmp_out);
    }
    std::cout << "The resulting BCR BWT was stored in " << o_file << std::endl;
}
