// Shared pieces of the .rl_bwt consumer tools (fresh equivalents of the reference's scripts/*.cpp, which need
// SDSL wavelet trees): a run-length BWT held as arrays, and LF-mapping over it with per-run rank prefixes and a
// binary search on run starts (O(log r) per step, any symbol width).
#pragma once
#include <algorithm>
#include <cstdint>
#include <exception>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "../rl_bwt_io.hpp"

namespace grlbwt {

struct RlBwt {
    RunList runs;
    uint64_t sb = 0, fb = 0, n = 0;
    std::vector<uint64_t> start;      // start[i] = position of run i; start[r] = n
    std::vector<uint64_t> before;     // before[i] = occurrences of runs.sym[i] in runs 0..i-1
    std::map<uint64_t, uint64_t> C;   // C[c] = number of symbols smaller than c
    uint64_t sep = 0, n_strings = 0;

    explicit RlBwt(const std::string& path) {
        read_rl_bwt(path, runs, sb, fb);
        const size_t r = runs.size();
        start.resize(r + 1);
        before.resize(r);
        std::map<uint64_t, uint64_t> cnt;
        uint64_t max_sym = 0;
        for (size_t i = 0; i < r; i++) max_sym = std::max(max_sym, runs.sym[i]);
        if (max_sym < (1ull << 26)) {  // dense counters: one array access per run instead of a tree walk (billions of runs at full size)
            std::vector<uint64_t> dense(max_sym + 1, 0);
            for (size_t i = 0; i < r; i++) {
                start[i] = n;
                before[i] = dense[runs.sym[i]];
                dense[runs.sym[i]] += runs.len[i];
                n += runs.len[i];
            }
            for (uint64_t c = 0; c <= max_sym; c++)
                if (dense[c]) cnt[c] = dense[c];
        } else {
            for (size_t i = 0; i < r; i++) {
                start[i] = n;
                before[i] = cnt[runs.sym[i]];
                cnt[runs.sym[i]] += runs.len[i];
                n += runs.len[i];
            }
        }
        start[r] = n;
        uint64_t acc = 0;
        for (auto& kv : cnt) { C[kv.first] = acc; acc += kv.second; }
        if (!cnt.empty()) { sep = cnt.begin()->first; n_strings = cnt.begin()->second; }  // the separator is the smallest symbol
    }
    // symbol at position i and the row LF(i) = C[c] + rank_c(i)
    std::pair<uint64_t, uint64_t> lf(uint64_t i) const {
        const size_t k = (size_t)(std::upper_bound(start.begin(), start.end(), i) - start.begin()) - 1;
        const uint64_t c = runs.sym[k];
        return {c, C.at(c) + before[k] + (i - start[k])};
    }
};

inline void put_cell(std::vector<unsigned char>& out, uint64_t v, int width) {
    for (int b = 0; b < width; b++) out.push_back((unsigned char)(v >> (8 * b)));
}

}  // namespace grlbwt

// a malformed input (bad header, truncated record) ends the tool with a message and status 1, not an abort
#define GRL_TOOL_MAIN(fn)                                             \
    int main(int argc, char** argv) {                                 \
        try { return fn(argc, argv); }                                \
        catch (const std::exception& e) { std::cerr << "Error: " << e.what() << std::endl; return 1; } \
    }
