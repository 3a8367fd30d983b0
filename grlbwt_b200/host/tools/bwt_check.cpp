// bwt_check TEXT file.rl_bwt [-a cell_bytes] [-k strings_to_invert]: size-independent properties every BCR BWT of
// TEXT must satisfy (SURVEY.md 8d "parity at sizes the oracle cannot reach"): header widths per App. C, sum of run
// lengths = n, runs maximal, per-symbol totals = the text's histogram, the first N symbols = last real symbol of
// each string in string order, and a sample of strings recovered by LF-mapping equals the text.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <unordered_map>
#include "rl_bwt_tools.hpp"

static int fail(const std::string& what) {
    std::cout << "FAILED: " << what << std::endl;
    return 1;
}

int main(int argc, char** argv) {
    int w = 1;
    uint64_t k_inv = 64;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-a") && i + 1 < argc) w = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) k_inv = strtoull(argv[++i], nullptr, 10);
        else pos.push_back(argv[i]);
    }
    if (pos.size() != 2 || !(w == 1 || w == 2 || w == 4 || w == 8)) {
        std::cout << "usage: ./bwt_check TEXT file.rlbwt [-a cell_bytes] [-k strings_to_invert]" << std::endl;
        return 0;
    }
    std::ifstream ifs(pos[0], std::ios::binary | std::ios::ate);
    if (!ifs) return fail("cannot open " + pos[0]);
    const uint64_t bytes = (uint64_t)ifs.tellg(), n = bytes / (uint64_t)w;
    ifs.seekg(0);
    std::vector<unsigned char> raw(bytes);
    ifs.read((char*)raw.data(), (std::streamsize)bytes);
    auto cell = [&](uint64_t i) { uint64_t v = 0; memcpy(&v, raw.data() + i * w, (size_t)w); return v; };
    const uint64_t sep = cell(n - 1);

    grlbwt::RlBwt bwt(pos[1]);
    // header (App. C)
    std::unordered_map<uint64_t, uint64_t> hist;
    uint64_t max_sym = 0, max_freq = 0;
    for (uint64_t i = 0; i < n; i++) { const uint64_t c = cell(i); hist[c]++; if (c > max_sym) max_sym = c; }
    for (auto& kv : hist) max_freq = std::max(max_freq, kv.second);
    if (w > 1) max_freq = n;
    const uint64_t sb = grlbwt::int_ceil((uint64_t)grlbwt::sym_width(max_sym + 4), 8), fb = grlbwt::int_ceil((uint64_t)grlbwt::sym_width(max_freq), 8);
    if (bwt.sb != sb || bwt.fb != fb) return fail("header widths " + std::to_string(bwt.sb) + "/" + std::to_string(bwt.fb) + " expected " + std::to_string(sb) + "/" + std::to_string(fb));
    if (bwt.n != n) return fail("sum of run lengths " + std::to_string(bwt.n) + " != n " + std::to_string(n));
    std::unordered_map<uint64_t, uint64_t> bh;
    for (size_t i = 0; i < bwt.runs.size(); i++) {
        if (i && bwt.runs.sym[i] == bwt.runs.sym[i - 1]) return fail("runs " + std::to_string(i - 1) + " and " + std::to_string(i) + " are not maximal");
        if (bwt.runs.len[i] == 0) return fail("empty run");
        bh[bwt.runs.sym[i]] += bwt.runs.len[i];
    }
    if (bh.size() != hist.size()) return fail("alphabet differs");
    for (auto& kv : hist)
        if (bh[kv.first] != kv.second) return fail("count of symbol " + std::to_string(kv.first) + " differs");
    if (bwt.sep != sep) return fail("separator differs");
    // string starts
    std::vector<uint64_t> starts;
    starts.push_back(0);
    for (uint64_t i = 0; i + 1 < n; i++)
        if (cell(i) == sep) starts.push_back(i + 1);
    const uint64_t N = starts.size();
    if (bwt.n_strings != N) return fail("number of strings differs");
    // first N symbols: the symbol before each string's terminator, in string order
    {
        size_t run = 0;
        uint64_t used = 0;
        for (uint64_t s = 0; s < N; s++) {
            const uint64_t end = (s + 1 < N ? starts[s + 1] : n) - 1;  // position of the terminator
            const uint64_t expect = end > starts[s] ? cell(end - 1) : sep;
            if (bwt.runs.sym[run] != expect) return fail("BWT[" + std::to_string(s) + "] is not the last symbol of string " + std::to_string(s));
            if (++used == bwt.runs.len[run]) { run++; used = 0; }
        }
    }
    // LF inversion of a sample of strings
    const uint64_t K = std::min<uint64_t>(k_inv, N);
    for (uint64_t q = 0; q < K; q++) {
        const uint64_t s = K == N ? q : q * (N / K);
        const uint64_t end = (s + 1 < N ? starts[s + 1] : n) - 1;
        auto res = bwt.lf(s);
        uint64_t p = end;
        while (res.first != sep) {
            if (p == starts[s]) return fail("string " + std::to_string(s) + " inverts to something longer");
            p--;
            if (cell(p) != res.first) return fail("string " + std::to_string(s) + " differs at offset " + std::to_string(p - starts[s]));
            res = bwt.lf(res.second);
        }
        if (p != starts[s]) return fail("string " + std::to_string(s) + " inverts to something shorter");
    }
    std::cout << "OK n=" << n << " strings=" << N << " runs=" << bwt.runs.size() << " inverted=" << K << std::endl;
    return 0;
}
