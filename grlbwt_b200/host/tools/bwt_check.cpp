// bwt_check TEXT file.rl_bwt [-a cell_bytes] [-k strings_to_invert]: size-independent properties every BCR BWT of
// TEXT must satisfy (SURVEY.md 8d "parity at sizes the oracle cannot reach"): header widths per App. C, sum of run
// lengths = n, runs maximal, per-symbol totals = the text's histogram, the first N symbols = last real symbol of
// each string in string order, and a sample of strings recovered by LF-mapping equals the text.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <unordered_map>
#include "rl_bwt_tools.hpp"

static int fail(const std::string& what) {
    std::cout << "FAILED: " << what << std::endl;
    return 1;
}

static int run(int argc, char** argv) {
    int w = 1;
    uint64_t k_inv = 64, max_steps = 50000000;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-a") && i + 1 < argc) w = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) k_inv = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "-m") && i + 1 < argc) max_steps = strtoull(argv[++i], nullptr, 10);
        else pos.push_back(argv[i]);
    }
    if (pos.size() != 2 || !(w == 1 || w == 2 || w == 4 || w == 8)) {
        std::cout << "usage: ./bwt_check TEXT file.rlbwt [-a cell_bytes] [-k strings_to_invert] [-m max_LF_steps]" << std::endl;
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
    if (w <= 2) {  // dense per-thread counters: the text has billions of cells at full size
        const size_t S = (size_t)1 << (8 * w), T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::vector<uint64_t>> part(T, std::vector<uint64_t>(S, 0));
        std::vector<std::thread> th;
        for (size_t t = 0; t < T; t++)
            th.emplace_back([&, t] {
                std::vector<uint64_t>& h = part[t];
                const uint64_t b = n / T * t, e = t + 1 == T ? n : n / T * (t + 1);
                if (w == 1) for (uint64_t i = b; i < e; i++) h[raw[i]]++;
                else { const uint16_t* p = (const uint16_t*)raw.data(); for (uint64_t i = b; i < e; i++) h[p[i]]++; }
            });
        for (auto& x : th) x.join();
        for (size_t c = 0; c < S; c++) {
            uint64_t tot = 0;
            for (size_t t = 0; t < T; t++) tot += part[t][c];
            if (tot) { hist[c] = tot; max_sym = c; }
        }
    } else
        for (uint64_t i = 0; i < n; i++) { const uint64_t c = cell(i); hist[c]++; if (c > max_sym) max_sym = c; }
    for (auto& kv : hist) max_freq = std::max(max_freq, kv.second);
    if (w > 1) max_freq = n;
    const uint64_t sb = grlbwt::int_ceil((uint64_t)grlbwt::sym_width(max_sym + 4), 8), fb = grlbwt::int_ceil((uint64_t)grlbwt::sym_width(max_freq), 8);
    if (bwt.sb != sb || bwt.fb != fb) return fail("header widths " + std::to_string(bwt.sb) + "/" + std::to_string(bwt.fb) + " expected " + std::to_string(sb) + "/" + std::to_string(fb));
    if (bwt.n != n) return fail("sum of run lengths " + std::to_string(bwt.n) + " != n " + std::to_string(n));
    for (size_t i = 0; i < bwt.runs.size(); i++) {
        if (i && bwt.runs.sym[i] == bwt.runs.sym[i - 1]) return fail("runs " + std::to_string(i - 1) + " and " + std::to_string(i) + " are not maximal");
        if (bwt.runs.len[i] == 0) return fail("empty run");
    }
    // per-symbol totals of the BWT (RlBwt::C holds the exclusive prefix sums of them, in symbol order)
    if (bwt.C.size() != hist.size()) return fail("alphabet differs");
    for (auto it = bwt.C.begin(); it != bwt.C.end(); ++it) {
        auto nx = std::next(it);
        const uint64_t tot = (nx == bwt.C.end() ? bwt.n : nx->second) - it->second;
        auto h = hist.find(it->first);
        if (h == hist.end() || h->second != tot) return fail("count of symbol " + std::to_string(it->first) + " differs");
    }
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
    // (every string gets the same share of the step budget -m: collections of few, very long strings are checked on the last
    // budget/K symbols of each sampled string instead of walking gigabases)
    const uint64_t K = std::min<uint64_t>(k_inv, N);
    const uint64_t per_string = K ? std::max<uint64_t>(1, max_steps / K) : 0;
    uint64_t steps = 0, whole = 0;
    for (uint64_t q = 0; q < K; q++) {
        const uint64_t s = K == N ? q : q * (N / K);
        const uint64_t end = (s + 1 < N ? starts[s + 1] : n) - 1;
        auto res = bwt.lf(s);
        uint64_t p = end, used = 0;
        bool cut = false;
        while (res.first != sep) {
            if (p == starts[s]) return fail("string " + std::to_string(s) + " inverts to something longer");
            p--;
            if (cell(p) != res.first) return fail("string " + std::to_string(s) + " differs at offset " + std::to_string(p - starts[s]));
            if (++used >= per_string) { cut = p != starts[s]; break; }
            res = bwt.lf(res.second);
        }
        steps += used;
        if (!cut) {
            if (p != starts[s]) return fail("string " + std::to_string(s) + " inverts to something shorter");
            whole++;
        }
    }
    std::cout << "OK n=" << n << " strings=" << N << " runs=" << bwt.runs.size() << " inverted=" << K << " (whole strings " << whole << ", LF steps " << steps << ")"
              << std::endl;
    return 0;
}
GRL_TOOL_MAIN(run)
