// bwt_stats file.rl_bwt: the report of the reference's scripts/bwt_stats.cpp:9-108 (runs and text frequency per symbol, share of
// the run lengths that fit 1 / 2 / 3+ bytes, deciles of the run lengths, longest / shortest run), for any symbol width (the
// reference indexes 256-entry arrays with the symbol), followed by a summary of the file itself (header widths, strings).
#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <tuple>
#include "rl_bwt_tools.hpp"

static int run(int argc, char** argv) {
    if (argc != 2) {
        std::cout << "usage: ./bwt_stats file.rlbwt\n"
                     "file.rlbwt is the bcr bwt file\n" << std::endl;
        return 0;
    }
    std::cout << "Reading the input BWT" << std::endl;
    grlbwt::RlBwt bwt(argv[1]);
    const size_t r = bwt.runs.size();
    uint64_t longest = 0, shortest = ~0ull, one_byte = 0, two_bytes = 0, three_bytes = 0, lt256 = 0, lt65536 = 0;
    std::map<uint64_t, std::pair<uint64_t, uint64_t>> per_sym;  // symbol -> (runs, text frequency)
    std::vector<uint64_t> lens(bwt.runs.len);
    for (size_t i = 0; i < r; i++) {
        const uint64_t len = bwt.runs.len[i];
        longest = std::max(longest, len);
        shortest = std::min(shortest, len);
        if (len <= 255) one_byte++; else if (len <= 65535) two_bytes++; else three_bytes++;
        lt256 += len < 256;
        lt65536 += len < 65536;
        auto& e = per_sym[bwt.runs.sym[i]];
        e.first++;
        e.second += len;
    }
    std::sort(lens.begin(), lens.end());
    std::vector<std::tuple<uint64_t, uint64_t, uint64_t>> rl_stats;
    for (const auto& kv : per_sym) rl_stats.emplace_back(kv.first, kv.second.first, kv.second.second);
    std::stable_sort(rl_stats.begin(), rl_stats.end(), [](const auto& a, const auto& b) { return std::get<1>(a) < std::get<1>(b); });  // fewest runs first

    std::cout << "Number of runs: " << r << std::endl;
    std::cout << "Text alphabet: " << per_sym.size() << std::endl;
    std::cout << "Run stats:" << std::endl;
    uint64_t k = 1, space_acc = 0;
    for (const auto& t : rl_stats) {
        const uint64_t space = 32 * std::get<1>(t) * 2;  // two 32-bit words per run (:81)
        std::cout << k++ << ") Symbol:" << std::get<0>(t) << "\t\tnumber of runs in the BWT:" << std::get<1>(t) << "\t\ttext frequency:" << std::get<2>(t) << " | " << space << " "
                  << space_acc << std::endl;
        space_acc += space;
    }
    const double l = (double)r;
    std::cout << "Text size: " << bwt.n << std::endl;
    std::cout << "n/r: " << (r ? double(bwt.n) / l : 0.0) << std::endl;
    std::cout << (r ? double(one_byte) / l * 100 : 0.0) << "% of the run lenghts fit 1 byte" << std::endl;
    std::cout << (r ? double(two_bytes) / l * 100 : 0.0) << "% of the run lenghts fit 2 bytes" << std::endl;
    std::cout << (r ? double(three_bytes) / l * 100 : 0.0) << "% of the run lenghts fit 3 or more bytes" << std::endl;
    std::cout << "Deciles: " << std::endl;
    double prop = 0.1;
    for (int i = 0; i < 9 && r; i++) {
        const size_t q = std::min<size_t>((size_t)std::ceil(l * prop), r - 1);  // (the reference reads one past the end when r < 10)
        std::cout << "  q" << (i + 1) << ": " << lens[q] << std::endl;
        prop += 0.1;
    }
    std::cout << "Longest run: " << longest << std::endl;
    std::cout << "Shortest run: " << (r ? shortest : 0) << std::endl;

    std::cout << "\nBWT size (n):            " << bwt.n << "\n"
              << "Number of runs (r):      " << r << "\n"
              << "Alphabet size:           " << bwt.C.size() << "\n"
              << "Separator symbol:        " << bwt.sep << "\n"
              << "Number of strings:       " << bwt.n_strings << "\n"
              << "Bytes per run symbol:    " << bwt.sb << "\n"
              << "Bytes per run length:    " << bwt.fb << "\n"
              << "Runs shorter than 2^8:   " << (r ? 100.0 * double(lt256) / double(r) : 0.0) << " %\n"
              << "Runs shorter than 2^16:  " << (r ? 100.0 * double(lt65536) / double(r) : 0.0) << " %" << std::endl;
    return 0;
}
GRL_TOOL_MAIN(run)
