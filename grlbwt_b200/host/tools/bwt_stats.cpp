// bwt_stats file.rl_bwt: size, runs, alphabet and run-length distribution of a .rl_bwt (reference scripts/bwt_stats.cpp)
#include <iostream>
#include "rl_bwt_tools.hpp"

static int run(int argc, char** argv) {
    if (argc != 2) {
        std::cout << "usage: ./bwt_stats file.rlbwt" << std::endl;
        return 0;
    }
    grlbwt::RlBwt bwt(argv[1]);
    const size_t r = bwt.runs.size();
    uint64_t longest = 0, lt256 = 0, lt65536 = 0;
    for (size_t i = 0; i < r; i++) {
        longest = std::max<uint64_t>(longest, bwt.runs.len[i]);
        lt256 += bwt.runs.len[i] < 256;
        lt65536 += bwt.runs.len[i] < 65536;
    }
    std::cout << "BWT size (n):            " << bwt.n << "\n"
              << "Number of runs (r):      " << r << "\n"
              << "n/r:                     " << (r ? double(bwt.n) / double(r) : 0.0) << "\n"
              << "Alphabet size:           " << bwt.C.size() << "\n"
              << "Separator symbol:        " << bwt.sep << "\n"
              << "Number of strings:       " << bwt.n_strings << "\n"
              << "Bytes per run symbol:    " << bwt.sb << "\n"
              << "Bytes per run length:    " << bwt.fb << "\n"
              << "Longest run:             " << longest << "\n"
              << "Runs shorter than 2^8:   " << (r ? 100.0 * double(lt256) / double(r) : 0.0) << " %\n"
              << "Runs shorter than 2^16:  " << (r ? 100.0 * double(lt65536) / double(r) : 0.0) << " %" << std::endl;
    return 0;
}
GRL_TOOL_MAIN(run)
