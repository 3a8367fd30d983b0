// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference parse phase
// (/root/reference/lib/exact_algo/exact_par_phase.cpp, pulled in by #include so the template
// exact_algo::par_phase_int is reachable) one round at a time, so that every round's parse,
// tot_phrases and string count can be dumped as golden fixtures, and so the
// reference's parse phase alone can be timed on the host cores (bench.py --impl reference).
// The round sequencing below restates the loop of exact_algo::par_phase
// (exact_par_phase.cpp:285-372): collection_stats -> round 1 with the first-round parser ->
// later rounds with the parser chosen by sym_width(n_syms)+1.
//
// usage: ref_harness dump  <input> <alph_bytes> <threads> <out_dir>
//        ref_harness parse <input> <alph_bytes> <threads> <tmp_dir>      (prints "PAR_PHASE_SECONDS x")
#include "exact_par_phase.cpp"
#include <filesystem>
#include <cstdio>

namespace fs = std::filesystem;

template <class sym_t>
static int run(const std::string& mode, std::string input, size_t threads, const std::string& out_dir) {
    tmp_workspace ws(mode == "dump" ? out_dir : out_dir, true, "grl.ref");
    auto t0 = std::chrono::steady_clock::now();
    str_collection sc = collection_stats<sym_t>(input);
    size_t hbuff = std::max<size_t>(64 * threads, size_t(ceil(float(sc.n_syms) * 0.15f)));

    std::string out_f = ws.get_file("tmp_output");
    std::string in_f = ws.get_file("tmp_input");

    bv_t sym_desc(sc.max_sym + 1, false);
    sym_desc[sc.min_sym] = true;

    parsing_info pi;
    pi.max_sym_freq = sc.max_sym_freq;
    pi.tot_phrases = sc.max_sym + 1;
    pi.str_ptrs.swap(sc.str_ptrs);
    pi.str_ptrs.push_back((long)sc.n_syms);
    pi.longest_str = sc.longest_string;
    pi.active_strings = sc.n_strings;

    FILE* meta = nullptr;
    if (mode == "dump") {
        meta = fopen((out_dir + "/rounds.txt").c_str(), "w");
        fprintf(meta, "# n_syms %zu n_strings %zu min %zu max %zu max_sym_freq %zu longest %zu\n", sc.n_syms,
                sc.n_strings, sc.min_sym, sc.max_sym, sc.max_sym_freq, sc.longest_string);
    }

    auto dump_round = [&](size_t round) {
        if (!meta) return;
        size_t bps = sym_width(pi.tot_phrases) + 1;
        size_t cell = bps <= 8 ? 1 : bps <= 16 ? 2 : bps <= 32 ? 4 : 8;
        std::string dst = out_dir + "/parse_r" + std::to_string(round) + ".bin";
        fs::copy_file(in_f, dst, fs::copy_options::overwrite_existing);
        fprintf(meta, "round %zu lms_phrases %zu tot_phrases %zu cell_bytes %zu parse_cells %zu n_strings %zu longest %zu\n",
                round, pi.lms_phrases, pi.tot_phrases, cell, (size_t)fs::file_size(dst) / cell,
                pi.str_ptrs.size() - 1, pi.longest_str);
        std::string sp = out_dir + "/str_ptrs_r" + std::to_string(round) + ".bin";
        FILE* f = fopen(sp.c_str(), "wb");
        fwrite(pi.str_ptrs.data(), sizeof(long), pi.str_ptrs.size(), f);
        fclose(f);
    };

    size_t round = 1;
    using first_parser = lms_parsing<i_file_stream<sym_t>, string_t, true>;
    size_t n_syms = exact_algo::par_phase_int<first_parser>(input, in_f, pi, hbuff, threads, sym_desc, ws);
    dump_round(round);
    while (n_syms > 0) {
        round++;
        size_t bps = sym_width(n_syms) + 1;
        if (bps <= 8) n_syms = exact_algo::par_phase_int<uint8t_parser_t>(in_f, out_f, pi, hbuff, threads, sym_desc, ws);
        else if (bps <= 16) n_syms = exact_algo::par_phase_int<uint16t_parser_t>(in_f, out_f, pi, hbuff, threads, sym_desc, ws);
        else if (bps <= 32) n_syms = exact_algo::par_phase_int<uint32t_parser_t>(in_f, out_f, pi, hbuff, threads, sym_desc, ws);
        else n_syms = exact_algo::par_phase_int<uint64t_parser_t>(in_f, out_f, pi, hbuff, threads, sym_desc, ws);
        remove(in_f.c_str());
        rename(out_f.c_str(), in_f.c_str());
        dump_round(round);
    }
    auto t1 = std::chrono::steady_clock::now();
    if (meta) fclose(meta);
    printf("PAR_PHASE_ROUNDS %zu\n", round);
    printf("PAR_PHASE_SECONDS %.6f\n", std::chrono::duration<double>(t1 - t0).count());
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 6) {
        fprintf(stderr, "usage: %s dump|parse <input> <alph_bytes> <threads> <dir>\n", argv[0]);
        return 2;
    }
    std::string mode = argv[1], input = argv[2], dir = argv[5];
    int a = atoi(argv[3]);
    size_t t = (size_t)atoi(argv[4]);
    if (a == 1) return run<uint8_t>(mode, input, t, dir);
    if (a == 2) return run<uint16_t>(mode, input, t, dir);
    if (a == 4) return run<uint32_t>(mode, input, t, dir);
    if (a == 8) return run<uint64_t>(mode, input, t, dir);
    return 2;
}
