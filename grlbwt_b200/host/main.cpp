// grlbwt / grlbwt-cli: same command line as the reference tool (main.cpp:43-154 there):
//   grlbwt TEXT [-o out] [-a 1|2|4|8] [-t N] [-f frac] [-b 0..5] [-T dir] [-v]   (+ -g/--gpu DEVICE)
// Output naming rule (main.cpp:112-113): -o, else filename(TEXT); the extension is replaced by
// ".rl_bwt"; relative to the current directory.
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>

#include "grl_bwt.hpp"

struct arguments {
    std::string input_file, output_file, tmp_dir = "/tmp";
    size_t n_threads = 1;
    uint8_t b_f_r = 1;
    float hbuff_frac = 0.15f;
    bool ver = false;
    uint8_t alph_bytes = 1;
    std::vector<int> devices{0};   // one rank per entry
    int comm_kind = 0;
    std::string version = "v1.0.1 alpha (B200 parse phase)";
};

static void usage(const char* prog) {
    std::cout << "Repetition-aware BWT construction\n"
              << "Usage: " << prog << " [OPTIONS] TEXT\n\n"
              << "Positionals:\n"
              << "  TEXT                   Input file in one-string-per-line format\n\n"
              << "Options:\n"
              << "  -h,--help              Print this help message and exit\n"
              << "  -o,--output-file       Output file\n"
              << "  -a,--alphabet          Number of bytes for the alphabet (def. 1)\n"
              << "  -t,--threads           Maximum number of working threads\n"
              << "  -f,--hbuff             Hashing step will use at most INPUT_SIZE*f bytes. O means no limit (def. 0.5)\n"
              << "  -b,--run-len-bytes     Max. number of bytes to encode the run lengths in the recursive BWTs (def. 1)\n"
              << "  -T,--tmp               Temporary folder (def. /tmp/grl.bwt.xxxx)\n"
              << "  -g,--gpu               CUDA device(s) that run the parse phase: one rank per entry of a comma list (def. 0)\n"
              << "  --gpus                 Number of GPUs: ranks on devices 0..N-1 (shards of whole strings, partitioned dictionary)\n"
              << "  --comm                 Exchange between the ranks: nccl, local (in-process peer copies over NVLink), auto (def.: local when\n"
              << "                         every pair of GPUs has peer access, else nccl)\n"
              << "  -v,--version           Print the software version and exit\n";
}

[[noreturn]] static void bad(const std::string& msg, int code) {
    std::cerr << msg << "\nRun with --help for more information." << std::endl;
    exit(code);
}

static bool parse_args(int argc, char** argv, arguments& a) {
    bool have_text = false;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        auto value = [&](const std::string& name) -> std::string {
            if (i + 1 >= argc) bad(name + ": 1 required", 114);
            return argv[++i];
        };
        try {
            if (s == "-h" || s == "--help") { usage(argv[0]); exit(0); }
            else if (s == "-v" || s == "--version") a.ver = true;
            else if (s == "-o" || s == "--output-file") a.output_file = value(s);
            else if (s == "-a" || s == "--alphabet") {
                std::string v = value(s);
                if (!(v == "1" || v == "2" || v == "4" || v == "8")) bad("--alphabet: " + v + " is not a valid number of bytes for a native integer type", 105);
                a.alph_bytes = (uint8_t)std::stoi(v);
            } else if (s == "-t" || s == "--threads") a.n_threads = (size_t)std::stoul(value(s));
            else if (s == "-f" || s == "--hbuff") {
                a.hbuff_frac = std::stof(value(s));
                if (a.hbuff_frac < 0.f || a.hbuff_frac > 1.f) bad("--hbuff: Value " + std::to_string(a.hbuff_frac) + " not in range 0 to 1", 105);
            } else if (s == "-b" || s == "--run-len-bytes") {
                int b = std::stoi(value(s));
                if (b < 0 || b > 5) bad("--run-len-bytes: Value " + std::to_string(b) + " not in range 0 to 5", 105);
                a.b_f_r = (uint8_t)b;
            } else if (s == "-T" || s == "--tmp") {
                a.tmp_dir = value(s);
                if (!std::filesystem::is_directory(a.tmp_dir)) bad("--tmp: Directory does not exist: " + a.tmp_dir, 105);
            } else if (s == "-g" || s == "--gpu") {
                const std::string v = value(s);
                a.devices.clear();
                size_t p0 = 0;
                while (p0 <= v.size()) {
                    const size_t p1 = v.find(',', p0);
                    a.devices.push_back(std::stoi(v.substr(p0, p1 == std::string::npos ? std::string::npos : p1 - p0)));
                    if (p1 == std::string::npos) break;
                    p0 = p1 + 1;
                }
                if (a.devices.empty() || a.devices.size() > 31) bad("--gpu: between 1 and 31 devices", 105);
            } else if (s == "--gpus") {
                const int n = std::stoi(value(s));
                if (n < 1 || n > 31) bad("--gpus: Value " + std::to_string(n) + " not in range 1 to 31", 105);
                a.devices.clear();
                for (int d = 0; d < n; d++) a.devices.push_back(d);
            } else if (s == "--comm") {
                const std::string v = value(s);
                if (v == "auto") a.comm_kind = 0; else if (v == "local") a.comm_kind = 1; else if (v == "nccl") a.comm_kind = 2;
                else bad("--comm: " + v + " is not one of auto, nccl, local", 105);
            }
            else if (!s.empty() && s[0] == '-' && s.size() > 1) bad("The following argument was not expected: " + s, 109);
            else {
                if (have_text) bad("The following argument was not expected: " + s, 109);
                a.input_file = s;
                have_text = true;
            }
        } catch (const std::invalid_argument&) {
            bad("Could not convert: " + s, 104);
        } catch (const std::out_of_range&) {
            bad("Could not convert: " + s, 104);
        }
    }
    if (a.ver) return true;
    if (!have_text) bad("TEXT is required", 106);
    if (!std::filesystem::is_regular_file(a.input_file)) bad("TEXT: File does not exist: " + a.input_file, 105);
    return true;
}

template <class sym_type>
static void run_int(std::string input_collection, arguments& args) {
    tmp_workspace tmp_ws(args.tmp_dir, true, "grl.bwt");
    std::cout << "Temporary folder: " << tmp_ws.folder() << std::endl;
    std::cout << "BWT type:         BCR exact" << std::endl;
    grl_bwt_algo<sym_type, false>(input_collection, args.output_file, tmp_ws, args.n_threads, args.hbuff_frac, args.b_f_r, args.devices, args.comm_kind);
}

int main(int argc, char** argv) {
    arguments args;
    parse_args(argc, argv, args);
    if (args.ver) {
        std::cout << args.version << std::endl;
        return 0;
    }
    std::cout << "Input file:       " << args.input_file << std::endl;
    if (args.output_file.empty()) args.output_file = std::filesystem::path(args.input_file).filename().string();
    args.output_file = std::filesystem::path(args.output_file).replace_extension(".rl_bwt").string();
    std::cout << (args.alph_bytes > 1 ? "Alphabet type:    integer" : "Alphabet type:    byte") << std::endl;
    try {
        if (args.alph_bytes == 1) run_int<uint8_t>(args.input_file, args);
        else if (args.alph_bytes == 2) run_int<uint16_t>(args.input_file, args);
        else if (args.alph_bytes == 4) run_int<uint32_t>(args.input_file, args);
        else run_int<uint64_t>(args.input_file, args);
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
