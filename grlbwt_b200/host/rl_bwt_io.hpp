// .rl_bwt run-length BWT format (bit-exact contract, reference include/bwt_io.h:184,377-382,448-490,541-550):
// bytes 0-7 sb (u64 LE), bytes 8-15 fb (u64 LE), then r records of sb symbol bytes + fb length bytes, LE.
// Fresh sequential writer/reader; the reference's in-place block cache is not needed here.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace grlbwt {

inline int sym_width(uint64_t v) { return v == 0 ? 0 : 64 - __builtin_clzll(v); }  // cdt_common.cpp:6-9
inline uint64_t int_ceil(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

struct RunList {
    std::vector<uint64_t> sym, len;
    size_t size() const { return sym.size(); }
    // appends, merging with the previous run when the symbol repeats (bwt_io.h:448-498 call sites)
    inline void push(uint64_t s, uint64_t l) {
        if (l == 0) return;
        if (!sym.empty() && sym.back() == s) { len.back() += l; return; }
        sym.push_back(s);
        len.push_back(l);
    }
};

static_assert(__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__, "the .rl_bwt records are packed with little-endian stores");

// the image of the file in memory: 16 + n_runs * (sb + fb) bytes at `dst` (same packing as write_rl_bwt below)
template <class SymT, class LenT>
inline void pack_rl_bwt(unsigned char* dst, const SymT* sym, const LenT* len, uint64_t n_runs, uint64_t sb, uint64_t fb) {
    const uint64_t hdr[2] = {sb, fb};
    memcpy(dst, hdr, 16);
    unsigned char* p = dst + 16;
    const uint64_t rec = sb + fb, tail = n_runs < 4 ? n_runs : 4;  // the two 8-byte stores of a record reach 8 - fb <= 7 bytes (< 4 records) past its end
    for (uint64_t i = 0; i + tail < n_runs; i++, p += rec) {
        const uint64_t s64 = (uint64_t)sym[i], l64 = (uint64_t)len[i];
        memcpy(p, &s64, 8);
        memcpy(p + sb, &l64, 8);
    }
    for (uint64_t i = n_runs - tail; i < n_runs; i++, p += rec) {  // the last records byte-exact: nothing is written past the image
        const uint64_t s64 = (uint64_t)sym[i], l64 = (uint64_t)len[i];
        memcpy(p, &s64, sb);
        memcpy(p + sb, &l64, fb);
    }
}

template <class SymT, class LenT>
inline void write_rl_bwt(const std::string& path, const SymT* sym, const LenT* len, uint64_t n_runs, uint64_t sb, uint64_t fb) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path + " for writing");
    const size_t rec = sb + fb;
    const size_t CAP = (size_t)1 << 24;           // records are packed with two 8-byte stores each, hence 16 bytes of slack
    std::vector<unsigned char> buf(CAP + 16);
    unsigned char* p = buf.data();
    uint64_t hdr[2] = {sb, fb};
    bool ok = fwrite(hdr, 8, 2, f) == 2;
    for (uint64_t i = 0; i < n_runs && ok; i++) {
        const uint64_t s64 = (uint64_t)sym[i], l64 = (uint64_t)len[i];
        memcpy(p, &s64, 8);                       // little endian hosts only: the low sb bytes are the record's symbol ...
        memcpy(p + sb, &l64, 8);                  // ... overwritten from offset sb on by the low fb bytes of the length
        p += rec;
        if ((size_t)(p - buf.data()) + rec > CAP) { ok = fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) == (size_t)(p - buf.data()); p = buf.data(); }
    }
    if (ok && p != buf.data()) ok = fwrite(buf.data(), 1, (size_t)(p - buf.data()), f) == (size_t)(p - buf.data());
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("short write on " + path);
}

inline void read_rl_bwt(const std::string& path, RunList& out, uint64_t& sb, uint64_t& fb) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    uint64_t hdr[2];
    if (fread(hdr, 8, 2, f) != 2) { fclose(f); throw std::runtime_error("truncated header in " + path); }
    sb = hdr[0]; fb = hdr[1];
    if (sb == 0 || sb > 8 || fb == 0 || fb > 8) { fclose(f); throw std::runtime_error("bad header in " + path); }
    unsigned char rec[16];
    out.sym.clear(); out.len.clear();
    size_t got;
    while ((got = fread(rec, 1, sb + fb, f)) == sb + fb) {
        uint64_t s = 0, l = 0;
        memcpy(&s, rec, sb);
        memcpy(&l, rec + sb, fb);
        out.sym.push_back(s);
        out.len.push_back(l);
    }
    fclose(f);
    if (got != 0) throw std::runtime_error("truncated record at the end of " + path);  // the file must be 16 + r * (sb + fb) bytes
}

}  // namespace grlbwt
