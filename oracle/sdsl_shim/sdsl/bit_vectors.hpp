// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for the subset of sdsl-lite that
// grlBWT's exact path touches, so the UNMODIFIED reference sources under /root/reference
// compile in an image without sdsl-lite. Nothing in the product path
// includes this header. On-disk layouts written through it are private to one reference run.
#pragma once
#include <vector>
#include <string>
#include <cstdint>
#include <iostream>
#include <fstream>
#include <typeinfo>
#include <functional>
#include <algorithm>
#include <limits>
#include <cstring>
#include <cmath>
#include <iomanip>
#include <map>
#include <thread>
#include <chrono>

namespace sdsl {

struct structure_tree_node {};
struct structure_tree {
    static structure_tree_node* add_child(structure_tree_node*, const std::string&, const std::string&) { return nullptr; }
};

namespace bits {
inline uint32_t hi(uint64_t x) { return x == 0 ? 0 : 63 - __builtin_clzll(x); }
}

template <class T>
size_t write_member(const T& t, std::ostream& out, structure_tree_node* = nullptr, std::string = "") {
    out.write((const char*)&t, sizeof(T));
    return sizeof(T);
}
template <class T>
void read_member(T& t, std::istream& in) { in.read((char*)&t, sizeof(T)); }

class rank_support_v1;

class bit_vector {
public:
    typedef size_t size_type;
    typedef rank_support_v1 rank_1_type;
    std::vector<uint64_t> w;
    size_t n = 0;

    struct ref {
        uint64_t* p;
        uint64_t m;
        operator bool() const { return (*p & m) != 0; }
        ref& operator=(bool b) { if (b) *p |= m; else *p &= ~m; return *this; }
        ref& operator=(const ref& o) { return *this = bool(o); }
    };

    bit_vector() = default;
    bit_vector(size_t n_, bool v = false) : w((n_ + 63) / 64, v ? ~0ULL : 0ULL), n(n_) {}
    size_t size() const { return n; }
    bool operator[](size_t i) const { return (w[i >> 6] >> (i & 63)) & 1ULL; }
    ref operator[](size_t i) { return ref{&w[i >> 6], 1ULL << (i & 63)}; }
    void resize(size_t n_) {
        w.resize((n_ + 63) / 64, 0);
        n = n_;
        if (n & 63) w.back() &= ((1ULL << (n & 63)) - 1);
    }
    void swap(bit_vector& o) { w.swap(o.w); std::swap(n, o.n); }
    size_t serialize(std::ostream& out, structure_tree_node* = nullptr, std::string = "") const {
        out.write((const char*)&n, sizeof(n));
        out.write((const char*)w.data(), w.size() * 8);
        return 8 + w.size() * 8;
    }
    void load(std::istream& in) {
        in.read((char*)&n, sizeof(n));
        w.assign((n + 63) / 64, 0);
        in.read((char*)w.data(), w.size() * 8);
    }
};

class rank_support_v1 {
    const bit_vector* bv = nullptr;
    std::vector<size_t> blk;
public:
    rank_support_v1() = default;
    explicit rank_support_v1(const bit_vector* b) : bv(b) {
        blk.resize(b->w.size() + 1);
        size_t a = 0;
        for (size_t i = 0; i < b->w.size(); i++) { blk[i] = a; a += __builtin_popcountll(b->w[i]); }
        blk[b->w.size()] = a;
    }
    size_t operator()(size_t i) const {
        size_t r = blk[i >> 6];
        if (i & 63) r += __builtin_popcountll(bv->w[i >> 6] & ((1ULL << (i & 63)) - 1));
        return r;
    }
    size_t rank(size_t i) const { return (*this)(i); }
    void swap(rank_support_v1& o) { std::swap(bv, o.bv); blk.swap(o.blk); }
};

namespace util {
template <class T> void clear(T& t) { T tmp; t.swap(tmp); }
inline void set_to_value(bit_vector& b, bool v) {
    for (auto& x : b.w) x = v ? ~0ULL : 0ULL;
    if (b.n & 63) b.w.back() &= ((1ULL << (b.n & 63)) - 1);
}
template <class T> std::string class_name(const T&) { return typeid(T).name(); }
}  // namespace util

template <class T>
bool store_to_file(const T& t, const std::string& f) {
    std::ofstream o(f, std::ios::binary);
    if (!o) return false;
    t.serialize(o);
    return true;
}
template <class T>
bool load_from_file(T& t, const std::string& f) {
    std::ifstream i(f, std::ios::binary);
    if (!i) return false;
    t.load(i);
    return true;
}

}  // namespace sdsl
