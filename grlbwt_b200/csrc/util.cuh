// Shared device/host helpers for the grlgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef unsigned long long u64;  // matches CUDA's atomic overloads

namespace grl {

// Error transport inside the library: thrown on the host side, converted to a status code at the C ABI.
struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

#define GRL_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw grl::Error(-3 /*GRLGPU_ERR_CUDA*/, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                                         " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define GRL_KERNEL_CHECK() GRL_CUDA(cudaGetLastError())

static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }
static inline int bit_width64(u64 v) { return v == 0 ? 0 : 64 - __builtin_clzll(v); }  // = reference sym_width()

// Device memory of one context: a few large cudaMalloc'ed slabs carved by a first-fit free list with
// coalescing. Every buffer of every round comes from here, so after the first step no allocation
// reaches the driver (HBM is laid out once and reused round after round). All work of a context is
// issued on one stream, so a block freed on the host may be handed out again immediately: the
// kernels that used it were enqueued before the kernels that will.
struct DevicePool {
    struct Slab { char* base; u64 size; std::map<u64, u64> free_ranges; };  // offset -> length
    std::vector<Slab> slabs;
    std::map<void*, std::pair<int, u64>> live;  // ptr -> (slab, size)
    u64 reserved = 0, in_use = 0, peak = 0;
    static constexpr u64 ALIGN = 512, MIN_SLAB = 1ull << 28;
    void* alloc(u64 bytes) {
        bytes = (std::max<u64>(bytes, 1) + ALIGN - 1) / ALIGN * ALIGN;
        for (int pass = 0; pass < 2; pass++) {
            int best_s = -1; u64 best_off = 0, best_len = ~0ULL;
            for (int si = 0; si < (int)slabs.size(); si++)
                for (auto& fr : slabs[si].free_ranges)
                    if (fr.second >= bytes && fr.second < best_len) { best_s = si; best_off = fr.first; best_len = fr.second; }
            if (best_s >= 0) {
                Slab& sl = slabs[best_s];
                sl.free_ranges.erase(best_off);
                if (best_len > bytes) sl.free_ranges[best_off + bytes] = best_len - bytes;
                void* p = sl.base + best_off;
                live[p] = {best_s, bytes};
                in_use += bytes; peak = std::max(peak, in_use);
                return p;
            }
            // grow: a new slab at least as large as the request (and as what is already reserved, up to 16 GB)
            u64 sz = std::max<u64>({bytes, MIN_SLAB, std::min<u64>(reserved, 16ull << 30)});
            void* q = nullptr;
            cudaError_t e = cudaMalloc(&q, sz);
            if (e != cudaSuccess && sz > bytes) { cudaGetLastError(); sz = bytes; e = cudaMalloc(&q, sz); }
            if (e != cudaSuccess) {
                cudaGetLastError();
                throw Error(-4 /*GRLGPU_ERR_NOMEM*/, "out of device memory: request of " + std::to_string(bytes) + " bytes with " +
                                                         std::to_string(reserved) + " reserved");
            }
            Slab sl; sl.base = (char*)q; sl.size = sz; sl.free_ranges[0] = sz;
            slabs.push_back(std::move(sl));
            reserved += sz;
        }
        throw Error(-4, "device pool: allocation failed");
    }
    void free(void* p) {
        auto it = live.find(p);
        if (it == live.end()) return;
        Slab& sl = slabs[it->second.first];
        u64 off = (u64)((char*)p - sl.base), len = it->second.second;
        in_use -= len;
        live.erase(it);
        auto nx = sl.free_ranges.lower_bound(off);
        if (nx != sl.free_ranges.end() && off + len == nx->first) { len += nx->second; nx = sl.free_ranges.erase(nx); }
        if (nx != sl.free_ranges.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == off) { pv->second += len; return; }
        }
        sl.free_ranges[off] = len;
    }
    void release_all() {  // caller guarantees the device is idle
        for (auto& sl : slabs) cudaFree(sl.base);
        slabs.clear(); live.clear(); reserved = in_use = 0;
    }
    ~DevicePool() { release_all(); }
};
inline thread_local DevicePool* g_pool = nullptr;  // set by the C ABI guard for the calling context

// RAII device buffer drawn from the context's DevicePool (plain cudaMalloc when no pool is active: self tests)
template <class T>
struct DevBuf {
    T* p = nullptr;
    u64 n = 0;
    cudaStream_t st = nullptr;
    DevicePool* pool = nullptr;
    DevBuf() = default;
    DevBuf(u64 count, cudaStream_t s) { alloc(count, s); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), st(o.st), pool(o.pool) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; st = o.st; pool = o.pool; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(u64 count, cudaStream_t s) {
        release();
        st = s;
        n = count;
        pool = g_pool;
        const u64 bytes = (count ? count : 1) * sizeof(T);
        if (pool) p = (T*)pool->alloc(bytes);
        else {
            void* q = nullptr;
            GRL_CUDA(cudaMalloc(&q, bytes));
            p = (T*)q;
        }
    }
    void zero() { GRL_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), st)); }
    void fill_ff() { GRL_CUDA(cudaMemsetAsync(p, 0xFF, (n ? n : 1) * sizeof(T), st)); }
    void release() {
        if (p) {
            if (pool) pool->free(p);
            else { cudaStreamSynchronize(st); cudaFree(p); }
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    u64 bytes() const { return n * sizeof(T); }
};

// Launch accounting: every kernel launch of the library goes through GRL_LAUNCH, which counts it
// and, when timing is enabled (grlgpu_profile_enable), brackets it with CUDA events on the launch
// stream so bench.py can report per-kernel durations measured live, outside any profiler.
struct Profiler {
    struct Rec { const char* name; cudaEvent_t a, b; u64 bytes; };
    struct Acc { u64 launches = 0; double ms = 0; u64 bytes = 0; };
    bool timing = false;
    u64 launches = 0;
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, Acc> acc;
    cudaEvent_t get_event() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        GRL_CUDA(cudaEventCreate(&e));
        return e;
    }
    void resolve() {  // call after a stream synchronisation point
        for (auto& r : pending) {
            float ms = 0;
            if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                Acc& a = acc[r.name];
                a.launches++; a.ms += ms; a.bytes += r.bytes;
            }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        pending.clear();
    }
    void reset() { resolve(); acc.clear(); launches = 0; }
    ~Profiler() { resolve(); for (auto e : pool) cudaEventDestroy(e); }
};
inline thread_local Profiler* g_prof = nullptr;

struct ProfScope {
    Profiler* p; cudaStream_t st; cudaEvent_t a{}, b{}; const char* name; u64 bytes;
    ProfScope(const char* nm, u64 by, cudaStream_t s) : p(g_prof), st(s), name(nm), bytes(by) {
        if (!p) return;
        p->launches++;
        if (p->timing) { a = p->get_event(); b = p->get_event(); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (p && p->timing) { cudaEventRecord(b, st); p->pending.push_back({name, a, b, bytes}); }
    }
};

// GRL_LAUNCH("name", expected_dram_bytes, (kernel<T...>), grid, block, smem, stream, args...)
#define GRL_LAUNCH(name, bytes, kernel, grid, block, smem, st, ...)            \
    do {                                                                       \
        grl::ProfScope _ps(name, (u64)(bytes), st);                            \
        kernel<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__);                \
        GRL_KERNEL_CHECK();                                                    \
    } while (0)

#ifdef __CUDACC__
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <class T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T x = __shfl_xor_sync(0xffffffffu, v, o);
        v = x > v ? x : v;
    }
    return v;
}
template <class T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T x = __shfl_xor_sync(0xffffffffu, v, o);
        v = x < v ? x : v;
    }
    return v;
}
// inclusive warp scan (sum)
template <class T>
__device__ __forceinline__ T warp_inclusive_sum(T v) {
    const u32 l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T x = __shfl_up_sync(0xffffffffu, v, o);
        if (l >= (u32)o) v += x;
    }
    return v;
}

// Block-wide exclusive sum for blocks of up to 1024 threads. Returns the exclusive prefix of `v`
// for the calling thread and the block total in `total`. `smem` must hold 33 T's.
template <class T>
__device__ __forceinline__ T block_exclusive_sum(T v, T* smem, T& total) {
    const u32 l = lane_id(), w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    T inc = warp_inclusive_sum(v);
    if (l == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = l < nw ? smem[l] : T(0);
        T xi = warp_inclusive_sum(x);
        smem[l] = xi - x;  // exclusive prefix of warp totals
        if (l == 31) smem[32] = xi;
    }
    __syncthreads();
    T res = smem[w] + inc - v;
    total = smem[32];
    __syncthreads();
    return res;
}

// 64-bit mixing (murmur3 fmix64). Hash values never reach the output.
__device__ __forceinline__ u64 mix64(u64 k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
#endif  // __CUDACC__

}  // namespace grl
