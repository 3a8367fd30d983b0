// Shared device/host helpers for the grlgpu kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

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

// RAII device buffer (stream-ordered pool allocation; the pool keeps freed blocks, so per-round
// allocation does not go back to the driver).
template <class T>
struct DevBuf {
    T* p = nullptr;
    u64 n = 0;
    cudaStream_t st = nullptr;
    DevBuf() = default;
    DevBuf(u64 count, cudaStream_t s) { alloc(count, s); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), st(o.st) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; st = o.st; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(u64 count, cudaStream_t s) {
        release();
        st = s;
        n = count;
        void* q = nullptr;
        GRL_CUDA(cudaMallocAsync(&q, (count ? count : 1) * sizeof(T), s));
        p = (T*)q;
    }
    void zero() { GRL_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), st)); }
    void fill_ff() { GRL_CUDA(cudaMemsetAsync(p, 0xFF, (n ? n : 1) * sizeof(T), st)); }
    void release() {
        if (p) cudaFreeAsync(p, st);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    u64 bytes() const { return n * sizeof(T); }
};

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
