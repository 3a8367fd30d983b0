// Shared device/host helpers for the grlgpu kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
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

// Device memory of one context: ONE contiguous virtual address range (CUDA virtual memory
// management: cuMemAddressReserve + cuMemCreate/cuMemMap) whose physical backing grows in 256 MB
// chunks up to what the rounds need, carved by a best-fit free list with coalescing. Every buffer
// of every round comes from here, so after the first step no allocation reaches the driver and a
// multi-GB buffer never fails because of fragmentation across separately malloc'ed slabs (HBM is
// laid out once and reused round after round). All work of a context is issued on one stream, so a
// block freed on the host may be handed out again immediately: the kernels that used it were
// enqueued before the kernels that will. The driver entry points are resolved through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda and still loads on a CPU-only box.
struct DevicePool {
    typedef CUresult (*fn_reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
    typedef CUresult (*fn_create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
    typedef CUresult (*fn_map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    typedef CUresult (*fn_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
    typedef CUresult (*fn_unmap)(CUdeviceptr, size_t);
    typedef CUresult (*fn_release)(CUmemGenericAllocationHandle);
    typedef CUresult (*fn_addrfree)(CUdeviceptr, size_t);
    typedef CUresult (*fn_gran)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
    fn_reserve p_reserve = nullptr; fn_create p_create = nullptr; fn_map p_map = nullptr; fn_access p_access = nullptr;
    fn_unmap p_unmap = nullptr; fn_release p_release = nullptr; fn_addrfree p_addrfree = nullptr; fn_gran p_gran = nullptr;

    int device = 0;
    CUdeviceptr base = 0;
    u64 va_size = 0, mapped = 0, chunk = 0;
    std::vector<CUmemGenericAllocationHandle> handles;
    std::map<u64, u64> free_ranges;       // offset -> length, inside [0, mapped)
    std::map<u64, u64> live;              // offset -> length
    u64 in_use = 0, peak = 0;
    double grow_ms = 0;                   // host time spent mapping physical memory (GRLGPU_TRACE prints it when the pool is released)
    std::vector<int> peers;               // other devices that may read / write this pool directly (NVLink P2P, LocalComm pulls)
    static constexpr u64 ALIGN = 512;

    template <class F>
    static void resolve(const char* name, F& f) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint(name, &q, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !q)
            throw Error(-3, std::string("driver entry point unavailable: ") + name);
        f = (F)q;
    }
    // Mapped memory of destroyed contexts, kept per device for the next context of this process (mapping 80 GB costs 0.3-0.9 s, more
    // than a whole parse phase): a context that is created later adopts it instead of mapping afresh. GRLGPU_POOL_CACHE=0 turns it
    // off; grlgpu_trim() hands everything back; an allocation that would otherwise fail releases the device's cached mappings first.
    struct Spare {
        int device; CUdeviceptr base; u64 va_size, mapped, chunk;
        std::vector<CUmemGenericAllocationHandle> handles;
        std::vector<int> peers;
    };
    struct SpareCache {
        std::mutex m;
        std::vector<Spare> spares;
    };
    static SpareCache& cache() { static SpareCache* c = new SpareCache(); return *c; }  // leaked on purpose: no driver calls during static destruction
    static bool cache_enabled() { static const bool v = [] { const char* e = getenv("GRLGPU_POOL_CACHE"); return !e || atoi(e) != 0; }(); return v; }
    bool adopt_spare() {
        SpareCache& c = cache();
        std::lock_guard<std::mutex> lk(c.m);
        for (size_t i = 0; i < c.spares.size(); i++) {
            if (c.spares[i].device != device) continue;
            Spare sp = std::move(c.spares[i]);
            c.spares.erase(c.spares.begin() + (long)i);
            base = sp.base; va_size = sp.va_size; mapped = sp.mapped; chunk = sp.chunk;
            handles = std::move(sp.handles); peers = std::move(sp.peers);
            free_ranges.clear(); live.clear();
            if (mapped) free_ranges[0] = mapped;
            in_use = 0; peak = 0;
            return true;
        }
        return false;
    }
    void release_spare(Spare& sp) {
        for (size_t i = 0; i < sp.handles.size(); i++) { p_unmap(sp.base + (u64)i * sp.chunk, sp.chunk); p_release(sp.handles[i]); }
        p_addrfree(sp.base, sp.va_size);
    }
    // hand the cached mappings of `dev` (or of every device: dev < 0) back to the driver; returns the bytes released
    u64 trim_spares(int dev) {
        SpareCache& c = cache();
        std::lock_guard<std::mutex> lk(c.m);
        u64 freed = 0;
        for (size_t i = 0; i < c.spares.size();) {
            if (dev >= 0 && c.spares[i].device != dev) { i++; continue; }
            freed += c.spares[i].mapped;
            release_spare(c.spares[i]);
            c.spares.erase(c.spares.begin() + (long)i);
        }
        return freed;
    }
    void init(int dev) {
        if (base) return;
        device = dev;
        resolve("cuMemAddressReserve", p_reserve); resolve("cuMemCreate", p_create); resolve("cuMemMap", p_map);
        resolve("cuMemSetAccess", p_access); resolve("cuMemUnmap", p_unmap); resolve("cuMemRelease", p_release);
        resolve("cuMemAddressFree", p_addrfree); resolve("cuMemGetAllocationGranularity", p_gran);
        if (cache_enabled() && adopt_spare()) return;
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        size_t gran = 0;
        if (p_gran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = 2ull << 20;
        chunk = ((256ull << 20) + gran - 1) / gran * gran;
        size_t free_b = 0, total_b = 0;
        GRL_CUDA(cudaMemGetInfo(&free_b, &total_b));
        va_size = ((u64)total_b + chunk - 1) / chunk * chunk;   // never more physical memory than the device has
        if (p_reserve(&base, va_size, 0, 0, 0) != CUDA_SUCCESS) { base = 0; throw Error(-3, "cuMemAddressReserve failed"); }
    }
    void grow(u64 need_bytes) {  // map more physical chunks so that a free range of need_bytes exists at the top
        u64 top_free = 0;
        if (!free_ranges.empty()) {
            auto last = std::prev(free_ranges.end());
            if (last->first + last->second == mapped) top_free = last->second;
        }
        u64 add = need_bytes > top_free ? need_bytes - top_free : 0;
        add = (add + chunk - 1) / chunk * chunk;
        if (mapped + add > va_size)
            throw Error(-4, "out of device memory: request of " + std::to_string(need_bytes) + " bytes with " + std::to_string(mapped) + " mapped");
        CUmemAllocationProp prop = {};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = device;
        std::vector<CUmemAccessDesc> acc = access_descs();
        const auto t_grow = std::chrono::steady_clock::now();
        struct GrowClock {
            double& ms; std::chrono::steady_clock::time_point t0;
            ~GrowClock() { ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
        } gclk{grow_ms, t_grow};
        for (u64 done = 0; done < add; done += chunk) {
            CUmemGenericAllocationHandle h;
            CUresult cr = p_create(&h, chunk, &prop, 0);
            if (cr != CUDA_SUCCESS && trim_spares(device)) cr = p_create(&h, chunk, &prop, 0);  // memory parked by earlier contexts of this process
            if (cr != CUDA_SUCCESS)
                throw Error(-4, "out of device memory: request of " + std::to_string(need_bytes) + " bytes with " + std::to_string(mapped) + " mapped");
            if (p_map(base + mapped, chunk, 0, h, 0) != CUDA_SUCCESS || p_access(base + mapped, chunk, acc.data(), acc.size()) != CUDA_SUCCESS) {
                p_release(h);
                throw Error(-3, "cuMemMap failed");
            }
            handles.push_back(h);
            add_free(mapped, chunk);
            mapped += chunk;
        }
    }
    std::vector<CUmemAccessDesc> access_descs() const {
        std::vector<CUmemAccessDesc> v;
        CUmemAccessDesc a = {};
        a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        a.location.id = device;
        a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        v.push_back(a);
        for (int d : peers) { a.location.id = d; v.push_back(a); }
        return v;
    }
    // devices whose copy engines / kernels may touch this pool over NVLink (only those that report peer capability)
    void set_peers(const std::vector<int>& devs) {
        peers.clear();
        for (int d : devs) {
            int can = 0;
            if (d != device && cudaDeviceCanAccessPeer(&can, d, device) == cudaSuccess && can && std::find(peers.begin(), peers.end(), d) == peers.end()) peers.push_back(d);
        }
        if (mapped && !peers.empty()) {
            std::vector<CUmemAccessDesc> acc = access_descs();
            if (p_access(base, mapped, acc.data(), acc.size()) != CUDA_SUCCESS) peers.clear();  // stay correct (staged copies) if the mapping is refused
        }
    }
    void add_free(u64 off, u64 len) {
        auto nx = free_ranges.lower_bound(off);
        if (nx != free_ranges.end() && off + len == nx->first) { len += nx->second; nx = free_ranges.erase(nx); }
        if (nx != free_ranges.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == off) { pv->second += len; return; }
        }
        free_ranges[off] = len;
    }
    void* alloc(u64 bytes) {
        bytes = (std::max<u64>(bytes, 1) + ALIGN - 1) / ALIGN * ALIGN;
        for (int pass = 0; pass < 2; pass++) {
            u64 best_off = 0, best_len = ~0ULL;
            for (auto& fr : free_ranges)
                if (fr.second >= bytes && fr.second < best_len) { best_off = fr.first; best_len = fr.second; }
            if (best_len != ~0ULL) {
                free_ranges.erase(best_off);
                if (best_len > bytes) free_ranges[best_off + bytes] = best_len - bytes;
                live[best_off] = bytes;
                in_use += bytes; peak = std::max(peak, in_use);
                return (void*)(base + best_off);
            }
            grow(bytes);
        }
        throw Error(-4, "device pool: allocation failed");
    }
    void free(void* p) {
        const u64 off = (u64)((CUdeviceptr)p - base);
        auto it = live.find(off);
        if (it == live.end()) return;
        const u64 len = it->second;
        live.erase(it);
        in_use -= len;
        add_free(off, len);
    }
    void release_all() {  // caller guarantees the device is idle
        if (!base) return;
        if (getenv("GRLGPU_TRACE")) fprintf(stderr, "[grlgpu] device %d pool: %.1f GB mapped in %zu chunks (%zu peer device%s), %.1f ms spent mapping\n", device, mapped / 1e9,
                                            handles.size(), peers.size(), peers.size() == 1 ? "" : "s", grow_ms);
        for (size_t i = 0; i < handles.size(); i++) { p_unmap(base + (u64)i * chunk, chunk); p_release(handles[i]); }
        p_addrfree(base, va_size);
        handles.clear(); free_ranges.clear(); live.clear();
        base = 0; mapped = 0; in_use = 0;
    }
    ~DevicePool() {
        if (base && mapped && live.empty() && cache_enabled()) {  // park the mapping for the next context on this device
            SpareCache& c = cache();
            std::lock_guard<std::mutex> lk(c.m);
            size_t same = 0;
            for (const Spare& sp : c.spares) same += sp.device == device;
            if (same < 8) {
                if (getenv("GRLGPU_TRACE")) fprintf(stderr, "[grlgpu] device %d pool: %.1f GB kept mapped for the next context (%.1f ms spent mapping)\n", device, mapped / 1e9, grow_ms);
                c.spares.push_back(Spare{device, base, va_size, mapped, chunk, std::move(handles), std::move(peers)});
                base = 0; mapped = 0; handles.clear(); free_ranges.clear();
                return;
            }
        }
        release_all();
    }
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

// Small device -> host readbacks (counts, flags) between the stages of a round. They go through a kernel that writes
// to mapped pinned memory instead of a copy-engine transfer: a DMA job would queue behind the multi-GB asynchronous
// level copies of grlgpu_fetch_level_async and stall the round for their whole duration.
static __global__ void copy_small_kernel(const u32* __restrict__ src, u32* __restrict__ dst, u32 n_words) {
    if (threadIdx.x < n_words) dst[threadIdx.x] = src[threadIdx.x];
}
inline void d2h_small(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t st) {
    static thread_local u32* slots = nullptr;
    if (bytes % 4 || bytes > 256) throw Error(GRLGPU_ERR_ARG, "d2h_small: unsupported size");
    if (!slots) GRL_CUDA(cudaHostAlloc((void**)&slots, 256, cudaHostAllocMapped | cudaHostAllocPortable));
    copy_small_kernel<<<1, 64, 0, st>>>((const u32*)dev_src, slots, (u32)(bytes / 4));
    GRL_CUDA(cudaGetLastError());
    GRL_CUDA(cudaStreamSynchronize(st));
    memcpy(host_dst, slots, bytes);
}


// the same for readbacks of up to a few hundred KB (per-peer counts, key samples, gathered scalars of the multi-GPU rounds)
static __global__ void copy_words_kernel(const u32* __restrict__ src, u32* __restrict__ dst, u64 n_words) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (u64)gridDim.x * blockDim.x) dst[i] = src[i];
}
inline void d2h_mapped(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t st) {
    static thread_local u8* buf = nullptr;
    static thread_local size_t cap = 0;
    if (bytes == 0) return;
    if (bytes % 4 || ((uintptr_t)dev_src & 3)) throw Error(GRLGPU_ERR_ARG, "d2h_mapped: unsupported size or alignment");
    if (bytes > cap) {
        if (buf) GRL_CUDA(cudaFreeHost(buf));
        cap = std::max<size_t>(bytes, (size_t)1 << 20);
        GRL_CUDA(cudaHostAlloc((void**)&buf, cap, cudaHostAllocMapped | cudaHostAllocPortable));
    }
    const u64 words = bytes / 4;
    copy_words_kernel<<<(unsigned)std::min<u64>(64, (words + 255) / 256), 256, 0, st>>>((const u32*)dev_src, (u32*)buf, words);
    GRL_CUDA(cudaGetLastError());
    GRL_CUDA(cudaStreamSynchronize(st));
    memcpy(host_dst, buf, bytes);
}

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

// Random gathers: a plain load that misses the L2 pulls the whole 128-byte line from DRAM (4 sectors looked up per
// 1-sector request in ncu); the 64-byte prefetch-size hint halves that for records read once at random.
__device__ __forceinline__ ulonglong2 ld_gather16(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L2::64B.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_gather4(const u32* p) {
    u32 v;
    asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 ld_gather8(const u64* p) {
    u64 v;
    asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

template <class T>
__device__ __forceinline__ T ld_gather(const T* p) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "ld_gather: 4- or 8-byte types");
    if constexpr (sizeof(T) == 4) return (T)ld_gather4(reinterpret_cast<const u32*>(p));
    else return (T)ld_gather8(reinterpret_cast<const u64*>(p));
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
