// Exchange layer of the multi-GPU parse rounds. Two backends behind one interface:
//   NcclComm   one rank per process OR per host thread; grouped ncclSend/ncclRecv = all-to-all-v over NVLink/NVSwitch.
//              NCCL is resolved at run time (dlopen of libnccl.so.2: the copy torch already loaded when the caller is
//              bench.py under torchrun, else the system library), so the library has no link-time dependency on it and
//              still loads on a box without NCCL.
//   LocalComm  ranks are host threads of ONE process; every rank pulls its pieces straight out of the peers' send buffers
//              with cudaMemcpyPeerAsync (NVLink P2P between GPUs; a device-to-device copy when several ranks share one
//              GPU, which is how the N > 1 path is tested on a 1-GPU box). No NCCL involved.
// The reference's counterpart is the std::thread fan-out / serial join of mt_parse_strat_t (parsing_strategies.h:244-386).
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; every call goes through the table resolved below

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <vector>

#include "util.cuh"

namespace grl {

struct Comm {
    int rank = 0, world = 1, device = 0;
    u64 bytes_sent = 0, n_bulk = 0, n_small = 0;  // accounting for bench.py: bulk bytes this rank sent, collectives issued
    double ms_bulk = 0, ms_small = 0;             // host wall time spent inside them (each call ends with the data in place)
    virtual ~Comm() {}
    virtual const char* kind() const = 0;
    // small host-side all-gather: all[p * bytes .. (p+1) * bytes) = rank p's `mine`. Blocking.
    virtual void all_gather_host(const void* mine, size_t bytes, void* all, cudaStream_t st) = 0;
    // device all-to-all-v in BYTES: send_off / recv_off are world + 1 prefix offsets (by destination / by source).
    // On return the received data is visible to work enqueued on `st` afterwards and d_send may be reused by it.
    virtual void all_to_all_v(const void* d_send, const u64* send_off, void* d_recv, const u64* recv_off, cudaStream_t st) = 0;
    // several arrays that share the per-peer ELEMENT counts (a structure of arrays): one exchange instead of one per array
    struct SoaPart { const void* send; void* recv; u64 elem_bytes; };
    virtual void all_to_all_soa(const SoaPart* parts, int n_parts, const u64* send_cnt, const u64* recv_cnt, cudaStream_t st) {
        std::vector<u64> so((size_t)world + 1), ro((size_t)world + 1);
        for (int a = 0; a < n_parts; a++) {
            so[0] = ro[0] = 0;
            for (int p = 0; p < world; p++) { so[(size_t)p + 1] = so[(size_t)p] + send_cnt[p] * parts[a].elem_bytes; ro[(size_t)p + 1] = ro[(size_t)p] + recv_cnt[p] * parts[a].elem_bytes; }
            all_to_all_v(parts[a].send, so.data(), parts[a].recv, ro.data(), st);
        }
    }
};

// ---------------------------------------------------------------- in-process ranks (host threads)
struct LocalGroup {
    int world;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    u64 gen = 0;
    bool failed = false;
    std::vector<const void*> ptr;
    std::vector<const u64*> off;
    std::vector<int> dev;
    explicit LocalGroup(int w) : world(w), ptr((size_t)w, nullptr), off((size_t)w, nullptr), dev((size_t)w, 0) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        if (failed) throw Error(-5, "multi-GPU group aborted: another rank failed");
        const u64 g = gen;
        if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g || failed; });
        if (failed) throw Error(-5, "multi-GPU group aborted: another rank failed");
    }
    void abort() {
        std::lock_guard<std::mutex> lk(m);
        failed = true;
        cv.notify_all();
    }
};

struct LocalComm : Comm {
    LocalGroup* g;
    LocalComm(LocalGroup* grp, int r, int dev_) : g(grp) { rank = r; world = grp->world; device = dev_; }
    const char* kind() const override { return "in-process ranks, cudaMemcpyPeerAsync pulls"; }
    void all_gather_host(const void* mine, size_t bytes, void* all, cudaStream_t) override {
        n_small++;
        const auto t0 = std::chrono::steady_clock::now();
        g->ptr[(size_t)rank] = mine;
        g->wait();
        for (int p = 0; p < world; p++) memcpy((char*)all + (size_t)p * bytes, g->ptr[(size_t)p], bytes);
        g->wait();
        ms_small += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void all_to_all_v(const void* d_send, const u64* send_off, void* d_recv, const u64* recv_off, cudaStream_t st) override {
        n_bulk++;
        bytes_sent += send_off[world] - (send_off[rank + 1] - send_off[rank]);
        GRL_CUDA(cudaStreamSynchronize(st));  // my send buffer is complete
        const auto t0 = std::chrono::steady_clock::now();
        g->ptr[(size_t)rank] = d_send;
        g->off[(size_t)rank] = send_off;
        g->dev[(size_t)rank] = device;
        g->wait();
        for (int q = 0; q < world; q++) {
            const int p = (rank + q) % world;  // start with myself, then round-robin so the pulls spread over the peers
            const u64* po = g->off[(size_t)p];
            const u64 len = po[rank + 1] - po[rank];
            if (len != recv_off[p + 1] - recv_off[p]) { g->abort(); throw Error(-5, "all-to-all-v: send and receive sizes disagree"); }
            if (!len) continue;
            const char* src = (const char*)g->ptr[(size_t)p] + po[rank];
            char* dst = (char*)d_recv + recv_off[p];
            if (g->dev[(size_t)p] == device) GRL_CUDA(cudaMemcpyAsync(dst, src, len, cudaMemcpyDeviceToDevice, st));
            else GRL_CUDA(cudaMemcpyPeerAsync(dst, device, src, g->dev[(size_t)p], len, st));
        }
        GRL_CUDA(cudaStreamSynchronize(st));
        g->wait();  // every peer has pulled what it needed out of my send buffer
        ms_bulk += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

// ---------------------------------------------------------------- NCCL, resolved at run time
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string where;
    static NcclApi& get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            // the copy already mapped into this process (torch's bundled NCCL under bench.py) wins over the system one:
            // two NCCL builds in one process would each bring their own global state
            api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
            api.where = "libnccl.so.2 already loaded by the host process";
            if (!api.h) { api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); api.where = "libnccl.so.2 (system)"; }
            if (!api.h) { api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); api.where = "libnccl.so (system)"; }
            if (!api.h) return;
            auto sym = [&](const char* n) { return dlsym(api.h, n); };
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.Send = (decltype(api.Send))sym("ncclSend");
            api.Recv = (decltype(api.Recv))sym("ncclRecv");
            api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
            api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        });
        if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllGather || !api.GroupStart || !api.GroupEnd)
            throw Error(-3, "NCCL is not available (dlopen libnccl.so.2 failed or symbols are missing)");
        return api;
    }
};

#define GRL_NCCL(expr)                                                                                                          \
    do {                                                                                                                        \
        ncclResult_t _r = (expr);                                                                                               \
        if (_r != ncclSuccess)                                                                                                  \
            throw grl::Error(-3, std::string(#expr) + ": " + (api.GetErrorString ? api.GetErrorString(_r) : "NCCL error") + " (" + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");                                          \
    } while (0)

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    void* d_stage = nullptr;
    size_t stage_cap = 0;
    NcclComm(const void* id128, int r, int w, int dev_) {
        rank = r; world = w; device = dev_;
        NcclApi& api = NcclApi::get();
        ncclUniqueId id;
        static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
        memcpy(&id, id128, sizeof(id));
        GRL_CUDA(cudaSetDevice(device));
        GRL_NCCL(api.CommInitRank(&comm, world, id, rank));
    }
    ~NcclComm() override {
        if (d_stage) cudaFree(d_stage);
        if (comm) { NcclApi& api = NcclApi::get(); if (api.CommDestroy) api.CommDestroy(comm); }
    }
    const char* kind() const override { return "NCCL grouped send/recv (all-to-all-v) + all-gather"; }
    void all_gather_host(const void* mine, size_t bytes, void* all, cudaStream_t st) override {
        n_small++;
        const auto t0 = std::chrono::steady_clock::now();
        NcclApi& api = NcclApi::get();
        const size_t padded = (bytes + 15) / 16 * 16, need = padded * (size_t)(world + 1);
        if (need > stage_cap) {
            if (d_stage) GRL_CUDA(cudaFree(d_stage));
            GRL_CUDA(cudaMalloc(&d_stage, need * 2));
            stage_cap = need * 2;
        }
        char* d_in = (char*)d_stage;
        char* d_out = d_in + padded;
        GRL_CUDA(cudaMemcpyAsync(d_in, mine, bytes, cudaMemcpyHostToDevice, st));
        GRL_NCCL(api.AllGather(d_in, d_out, padded, ncclUint8, comm, st));
        std::vector<char> tmp(padded * (size_t)world);
        d2h_mapped(tmp.data(), d_out, tmp.size(), st);  // through mapped pinned memory: a DMA readback would queue behind the level copies
        for (int p = 0; p < world; p++) memcpy((char*)all + (size_t)p * bytes, tmp.data() + (size_t)p * padded, bytes);
        ms_small += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void all_to_all_v(const void* d_send, const u64* send_off, void* d_recv, const u64* recv_off, cudaStream_t st) override {
        n_bulk++;
        bytes_sent += send_off[world] - (send_off[rank + 1] - send_off[rank]);
        GRL_CUDA(cudaStreamSynchronize(st));  // (for the accounting below: the clock starts when this rank's data is ready)
        const auto t0 = std::chrono::steady_clock::now();
        NcclApi& api = NcclApi::get();
        const u64 self = send_off[rank + 1] - send_off[rank];
        if (self != recv_off[rank + 1] - recv_off[rank]) throw Error(-5, "all-to-all-v: send and receive sizes disagree");
        if (self) GRL_CUDA(cudaMemcpyAsync((char*)d_recv + recv_off[rank], (const char*)d_send + send_off[rank], self, cudaMemcpyDeviceToDevice, st));
        GRL_NCCL(api.GroupStart());
        for (int q = 1; q < world; q++) {
            const int to = (rank + q) % world, from = (rank - q + world) % world;
            const u64 sl = send_off[to + 1] - send_off[to], rl = recv_off[from + 1] - recv_off[from];
            if (sl) GRL_NCCL(api.Send((const char*)d_send + send_off[to], sl, ncclUint8, to, comm, st));
            if (rl) GRL_NCCL(api.Recv((char*)d_recv + recv_off[from], rl, ncclUint8, from, comm, st));
        }
        GRL_NCCL(api.GroupEnd());
        GRL_CUDA(cudaStreamSynchronize(st));
        ms_bulk += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    // all arrays of the structure in ONE NCCL group: one rendezvous per peer instead of one per array
    void all_to_all_soa(const SoaPart* parts, int n_parts, const u64* send_cnt, const u64* recv_cnt, cudaStream_t st) override {
        n_bulk++;
        GRL_CUDA(cudaStreamSynchronize(st));
        const auto t0 = std::chrono::steady_clock::now();
        NcclApi& api = NcclApi::get();
        std::vector<u64> s_el((size_t)world + 1, 0), r_el((size_t)world + 1, 0);
        for (int p = 0; p < world; p++) { s_el[(size_t)p + 1] = s_el[(size_t)p] + send_cnt[p]; r_el[(size_t)p + 1] = r_el[(size_t)p] + recv_cnt[p]; }
        if (send_cnt[rank] != recv_cnt[rank]) throw Error(-5, "all-to-all-v: send and receive sizes disagree");
        for (int a = 0; a < n_parts; a++) {
            const u64 eb = parts[a].elem_bytes;
            bytes_sent += (s_el[(size_t)world] - send_cnt[rank]) * eb;
            if (send_cnt[rank])
                GRL_CUDA(cudaMemcpyAsync((char*)parts[a].recv + r_el[(size_t)rank] * eb, (const char*)parts[a].send + s_el[(size_t)rank] * eb, send_cnt[rank] * eb,
                                         cudaMemcpyDeviceToDevice, st));
        }
        GRL_NCCL(api.GroupStart());
        for (int q = 1; q < world; q++) {
            const int to = (rank + q) % world, from = (rank - q + world) % world;
            for (int a = 0; a < n_parts; a++) {
                const u64 eb = parts[a].elem_bytes;
                if (send_cnt[to]) GRL_NCCL(api.Send((const char*)parts[a].send + s_el[(size_t)to] * eb, send_cnt[to] * eb, ncclUint8, to, comm, st));
                if (recv_cnt[from]) GRL_NCCL(api.Recv((char*)parts[a].recv + r_el[(size_t)from] * eb, recv_cnt[from] * eb, ncclUint8, from, comm, st));
            }
        }
        GRL_NCCL(api.GroupEnd());
        GRL_CUDA(cudaStreamSynchronize(st));
        ms_bulk += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

}  // namespace grl
