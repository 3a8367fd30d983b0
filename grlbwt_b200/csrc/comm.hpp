// Exchange layer of the multi-GPU parse rounds. Two backends behind one interface:
//   NcclComm   one rank per process OR per host thread; grouped ncclSend/ncclRecv = all-to-all-v over NVLink/NVSwitch.
//              NCCL is resolved at run time (dlopen of libnccl.so.2: the copy torch already loaded when the caller is
//              bench.py under torchrun, else the system library), so the library has no link-time dependency on it and
//              still loads on a box without NCCL.
//   IpcComm    one rank per PROCESS on one box (bench.py under torchrun): every rank stages what it sends in a window of device
//              memory that its peers have mapped through CUDA IPC, and the peers pull their pieces out of it with copy-engine
//              DMA over NVLink; the rendezvous (sizes, offsets, barriers, the small host-side all-gathers) lives in a POSIX
//              shared-memory segment, so an exchange costs microseconds of latency instead of an NCCL launch per step.
//   LocalComm  ranks are host threads of ONE process; every rank pulls its pieces straight out of the peers' send buffers
//              with cudaMemcpyPeerAsync (NVLink P2P between GPUs; a device-to-device copy when several ranks share one
//              GPU, which is how the N > 1 path is tested on a 1-GPU box). No NCCL involved.
// The reference's counterpart is the std::thread fan-out / serial join of mt_parse_strat_t (parsing_strategies.h:244-386).
#pragma once
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>  // types and prototypes only; every call goes through the table resolved below
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
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
    // tell the peers that this rank failed, where the backend can (ranks blocked in an exchange then fail instead of waiting)
    virtual void abort_group() {}
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
    void abort_group() override { g->abort(); }
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

// ---------------------------------------------------------------- one rank per process: shared-memory rendezvous + CUDA IPC windows
struct IpcComm : Comm {
    static constexpr int MAX_RANKS = 32, MAX_PARTS = 8, N_PULL_STREAMS = 4;
    static constexpr size_t SMALL_CAP = 64u << 10;
    static constexpr u32 MAGIC = 0x47524c31u;  // "GRL1"
    struct Header {
        std::atomic<u32> magic, count, sense, failed;
        u32 world, pad[11];
    };
    struct Slot {
        cudaIpcMemHandle_t handle;      // this rank's window (valid when window_bytes != 0)
        u64 window_bytes, need;         // need: bytes this rank stages in the exchange being set up
        u64 off[MAX_PARTS][MAX_RANKS];  // byte offset inside my window of array a's piece for destination p
        u64 len[MAX_PARTS][MAX_RANKS];  // its length in bytes
        alignas(64) char small[SMALL_CAP];
    };
    static_assert(std::atomic<u32>::is_always_lock_free, "the barrier lives in shared memory");
    static size_t shm_bytes(int w) { return sizeof(Header) + (size_t)w * sizeof(Slot); }

    std::string name;
    void* map = nullptr;
    size_t map_bytes = 0;
    Header* hdr = nullptr;
    Slot* slots = nullptr;
    u32 my_sense = 0;
    char* window = nullptr;
    u64 window_bytes = 0;
    std::vector<char*> peer_win;
    cudaStream_t pull_st[N_PULL_STREAMS] = {};
    double timeout_s = 300.0;

    bool use_cuda = true;  // false: the rendezvous only (barriers + small gathers), no device windows -- the CPU self-test of the segment code
    IpcComm(const char* session, int r, int w, int dev_, bool with_cuda = true) : name(session), peer_win((size_t)w, nullptr), use_cuda(with_cuda) {
        rank = r; world = w; device = dev_;
        if (w > MAX_RANKS) throw Error(-2, "at most 32 ranks");
        if (const char* e = getenv("GRLGPU_IPC_TIMEOUT_S")) timeout_s = std::max(1.0, atof(e));
        if (use_cuda) GRL_CUDA(cudaSetDevice(device));
        map_bytes = shm_bytes(w);
        const auto t0 = std::chrono::steady_clock::now();
        auto waited = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
        int fd = -1;
        if (rank == 0) {
            fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0) throw Error(-5, "shm_open(" + name + ") failed: " + strerror(errno));
            if (ftruncate(fd, (off_t)map_bytes) != 0) { close(fd); shm_unlink(name.c_str()); throw Error(-5, std::string("ftruncate of the rendezvous segment failed: ") + strerror(errno)); }
        } else {
            for (;;) {  // rank 0 may not have created (or sized) the segment yet
                fd = shm_open(name.c_str(), O_RDWR, 0600);
                if (fd >= 0) {
                    struct stat sb;
                    if (fstat(fd, &sb) == 0 && (size_t)sb.st_size == map_bytes) break;
                    close(fd);
                    fd = -1;
                }
                if (waited() > timeout_s) throw Error(-5, "rank 0 never created the rendezvous segment " + name);
                usleep(1000);
            }
        }
        map = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (map == MAP_FAILED) { map = nullptr; if (rank == 0) shm_unlink(name.c_str()); throw Error(-5, std::string("mmap of the rendezvous segment failed: ") + strerror(errno)); }
        hdr = (Header*)map;
        slots = (Slot*)((char*)map + sizeof(Header));
        if (rank == 0) {  // a fresh segment is zero-filled: counters, senses and window sizes start at 0
            hdr->world = (u32)world;
            hdr->magic.store(MAGIC, std::memory_order_release);
        } else {
            while (hdr->magic.load(std::memory_order_acquire) != MAGIC) {
                if (waited() > timeout_s) { release(); throw Error(-5, "the rendezvous segment was never initialised"); }
                usleep(200);
            }
            if (hdr->world != (u32)world) { release(); throw Error(-2, "ranks disagree on the world size"); }
        }
        if (use_cuda)
            for (auto& s : pull_st) GRL_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        try {
            barrier();
        } catch (...) {
            if (rank == 0) shm_unlink(name.c_str());
            release();
            throw;
        }
        if (rank == 0) shm_unlink(name.c_str());  // everyone has it mapped: the name can go, the memory lives until the last unmap
        // a first small window right away: a box where CUDA IPC does not work between these processes fails HERE, on every rank
        // together, where the caller can still choose another backend
        if (!use_cuda) return;
        try {
            regrow(16u << 20);
        } catch (...) {
            abort_group();
            release();
            throw;
        }
    }
    void release() {
        for (auto& p : peer_win) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
        if (window) { cudaFree(window); window = nullptr; }
        for (auto& s : pull_st) if (s) { cudaStreamDestroy(s); s = nullptr; }
        if (map) { munmap(map, map_bytes); map = nullptr; hdr = nullptr; }
        if (use_cuda) cudaGetLastError();
    }
    ~IpcComm() override {
        // peers must have closed their mappings of my window before it is freed: one last (short, best-effort) barrier in between
        for (auto& p : peer_win) if (p) { cudaIpcCloseMemHandle(p); p = nullptr; }
        if (hdr && !hdr->failed.load()) { timeout_s = std::min(timeout_s, 10.0); try { barrier(); } catch (...) {} }
        release();
    }
    const char* kind() const override { return "CUDA IPC windows pulled by copy-engine DMA over NVLink + shared-memory rendezvous"; }
    void fail(const std::string& why) {
        if (hdr) hdr->failed.store(1, std::memory_order_release);
        throw Error(-5, why);
    }
    void abort_group() override { if (hdr) hdr->failed.store(1, std::memory_order_release); }
    // sense-reversing barrier over the ranks (processes); every store before it is visible to every load after it
    void barrier() {
        my_sense ^= 1u;
        if (hdr->failed.load(std::memory_order_acquire)) throw Error(-5, "multi-GPU group aborted: another rank failed");
        if (hdr->count.fetch_add(1, std::memory_order_acq_rel) + 1 == (u32)world) {
            hdr->count.store(0, std::memory_order_relaxed);
            hdr->sense.store(my_sense, std::memory_order_release);
            return;
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (u64 spins = 0; hdr->sense.load(std::memory_order_acquire) != my_sense; spins++) {
            if (hdr->failed.load(std::memory_order_acquire)) throw Error(-5, "multi-GPU group aborted: another rank failed");
            if (spins < 4096) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                continue;
            }
            sched_yield();
            if ((spins & 1023) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
                fail("multi-GPU barrier timed out: a peer process is gone or stuck");
        }
    }
    void all_gather_host(const void* mine, size_t bytes, void* all, cudaStream_t) override {
        n_small++;
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t done = 0; done < bytes || done == 0; done += SMALL_CAP) {
            const size_t chunk = std::min(SMALL_CAP, bytes - done);
            memcpy(slots[rank].small, (const char*)mine + done, chunk);
            barrier();
            for (int p = 0; p < world; p++) memcpy((char*)all + (size_t)p * bytes + done, slots[p].small, chunk);
            barrier();
            if (bytes == 0) break;
        }
        ms_small += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    // every rank's window grows to the same size (collective: all ranks see the same `want`)
    void regrow(u64 want) {
        for (auto& p : peer_win) if (p) { GRL_CUDA(cudaIpcCloseMemHandle(p)); p = nullptr; }
        barrier();  // nobody maps my old window any more
        if (window) { GRL_CUDA(cudaFree(window)); window = nullptr; }
        window_bytes = (want + want / 4 + (1u << 20)) / 4096 * 4096;
        const cudaError_t e = cudaMalloc(&window, window_bytes);
        if (e != cudaSuccess) { cudaGetLastError(); fail("exchange window of " + std::to_string(window_bytes >> 20) + " MiB: " + cudaGetErrorString(e)); }
        GRL_CUDA(cudaIpcGetMemHandle(&slots[rank].handle, window));
        slots[rank].window_bytes = window_bytes;
        barrier();
        for (int q = 1; q < world; q++) {
            const int p = (rank + q) % world;
            void* ptr = nullptr;
            const cudaError_t eo = cudaIpcOpenMemHandle(&ptr, slots[p].handle, cudaIpcMemLazyEnablePeerAccess);
            if (eo != cudaSuccess) { cudaGetLastError(); fail(std::string("cudaIpcOpenMemHandle of a peer's window: ") + cudaGetErrorString(eo)); }
            peer_win[(size_t)p] = (char*)ptr;
        }
    }
    void all_to_all_v(const void* d_send, const u64* send_off, void* d_recv, const u64* recv_off, cudaStream_t st) override {
        std::vector<u64> sc((size_t)world), rc((size_t)world);
        for (int p = 0; p < world; p++) { sc[(size_t)p] = send_off[p + 1] - send_off[p]; rc[(size_t)p] = recv_off[p + 1] - recv_off[p]; }
        const SoaPart part{d_send, d_recv, 1};
        all_to_all_soa(&part, 1, sc.data(), rc.data(), st);
    }
    void all_to_all_soa(const SoaPart* parts, int n_parts, const u64* send_cnt, const u64* recv_cnt, cudaStream_t st) override {
        if (n_parts > MAX_PARTS) { Comm::all_to_all_soa(parts, n_parts, send_cnt, recv_cnt, st); return; }
        try {
            exchange(parts, n_parts, send_cnt, recv_cnt, st);
        } catch (...) {
            abort_group();  // (a CUDA error on this rank must not leave the peers in a barrier)
            throw;
        }
    }
    void exchange(const SoaPart* parts, int n_parts, const u64* send_cnt, const u64* recv_cnt, cudaStream_t st) {
        if (!use_cuda) throw Error(-1, "this communicator was created without device windows");
        n_bulk++;
        std::vector<u64> s_el((size_t)world + 1, 0), r_el((size_t)world + 1, 0);
        for (int p = 0; p < world; p++) { s_el[(size_t)p + 1] = s_el[(size_t)p] + send_cnt[p]; r_el[(size_t)p + 1] = r_el[(size_t)p] + recv_cnt[p]; }
        if (send_cnt[rank] != recv_cnt[rank]) fail("all-to-all-v: send and receive sizes disagree");
        Slot& me = slots[rank];
        u64 total = 0;
        for (int a = 0; a < n_parts; a++)
            for (int p = 0; p < world; p++) {
                const u64 len = p == rank ? 0 : send_cnt[p] * parts[a].elem_bytes;
                me.off[a][p] = total;
                me.len[a][p] = len;
                total += (len + 255) / 256 * 256;
                bytes_sent += len;
            }
        me.need = total;
        // the staging copies below are part of this rank's "send": the clock starts when its data is ready, as for the other backends
        GRL_CUDA(cudaStreamSynchronize(st));
        const auto t0 = std::chrono::steady_clock::now();
        barrier();
        u64 want = 0;
        for (int p = 0; p < world; p++) want = std::max(want, slots[p].need);
        if (want > window_bytes) regrow(want);
        for (int a = 0; a < n_parts; a++) {
            const u64 eb = parts[a].elem_bytes;
            for (int p = 0; p < world; p++)
                if (me.len[a][p]) GRL_CUDA(cudaMemcpyAsync(window + me.off[a][p], (const char*)parts[a].send + s_el[(size_t)p] * eb, me.len[a][p], cudaMemcpyDeviceToDevice, st));
            if (send_cnt[rank])
                GRL_CUDA(cudaMemcpyAsync((char*)parts[a].recv + r_el[(size_t)rank] * eb, (const char*)parts[a].send + s_el[(size_t)rank] * eb, send_cnt[rank] * eb,
                                         cudaMemcpyDeviceToDevice, st));
        }
        GRL_CUDA(cudaStreamSynchronize(st));
        barrier();  // every window is staged
        int k = 0;
        for (int q = 1; q < world; q++) {
            const int p = (rank + q) % world;  // staggered: at any moment the ranks pull from different peers
            const Slot& from = slots[p];
            for (int a = 0; a < n_parts; a++) {
                const u64 len = from.len[a][rank];
                if (len != recv_cnt[p] * parts[a].elem_bytes) fail("all-to-all-v: send and receive sizes disagree");
                if (!len) continue;
                GRL_CUDA(cudaMemcpyAsync((char*)parts[a].recv + r_el[(size_t)p] * parts[a].elem_bytes, peer_win[(size_t)p] + from.off[a][rank], len, cudaMemcpyDefault,
                                         pull_st[k++ % N_PULL_STREAMS]));
            }
        }
        for (auto& s : pull_st) GRL_CUDA(cudaStreamSynchronize(s));
        barrier();  // every peer has pulled what it needed: the windows may be overwritten
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
