// NCCL plumbing for the sharded apply: one process per GPU, the output-node list of every
// refinement iteration is split across ranks, norms and output coefficient blocks are exchanged over
// NVLink. NCCL is loaded at run time (dlopen of libnccl.so.2: the copy the host framework already loaded,
// else the system one), so the library has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

// minimal NCCL ABI (nccl.h 2.x): opaque comm, 128-byte unique id, enums as ints
struct NcclUniqueId {
    char internal[128];
};
using ncclComm_t = void *;
constexpr int kNcclInt8 = 0;    // ncclInt8 / ncclChar
constexpr int kNcclFloat64 = 8; // ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        // 1. the copy the host framework already mapped (two NCCL versions in one process do not mix: the
        //    second consumer would bind to the first copy by SONAME), 2. $MRX_NCCL_LIB, 3. the system library
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!a.lib) {
            const char *env = getenv("MRX_NCCL_LIB");
            if (env && env[0]) a.lib = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        }
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (a.lib) break;
            a.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        }
        if (!a.lib) MRX_ABORT(std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
        auto sym = [&](const char *s) {
            void *p = dlsym(a.lib, s);
            if (!p) MRX_ABORT(std::string("NCCL symbol missing: ") + s);
            return p;
        };
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
        a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return a;
}

void check(int rc, const char *what) {
    if (rc != 0) MRX_ABORT(std::string("NCCL error in ") + what + ": " + api().GetErrorString(rc));
}

} // namespace

} // namespace mrx

struct mrx_comm {
    int rank = 0, world = 1;
    mrx::ncclComm_t comm = nullptr;
    // ---- peer-memory exchange of output coefficient blocks (NVLink / NVSwitch, copy engines) ----
    // One cudaMalloc per rank holds kStageBufs staging buffers; its IPC handle is opened by every peer, so a rank PUSHES the
    // blocks it computed straight into the peers' HBM with DMA copies that use no SM and overlap the next refinement
    // iteration's kernels. peerBase[r] = base address of rank r's allocation in this process (own rank: local pointer).
    static constexpr int kStageBufs = 3;
    char *stageBase = nullptr;
    size_t stageBytes = 0; // per buffer
    std::vector<char *> peerBase;
    bool ipcTried = false, ipcOk = false;
    cudaStream_t pushStream = nullptr;
    // unpacking of an iteration's rows into the node store runs on its own stream, beside the next iteration's kernels
    cudaStream_t unpackStream = nullptr;
    cudaEvent_t evGathered = nullptr, evUnpacked[kStageBufs] = {nullptr, nullptr, nullptr};
    cudaEvent_t evReduced[kStageBufs] = {nullptr, nullptr, nullptr}, evPushed[kStageBufs] = {nullptr, nullptr, nullptr};
    char *ipcScratch = nullptr; // device buffer for the handle all-gather
    bool hostArena = false;     // this communicator's ranks share the process-wide host arena (mrx_comm_host_arena)
};

namespace mrx {

int shard_block() {
    static const int B = [] {
        const char *e = getenv("MRX_SHARD_BLOCK");
        const int b = e ? atoi(e) : 1;
        return b < 1 ? 1 : b;
    }();
    return B;
}
int comm_rank(const mrx_comm *c) { return c ? c->rank : 0; }
int comm_world(const mrx_comm *c) { return c ? c->world : 1; }

/// all-gather-v on device memory: segment r (count[r] bytes at base + off[r]) is broadcast from rank r
void comm_allgatherv(const mrx_comm *c, void *base, const size_t *off, const size_t *count, cudaStream_t st) {
    NcclApi &a = api();
    check(a.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < c->world; r++) {
        if (count[r] == 0) continue;
        char *p = static_cast<char *>(base) + off[r];
        check(a.Broadcast(p, p, count[r], kNcclInt8, r, c->comm, st), "ncclBroadcast");
    }
    check(a.GroupEnd(), "ncclGroupEnd");
}

/// in-place all-gather of equal segments: rank r's `bytes` live at base + r * bytes
void comm_allgather(const mrx_comm *c, void *base, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return;
    check(api().AllGather(static_cast<char *>(base) + (size_t)c->rank * bytes, base, bytes, kNcclInt8, c->comm, st), "ncclAllGather");
}

bool comm_peer_push_enabled(const mrx_comm *c) { return c && c->ipcOk; }
char *comm_stage(const mrx_comm *c, int buf) { return c->stageBase + (size_t)buf * c->stageBytes; }
size_t comm_stage_bytes(const mrx_comm *c) { return c->stageBytes; }

/// Collective (every rank calls it with the same `bytes`, which all ranks derive from the replicated topology): make the
/// staging buffers at least `bytes` each and map every peer's allocation. The caller guarantees that no push is in flight
/// (run_apply flushes the lagging exchange first). Falls back to the NCCL all-gather path if CUDA IPC is unavailable.
void comm_stage_reserve(mrx_comm *c, size_t bytes, cudaStream_t st) {
    if (bytes <= c->stageBytes) return;
    NcclApi &a = api();
    const int W = c->world;
    MRX_CUDA(cudaStreamSynchronize(st));
    if (c->pushStream) MRX_CUDA(cudaStreamSynchronize(c->pushStream));
    // nobody may still be writing into the buffers that are about to be unmapped
    if (!c->ipcScratch) MRX_CUDA(cudaMalloc(&c->ipcScratch, (size_t)W * sizeof(cudaIpcMemHandle_t) + 64));
    MRX_CUDA(cudaMemsetAsync(c->ipcScratch, 0, sizeof(double), st));
    check(a.AllReduce(c->ipcScratch, c->ipcScratch, 1, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce(barrier)");
    MRX_CUDA(cudaStreamSynchronize(st));
    if (c->stageBase) {
        for (int r = 0; r < W; r++)
            if (r != c->rank && c->peerBase[r]) cudaIpcCloseMemHandle(c->peerBase[r]);
    }
    // second barrier: every rank has unmapped the old peers' buffers BEFORE anybody frees its own (freeing memory that an
    // importing process still has mapped is undefined)
    check(a.AllReduce(c->ipcScratch, c->ipcScratch, 1, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce(barrier)");
    MRX_CUDA(cudaStreamSynchronize(st));
    if (c->stageBase) {
        MRX_CUDA(cudaFree(c->stageBase));
        c->stageBase = nullptr;
    }
    size_t nb = std::max(bytes + bytes / 2, (size_t)64 << 20);
    nb = (nb + 4095) / 4096 * 4096;
    MRX_CUDA(cudaMalloc(&c->stageBase, nb * mrx_comm::kStageBufs));
    c->stageBytes = nb;
    c->peerBase.assign(W, nullptr);
    c->peerBase[c->rank] = c->stageBase;
    if (!c->pushStream) {
        MRX_CUDA(cudaStreamCreateWithFlags(&c->pushStream, cudaStreamNonBlocking));
        for (int b = 0; b < mrx_comm::kStageBufs; b++) {
            MRX_CUDA(cudaEventCreateWithFlags(&c->evReduced[b], cudaEventDisableTiming));
            MRX_CUDA(cudaEventCreateWithFlags(&c->evPushed[b], cudaEventDisableTiming));
        }
    }
    const bool forceOff = getenv("MRX_NO_IPC") != nullptr;
    int ok = forceOff ? 0 : 1;
    cudaIpcMemHandle_t mine;
    std::memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, c->stageBase) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    // handles of all ranks (64 bytes each) through an in-place all-gather on device memory
    std::vector<cudaIpcMemHandle_t> all(W);
    MRX_CUDA(cudaMemcpyAsync(c->ipcScratch + (size_t)c->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    check(a.AllGather(c->ipcScratch + (size_t)c->rank * sizeof(mine), c->ipcScratch, sizeof(mine), kNcclInt8, c->comm, st), "ncclAllGather(ipc)");
    MRX_CUDA(cudaMemcpyAsync(all.data(), c->ipcScratch, sizeof(mine) * W, cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < W && ok; r++) {
        if (r == c->rank) continue;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        c->peerBase[r] = static_cast<char *>(p);
    }
    // the exchange mode must be the same everywhere: min over ranks
    double flag = ok ? 1.0 : 0.0;
    double *dflag = reinterpret_cast<double *>(c->ipcScratch + (size_t)W * sizeof(mine));
    dflag = reinterpret_cast<double *>(((uintptr_t)dflag + 7) & ~(uintptr_t)7);
    MRX_CUDA(cudaMemcpyAsync(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
    check(a.AllReduce(dflag, dflag, 1, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce(ipc ok)");
    MRX_CUDA(cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    c->ipcOk = (flag > W - 0.5);
    if (!c->ipcOk && !c->ipcTried && c->rank == 0 && !forceOff)
        std::fprintf(stderr, "[mrx] CUDA IPC peer mapping unavailable: coefficient exchange falls back to ncclAllGather\n");
    c->ipcTried = true;
}

/// push `bytes` at offset `off` of staging buffer `buf` into the same place of every peer's buffer (copy engines, push stream);
/// starts after evReduced[buf] (recorded by the caller on its compute stream), completion = evPushed[buf]
void comm_push(mrx_comm *c, int buf, size_t off, size_t bytes) {
    MRX_CUDA(cudaStreamWaitEvent(c->pushStream, c->evReduced[buf], 0));
    if (bytes > 0) {
        const char *src = c->stageBase + (size_t)buf * c->stageBytes + off;
        for (int d = 1; d < c->world; d++) {
            const int r = (c->rank + d) % c->world; // staggered: at any moment the ranks target distinct peers
            MRX_CUDA(cudaMemcpyAsync(c->peerBase[r] + (size_t)buf * c->stageBytes + off, src, bytes, cudaMemcpyDeviceToDevice, c->pushStream));
        }
    }
    MRX_CUDA(cudaEventRecord(c->evPushed[buf], c->pushStream));
}
cudaStream_t comm_unpack_stream(mrx_comm *c) {
    if (!c->unpackStream) {
        MRX_CUDA(cudaStreamCreateWithFlags(&c->unpackStream, cudaStreamNonBlocking));
        MRX_CUDA(cudaEventCreateWithFlags(&c->evGathered, cudaEventDisableTiming));
        for (int b = 0; b < mrx_comm::kStageBufs; b++) MRX_CUDA(cudaEventCreateWithFlags(&c->evUnpacked[b], cudaEventDisableTiming));
    }
    return c->unpackStream;
}
cudaEvent_t comm_ev_gathered(const mrx_comm *c) { return c->evGathered; }
cudaEvent_t comm_ev_unpacked(const mrx_comm *c, int buf) { return c->evUnpacked[buf]; }
cudaEvent_t comm_ev_reduced(const mrx_comm *c, int buf) { return c->evReduced[buf]; }
cudaEvent_t comm_ev_pushed(const mrx_comm *c, int buf) { return c->evPushed[buf]; }

void comm_allreduce_sum(const mrx_comm *c, double *buf, size_t n, cudaStream_t st) {
    check(api().AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce");
}

// ---- host arena shared by the ranks of one node -------------------------------------------------------------------------
// One anonymous shared-memory file (memfd: not bounded by the size of /dev/shm), created by rank 0 and mapped by every
// rank through /proc/<pid>/fd/<fd>, then registered with CUDA in every process: each GPU can DMA into the SAME host pages
// over its own PCIe link. Host chunks of an output tree with a shared mirror are carved from it in allocation order, which
// is the same on every rank (SPMD), so chunk i of the tree has the same offset everywhere.
namespace {
struct HostArena {
    char *base = nullptr;
    size_t bytes = 0, used = 0;
    int live = 0, fd = -1;
} g_arena;
} // namespace

void *host_arena_alloc(size_t bytes) {
    bytes = (bytes + 4095) & ~(size_t)4095;
    if (!g_arena.base || g_arena.used + bytes > g_arena.bytes)
        MRX_ABORT("shared host arena exhausted (mrx_comm_host_arena: ask for more bytes)");
    void *p = g_arena.base + g_arena.used;
    g_arena.used += bytes;
    g_arena.live++;
    return p;
}
void host_arena_free(void *) {
    if (--g_arena.live == 0) g_arena.used = 0; // bump allocator: space comes back when the last chunk is gone
}
bool comm_has_host_arena(const mrx_comm *c) { return c && c->hostArena && g_arena.base; }
long long host_arena_offset(const void *p) { return (long long)(static_cast<const char *>(p) - g_arena.base); }

int comm_host_arena(mrx_comm *c, size_t bytes) {
    NcclApi &a = api();
    cudaStream_t st = stream();
    bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    if (g_arena.base) {
        if (g_arena.bytes >= bytes || g_arena.live > 0) {
            c->hostArena = g_arena.bytes >= bytes;
            return c->hostArena ? 0 : 1;
        }
        cudaHostUnregister(g_arena.base);
        munmap(g_arena.base, g_arena.bytes);
        if (g_arena.fd >= 0) close(g_arena.fd);
        g_arena = HostArena{};
    }
    if (!c->ipcScratch) MRX_CUDA(cudaMalloc(&c->ipcScratch, (size_t)c->world * sizeof(cudaIpcMemHandle_t) + 64));
    long long info[2] = {0, -1};
    int ok = 1;
    if (c->rank == 0) {
        g_arena.fd = memfd_create("mrx_host_arena", 0);
        if (g_arena.fd < 0 || ftruncate(g_arena.fd, (off_t)bytes) != 0) ok = 0;
        info[0] = (long long)getpid();
        info[1] = ok ? g_arena.fd : -1;
    }
    long long *dinfo = reinterpret_cast<long long *>(c->ipcScratch);
    if (c->rank == 0) MRX_CUDA(cudaMemcpyAsync(dinfo, info, sizeof(info), cudaMemcpyHostToDevice, st));
    check(a.Broadcast(dinfo, dinfo, sizeof(info), kNcclInt8, 0, c->comm, st), "ncclBroadcast(arena)");
    MRX_CUDA(cudaMemcpyAsync(info, dinfo, sizeof(info), cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    if (info[1] < 0) ok = 0;
    if (ok && c->rank != 0) {
        char path[64];
        std::snprintf(path, sizeof(path), "/proc/%lld/fd/%lld", info[0], info[1]);
        g_arena.fd = open(path, O_RDWR);
        if (g_arena.fd < 0) ok = 0;
    }
    void *p = MAP_FAILED;
    if (ok) {
        p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, g_arena.fd, 0);
        if (p == MAP_FAILED) ok = 0;
    }
    if (ok) {
        int dev = 0, same = 0;
        MRX_CUDA(cudaGetDevice(&dev));
        cudaDeviceGetAttribute(&same, cudaDevAttrCanUseHostPointerForRegisteredMem, dev);
        if (!same || cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
            cudaGetLastError();
            munmap(p, bytes);
            ok = 0;
        }
    }
    // all ranks or none
    double flag = ok ? 1.0 : 0.0;
    double *dflag = reinterpret_cast<double *>(c->ipcScratch + 32);
    MRX_CUDA(cudaMemcpyAsync(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice, st));
    check(a.AllReduce(dflag, dflag, 1, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce(arena ok)");
    MRX_CUDA(cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    const bool all = flag > c->world - 0.5;
    if (ok && !all) {
        cudaHostUnregister(p);
        munmap(p, bytes);
    }
    if (!all) {
        if (g_arena.fd >= 0) close(g_arena.fd);
        g_arena = HostArena{};
        if (c->rank == 0) std::fprintf(stderr, "[mrx] shared host arena unavailable: the host mirror of a sharded apply stays on rank 0's PCIe link\n");
        return 1;
    }
    g_arena.base = static_cast<char *>(p);
    g_arena.bytes = bytes;
    c->hostArena = true;
    return 0;
}

} // namespace mrx

extern "C" {

int mrx_comm_unique_id(char *id128) {
    mrx::NcclUniqueId id;
    mrx::check(mrx::api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, id.internal, 128);
    return 0;
}

mrx_comm *mrx_comm_create(int rank, int world, const char *id128) {
    mrx::require_device("mrx_comm_create");
    auto *c = new mrx_comm;
    c->rank = rank;
    c->world = world;
    mrx::NcclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    mrx::check(mrx::api().CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank");
    return c;
}

void mrx_comm_destroy(mrx_comm *c) {
    if (!c) return;
    if (c->stageBase) {
        cudaDeviceSynchronize();
        for (int r = 0; r < c->world; r++)
            if (r != c->rank && r < (int)c->peerBase.size() && c->peerBase[r]) cudaIpcCloseMemHandle(c->peerBase[r]);
        // destruction is not collective: a peer may still have this buffer mapped, and freeing exported memory before every
        // importer has closed it is undefined -> an exported buffer is left to process teardown
        if (!c->ipcOk) cudaFree(c->stageBase);
    }
    if (c->ipcScratch) cudaFree(c->ipcScratch);
    if (c->pushStream) cudaStreamDestroy(c->pushStream);
    if (c->unpackStream) {
        cudaStreamDestroy(c->unpackStream);
        cudaEventDestroy(c->evGathered);
        for (int b = 0; b < mrx_comm::kStageBufs; b++) cudaEventDestroy(c->evUnpacked[b]);
    }
    for (int b = 0; b < mrx_comm::kStageBufs; b++) {
        if (c->evReduced[b]) cudaEventDestroy(c->evReduced[b]);
        if (c->evPushed[b]) cudaEventDestroy(c->evPushed[b]);
    }
    if (c->comm) mrx::api().CommDestroy(c->comm);
    delete c;
}

int mrx_comm_host_arena(mrx_comm *c, long long bytes) {
    mrx::require_device("mrx_comm_host_arena");
    if (!c || c->world < 2 || bytes <= 0) return 1;
    return mrx::comm_host_arena(c, (size_t)bytes);
}
int mrx_comm_rank(const mrx_comm *c) { return mrx::comm_rank(c); }
int mrx_comm_size(const mrx_comm *c) { return mrx::comm_world(c); }

/* cyclic distribution of one refinement iteration's work vector (n items) over `world` ranks, the layout the sharded
 * apply uses: rank r computes the items i = r, r + world, ...; *count = how many those are, *rows = items per rank after
 * padding to equal segments ((n + world - 1) / world). Item i lives in row (i % world) * rows + i / world of the rank-major
 * exchange buffers (norms and coefficient blocks). Pure host logic, shared with the CPU tests. */
void mrx_shard_cyclic(int n, int world, int rank, int *count, int *rows) {
    const int B = mrx::shard_block();
    if (rows) *rows = mrx::shard_rows(n, world, B);
    if (count) *count = mrx::shard_count(n, world, rank, B);
}
int mrx_shard_cyclic_row(int i, int n, int world) {
    const int B = mrx::shard_block();
    return mrx::shard_row(i, world, mrx::shard_rows(n, world, B), B);
}
int mrx_shard_block(void) { return mrx::shard_block(); }

/* contiguous, order-preserving partition of n weighted items into `world` ranges: begin[r]..begin[r+1]
 * (the split of one refinement iteration's work vector; pure host logic, also used by the CPU tests) */
void mrx_shard_partition(const long long *cost, int n, int world, int *begin) {
    long long total = 0;
    for (int i = 0; i < n; i++) total += cost[i] > 0 ? cost[i] : 0;
    begin[0] = 0;
    int i = 0;
    long long acc = 0;
    for (int r = 1; r < world; r++) {
        // boundary r: first index where the running cost reaches r/world of the total
        const long double want = (long double)total * r / world;
        while (i < n && (long double)acc + (cost[i] > 0 ? cost[i] : 0) * 0.5L < want) {
            acc += cost[i] > 0 ? cost[i] : 0;
            i++;
        }
        begin[r] = i;
    }
    begin[world] = n;
}

} // extern "C"
