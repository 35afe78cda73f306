// NCCL plumbing for the sharded apply: one process per GPU, the output-node list of every
// refinement iteration is split across ranks, norms and output coefficient blocks are exchanged over
// NVLink. NCCL is loaded at run time (dlopen of libnccl.so.2: the copy the host framework already loaded,
// else the system one), so the library has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../engine.hpp"
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
};

namespace mrx {

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

void comm_allreduce_sum(const mrx_comm *c, double *buf, size_t n, cudaStream_t st) {
    check(api().AllReduce(buf, buf, n, kNcclFloat64, kNcclSum, c->comm, st), "ncclAllReduce");
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
    if (c->comm) mrx::api().CommDestroy(c->comm);
    delete c;
}

int mrx_comm_rank(const mrx_comm *c) { return mrx::comm_rank(c); }
int mrx_comm_size(const mrx_comm *c) { return mrx::comm_world(c); }

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
