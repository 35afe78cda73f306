// Band enumeration on the device: ConvolutionCalculator::makeOperBand / fillOperBand
// (src/treebuilders/ConvolutionCalculator.cpp:142-222, non-periodic branch) and the lazy generation of
// finer input nodes it triggers (MWTree::getNode -> MWNode::retrieveNode -> genChildren,
// src/trees/MWTree.cpp:340-352, MWNode.cpp:1097-1120, FunctionNode.cpp:293-328), restated as bulk kernels.
//
// The reference walks, for every output node, the cube of (2 band_max + 1)^3 translations, fetching (and
// generating, under locks) the input node at the output node's scale for each. Here:
//   probe    lane per (output node, offset of its depth that at least one (term, gt, ft) can reach (band tables)):
//            clips to the world, descends the input tree's child pointers to the deepest existing node on
//            the way to (scale, l), and drops offsets whose largest operator-norm product times an upper
//            bound of |f| cannot pass the norm screening (the screening itself is re-done exactly in
//            pipe_screen; this is only a conservative early-out).
//   scan     offsets of the survivors in the neighbour list (node order, offset order) and of their candidates
//   emit     survivors are written as neighbour entries; entries whose node is still coarser than the
//            target scale are queued as pending.
//   resolve  thread per pending entry: continue the descent; a childless node on the way is flagged (once).
//   create   thread per flagged node: 8 generated children get slots, depth and norm bound; the (parent,
//            child0) item list feeds transform_kernel<2> which fills their scaling coefficients.
// resolve/create repeat until no entry is pending (one round per missing level).
#include <algorithm>

#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

__device__ __forceinline__ int descend(const int *__restrict__ child0, int node, int &depth, int target, int lx, int ly, int lz) {
    while (depth < target) {
        const int c0 = child0[node];
        if (c0 < 0) break;
        const int shift = target - depth - 1;
        node = c0 + (((lx >> shift) & 1) | (((ly >> shift) & 1) << 1) | (((lz >> shift) & 1) << 2));
        depth++;
    }
    return node;
}

// warp -> (output node j, chunk c of 32 offsets): chunkOff is the prefix of ceil(offCount / 32) over the nodes
__device__ __forceinline__ int find_node(const int *__restrict__ chunkOff, int nG, int w) {
    int lo = 0, hi = nG; // invariant: chunkOff[lo] <= w < chunkOff[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunkOff[mid] <= w) lo = mid;
        else hi = mid;
    }
    return lo;
}

constexpr int kPendingBit = 1 << 30;

// periodic::index_manipulation for scale >= 0 (periodic_utils.cpp:49-73): translation wrapped into [-2^n, 2^n)
__device__ __forceinline__ int wrap_cell(int l, int scale) {
    const int two_n = 1 << (scale + 1);
    int t = l + (two_n >> 1);
    t &= two_n - 1; // two_n is a power of two: the non-negative remainder
    return t - (two_n >> 1);
}

// probe: one lane per (output node, reachable offset). Every pointer chase of the iteration is in flight at once (the
// earlier warp-per-node loop serialised ~offCount / 32 dependent descents per node and was pure latency).
__global__ void __launch_bounds__(256) enum_probe_kernel(EnumParams E) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= E.nChunks) return;
    const int j = find_node(E.chunkOff, E.nG, w);
    const int4 gn = E.gNodes[j];
    const int dep = gn.x;
    const int o0 = E.offStart[dep], o1 = o0 + E.offCount[dep];
    const int oi = o0 + 32 * (w - E.chunkOff[j]) + lane;
    const int td = dep + E.depthShift; // depth in the function tree
    bool hit = false;
    int node = 0, nd = 0, nc = 0;
    if (oi < o1) {
        const OffEntry oe = E.offs[oi];
        int lx = gn.y + oe.dx, ly = gn.z + oe.dy, lz = gn.w + oe.dz;
        int lo[3], hi[3];
#pragma unroll
        for (int x = 0; x < 3; x++) {
            lo[x] = E.corner[x] * (1 << td);
            hi[x] = lo[x] + E.nboxes[x] * (1 << td) - 1;
            if (E.periodic) { // `reach` cells around the world instead of the world (ConvolutionCalculator.cpp:169-171)
                hi[x] = (lo[x] + E.nboxes[x] * (1 << td)) * E.reach - 1;
                lo[x] = lo[x] * E.reach;
            }
        }
        const double slack = 1.0 + 1e-9;
        // apply with precTrees: the node's own threshold (ConvolutionCalculator.cpp:244-245)
        const double gThrs = E.precFac ? E.prec * E.precFac[j] * E.sqrtTerm : E.gThrs;
        bool inb = lx >= lo[0] && lx <= hi[0] && ly >= lo[1] && ly <= hi[1] && lz >= lo[2] && lz <= hi[2];
        if (E.periodic) {
            const int half = 1 << td; // unit cell [-2^n, 2^n) (periodic::in_unit_cell, periodic_utils.cpp:35-47)
            const bool inCell = lx >= -half && lx < half && ly >= -half && ly < half && lz >= -half && lz < half;
            if (E.unitCell == 1 && !inCell) inb = false;
            if (E.unitCell == 2 && inCell) inb = false;
            lx = wrap_cell(lx, td); // MWTree::getNode looks the wrapped index up (MWTree.cpp:341)
            ly = wrap_cell(ly, td);
            lz = wrap_cell(lz, td);
        }
        if (inb && (!E.screenOn || oe.maxO * E.fMaxNorm * slack > gThrs)) {
            node = ((lx >> td) - E.corner[0]) + E.nboxes[0] * (((ly >> td) - E.corner[1]) + E.nboxes[1] * ((lz >> td) - E.corner[2]));
            node = descend(E.fChild0, node, nd, td, lx, ly, lz);
            // |f_ft| <= |node| for a real node; a generated node is an orthogonal projection of its real leaf
            // ancestor, so the ancestor's norm bounds it
            if (!E.screenOn || !(oe.maxO * E.fBound[node] * slack <= gThrs)) {
                hit = true;
                const int *coff = E.candOff + E.depthInfo[dep].cubeOff;
                nc = coff[oe.code + 1] - coff[oe.code];
            }
        }
    }
    E.pNode[(size_t)w * 32 + lane] = hit ? (node | (nd < td ? kPendingBit : 0)) : -1;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    int sum = nc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    // low 32 bits: neighbours, high 32 bits: candidates (both < 2^31 per iteration, checked by the host)
    if (lane == 0) E.chunkPacked[w] = (unsigned long long)__popc(bal) | ((unsigned long long)(unsigned)sum << 32);
}

// exclusive scans of the per-chunk neighbour and candidate counts (chunk order = node order, offset order: the neighbour
// list is deterministic). Two levels: CTAs of kScanTile chunks scan locally (coalesced), one CTA scans the tile totals.
constexpr int kScanTile = 2048; // 256 threads x 8 chunks

__device__ __forceinline__ unsigned long long cta_excl_scan(unsigned long long v, unsigned long long *sm, unsigned long long &total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long s = (lane < nw) ? sm[lane] : 0ull;
        unsigned long long si = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, si, off);
            if (lane >= off) si += t;
        }
        sm[lane] = si - s;
        if (lane == 31) sm[32] = si;
    }
    __syncthreads();
    const unsigned long long res = sm[warp] + incl - v;
    total = sm[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(256) enum_scan_tiles_kernel(EnumParams E) {
    __shared__ unsigned long long sm[33];
    const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
    unsigned long long v[8], local = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = (base + i < E.nChunks) ? E.chunkPacked[base + i] : 0ull;
        local += v[i];
    }
    unsigned long long total;
    unsigned long long run = cta_excl_scan(local, sm, total);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (base + i < E.nChunks) E.chunkScan[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0) E.tileTotal[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) enum_scan_top_kernel(EnumParams E, int nTiles) {
    __shared__ unsigned long long sm[33];
    const int tid = threadIdx.x;
    const int per = (nTiles + 1023) / 1024;
    const int b0 = min(tid * per, nTiles), b1 = min(b0 + per, nTiles);
    unsigned long long local = 0;
    for (int b = b0; b < b1; b++) local += E.tileTotal[b];
    unsigned long long total;
    unsigned long long run = cta_excl_scan(local, sm, total);
    for (int b = b0; b < b1; b++) {
        const unsigned long long t = E.tileTotal[b];
        E.tileBase[b] = run;
        run += t;
    }
    if (tid == 0) {
        E.tileBase[nTiles] = total;
        E.cnt->nNbr = (int)(unsigned)total;
        E.cnt->nCand = (unsigned)(total >> 32);
    }
}

__device__ __forceinline__ unsigned long long chunk_offset(const EnumParams &E, int c) {
    return (c < E.nChunks) ? E.chunkScan[c] + E.tileBase[c / kScanTile] : E.tileBase[(E.nChunks + kScanTile - 1) / kScanTile];
}

// thread per output node: its slice of the neighbour list
__global__ void __launch_bounds__(256) enum_desc_kernel(EnumParams E) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= E.nG) return;
    GDesc d;
    d.slot = E.gSlots[j];
    d.depth = E.gNodes[j].x;
    const int a = (int)(unsigned)chunk_offset(E, E.chunkOff[j]);
    const int b = (int)(unsigned)chunk_offset(E, E.chunkOff[j + 1]);
    d.nbrOff = a;
    d.nbrCnt = b - a;
    d.partial = -1;
    E.gdesc[j] = d;
}

__global__ void __launch_bounds__(256) enum_emit_kernel(EnumParams E) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= E.nChunks) return;
    const unsigned long long mine = E.chunkPacked[w];
    if ((unsigned)mine == 0u) return;
    const unsigned long long base = chunk_offset(E, w);
    const int j = find_node(E.chunkOff, E.nG, w);
    const int dep = E.gNodes[j].x;
    const int oi = E.offStart[dep] + 32 * (w - E.chunkOff[j]) + lane;
    const int pn = E.pNode[(size_t)w * 32 + lane];
    const bool hit = pn >= 0;
    int code = 0, nc = 0;
    if (hit) {
        code = E.offs[oi].code;
        const int *coff = E.candOff + E.depthInfo[dep].cubeOff;
        nc = coff[code + 1] - coff[code];
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    int incl = nc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (hit) {
        const int pos = (int)(unsigned)base + __popc(bal & ((1u << lane) - 1u));
        NbrEntry e;
        e.fslot = pn & ~kPendingBit;
        e.code = code;
        e.g = j;
        e.candBase = (int)((unsigned)(base >> 32) + (unsigned)(incl - nc));
        E.nbr[pos] = e;
        if (pn & kPendingBit) E.pending[atomicAdd(&E.cnt->nPending, 1)] = pos;
    }
}

__global__ void __launch_bounds__(256) resolve_kernel(EnumParams E, int nPending) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPending) return;
    const int idx = E.pending[p];
    if (idx < 0) return; // resolved in an earlier round
    NbrEntry e = E.nbr[idx];
    const int4 gn = E.gNodes[e.g];
    const int W = E.depthInfo[gn.x].W, cube = 2 * W + 1;
    int lx = gn.y + e.code % cube - W, ly = gn.z + (e.code / cube) % cube - W, lz = gn.w + e.code / (cube * cube) - W;
    const int td = gn.x + E.depthShift;
    if (E.periodic) {
        lx = wrap_cell(lx, td);
        ly = wrap_cell(ly, td);
        lz = wrap_cell(lz, td);
    }
    int nd = E.fDepth[e.fslot];
    const int node = descend(E.fChild0, e.fslot, nd, td, lx, ly, lz);
    E.nbr[idx].fslot = node;
    if (nd == td) {
        E.pending[p] = -1;
        return;
    }
    atomicAdd(&E.cnt->nUnresolved, 1);
    if (atomicCAS(&E.fFlag[node], 0, 1) == 0) E.newParents[atomicAdd(&E.cnt->nNewParents, 1)] = node;
}

__global__ void __launch_bounds__(256) create_kernel(EnumParams E, int nNew, int firstSlot) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nNew) return;
    const int p = E.newParents[r];
    const int c0 = firstSlot + 8 * r;
    E.fChild0[p] = c0;
    E.fFlag[p] = 0;
    const int dp = E.fDepth[p] + 1;
    const double b = E.fBound[p];
    for (int c = 0; c < 8; c++) {
        E.fChild0[c0 + c] = -1;
        E.fDepth[c0 + c] = dp;
        E.fBound[c0 + c] = b;
        E.fFlag[c0 + c] = 0;
    }
    E.genItems[2 * r] = p;
    E.genItems[2 * r + 1] = c0;
}

// ---- lazy residency of the input tree (engine.hpp DeviceTree::partial): the (node, block) pairs the apply is about to read
//      are gathered from the pinned host chunks by the device itself; only they cross PCIe. resident[8 node + block].
__global__ void __launch_bounds__(256) fetch_mark_kernel(const int *__restrict__ list, int n, int nRealF, int *__restrict__ resident,
                                                         int *__restrict__ fetchList, int *__restrict__ fetchCnt) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 8 * n) return;
    const int node = list[i >> 3]; // generation reads all eight blocks of a real leaf
    if (node < nRealF) {
        const int blk = node * 8 + (i & 7);
        if (atomicCAS(&resident[blk], 0, 1) == 0) fetchList[atomicAdd(fetchCnt, 1)] = blk;
    }
}

// resident[blk]: 0 not in HBM, 1 queued (pipe_fill / fetch_mark), 2 arrived (diagnostic: consumers are ordered behind the gather by
// stream events). Entries [*begin, *end) of the queue: the gather of one node sub-range of an iteration runs on its own stream
// beside the contraction of the sub-range before it (apply.cu), with one CTA per SM (30 registers: fits next to the three
// persistent contraction CTAs of an SM; measured beside the contraction: 51 GB/s, profiles/r02w).
__global__ void __launch_bounds__(256) fetch_nodes_kernel(double *__restrict__ coefs, const double *const *__restrict__ chunkTab,
                                                          const int *__restrict__ list, const int *__restrict__ begin,
                                                          const int *__restrict__ end, int ncoef, unsigned long long *__restrict__ total,
                                                          int *__restrict__ resident) {
    const int i0 = begin ? *begin : 0, n = *end;
    const int Kd = ncoef / 8;
    for (int i = i0 + blockIdx.x; i < n; i += gridDim.x) {
        const int blk = list[i], slot = blk >> 3, c = blk & 7;
        const double *src = chunkTab[slot >> 6] + (size_t)(slot & 63) * ncoef + (size_t)c * Kd;
        double *dst = coefs + (size_t)slot * ncoef + (size_t)c * Kd;
        if ((Kd & 1) == 0) {
            const double2 *s2 = reinterpret_cast<const double2 *>(src);
            double2 *d2 = reinterpret_cast<double2 *>(dst);
            for (int e = threadIdx.x; e < Kd / 2; e += 256) d2[e] = s2[e];
        } else {
            for (int e = threadIdx.x; e < Kd; e += 256) dst[e] = src[e];
        }
        if (threadIdx.x == 0) resident[blk] = 2;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *total += (unsigned long long)(n - i0); // blocks fetched
}

// host mirror of an apply output (mrx_tree_set_host_mirror): what the streamed per-iteration copies could not take yet -- the
// scaling block of every node (final after the closing TopDown) and all blocks of the branch nodes (rewritten by BottomUp) --
// written by the SMs straight into the pinned host chunks over PCIe
__global__ void __launch_bounds__(256) push_nodes_kernel(const double *__restrict__ coefs, double *const *__restrict__ chunkTab,
                                                         const int *__restrict__ items, int n, int ncoef) {
    const int Kd = ncoef / 8;
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const int it = items[i];
        const int slot = it & 0x7fffffff;
        const int cnt = (it < 0) ? ncoef : Kd;
        const double *src = coefs + (size_t)slot * ncoef;
        double *dst = chunkTab[slot >> 6] + (size_t)(slot & 63) * ncoef;
        if ((cnt & 1) == 0) {
            const double2 *s2 = reinterpret_cast<const double2 *>(src);
            double2 *d2 = reinterpret_cast<double2 *>(dst);
            for (int e = threadIdx.x; e < cnt / 2; e += 256) d2[e] = s2[e];
        } else {
            for (int e = threadIdx.x; e < cnt; e += 256) dst[e] = src[e];
        }
    }
}

} // namespace

void launch_push_nodes(const double *coefs, double *const *chunkTab, const int *items, int n, int ncoef, cudaStream_t st) {
    if (n <= 0) return;
    push_nodes_kernel<<<std::min(n, 1184), 256, 0, st>>>(coefs, chunkTab, items, n, ncoef);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_fetch_mark(const int *list, int n, int nRealF, int *resident, int *fetchList, int *fetchCnt, cudaStream_t st) {
    if (n <= 0) return;
    fetch_mark_kernel<<<(8 * n + 255) / 256, 256, 0, st>>>(list, n, nRealF, resident, fetchList, fetchCnt);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_fetch_nodes(double *coefs, const double *const *chunkTab, const int *list, const int *begin, const int *end, int ncoef,
                        unsigned long long *total, int *resident, cudaStream_t st, int grid) {
    fetch_nodes_kernel<<<grid > 0 ? grid : 1184, 256, 0, st>>>(coefs, chunkTab, list, begin, end, ncoef, total, resident);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_enum(const EnumParams &E, cudaStream_t st) {
    if (E.nG <= 0) return;
    const int nTiles = (E.nChunks + kScanTile - 1) / kScanTile;
    if (E.nChunks > 0) {
        enum_probe_kernel<<<(E.nChunks + 7) / 8, 256, 0, st>>>(E);
        enum_scan_tiles_kernel<<<nTiles, 256, 0, st>>>(E);
        launch_counter() += 2;
    }
    enum_scan_top_kernel<<<1, 1024, 0, st>>>(E, nTiles);
    enum_desc_kernel<<<(E.nG + 255) / 256, 256, 0, st>>>(E);
    launch_counter() += 2;
    if (E.nChunks > 0) {
        enum_emit_kernel<<<(E.nChunks + 7) / 8, 256, 0, st>>>(E);
        launch_counter()++;
    }
    MRX_CUDA(cudaGetLastError());
}

void launch_enum_resolve(const EnumParams &E, int nPending, cudaStream_t st) {
    if (nPending <= 0) return;
    resolve_kernel<<<(nPending + 255) / 256, 256, 0, st>>>(E, nPending);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_enum_create(const EnumParams &E, int nNew, int firstSlot, cudaStream_t st) {
    if (nNew <= 0) return;
    create_kernel<<<(nNew + 255) / 256, 256, 0, st>>>(E, nNew, firstSlot);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
