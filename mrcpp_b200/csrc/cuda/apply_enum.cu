// Band enumeration on the device: ConvolutionCalculator::makeOperBand / fillOperBand
// (src/treebuilders/ConvolutionCalculator.cpp:142-222, non-periodic branch) and the lazy generation of
// finer input nodes it triggers (MWTree::getNode -> MWNode::retrieveNode -> genChildren,
// src/trees/MWTree.cpp:340-352, MWNode.cpp:1097-1120, FunctionNode.cpp:293-328), restated as bulk kernels.
//
// The reference walks, for every output node, the cube of (2 band_max + 1)^3 translations, fetching (and
// generating, under locks) the input node at the output node's scale for each. Here:
//   enum     warp per output node; lanes over the offsets of its depth that at least one (term, gt, ft) can
//            reach (band tables). Each lane clips to the world, descends the input tree's child pointers
//            to the deepest existing node on the way to (scale, l), and drops offsets whose largest
//            operator-norm product times an upper bound of |f| cannot pass the norm screening (the screening
//            itself is re-done exactly in pipe_screen; this is only a conservative early-out). Survivors
//            are written in offset order as neighbour entries; entries whose node is still coarser than
//            the target scale are queued as pending.
//   resolve  thread per pending entry: continue the descent; a childless node on the way is flagged (once).
//   create   thread per flagged node: 8 generated children get slots, depth and norm bound; the (parent,
//            child0) item list feeds transform_kernel<2> which fills their scaling coefficients.
// resolve/create repeat until no entry is pending (one round per missing level).
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

__global__ void __launch_bounds__(256) enum_kernel(EnumParams E) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= E.nG) return;
    const int4 gn = E.gNodes[i];
    const int dep = gn.x;
    GDesc d;
    d.slot = E.gSlots[i];
    d.depth = dep;
    d.nbrOff = 0;
    d.nbrCnt = 0;
    d.partial = -1;
    if (dep < 0 || dep >= E.DM || E.depthInfo[dep].W < 0) { // deeper than every operator tree: empty band (:146-151)
        if (lane == 0) E.gdesc[i] = d;
        return;
    }
    const int o0 = E.offStart[dep], o1 = o0 + E.offCount[dep];
    const int td = dep + E.depthShift; // depth in the function tree
    int lo[3], hi[3];
#pragma unroll
    for (int x = 0; x < 3; x++) {
        lo[x] = E.corner[x] * (1 << td);
        hi[x] = lo[x] + E.nboxes[x] * (1 << td) - 1;
    }
    const int *coff = E.candOff + E.depthInfo[dep].cubeOff;
    const double slack = 1.0 + 1e-9;

    int nbrBase = 0;
    unsigned candBase = 0;
    int total = 0;
    for (int pass = 0; pass < 2; pass++) {
        int run = 0;
        unsigned candRun = 0;
        for (int base = o0; base < o1; base += 32) {
            const int oi = base + lane;
            bool hit = false;
            int node = 0, nd = 0, code = 0, nc = 0;
            if (oi < o1) {
                const OffEntry oe = E.offs[oi];
                const int lx = gn.y + oe.dx, ly = gn.z + oe.dy, lz = gn.w + oe.dz;
                const bool inb = lx >= lo[0] && lx <= hi[0] && ly >= lo[1] && ly <= hi[1] && lz >= lo[2] && lz <= hi[2];
                if (inb && (!E.screenOn || oe.maxO * E.fMaxNorm * slack > E.gThrs)) {
                    node = ((lx >> td) - E.corner[0]) + E.nboxes[0] * (((ly >> td) - E.corner[1]) + E.nboxes[1] * ((lz >> td) - E.corner[2]));
                    node = descend(E.fChild0, node, nd, td, lx, ly, lz);
                    // |f_ft| <= |node| for a real node; a generated node is an orthogonal projection of its real leaf
                    // ancestor, so the ancestor's norm bounds it
                    if (!E.screenOn || !(oe.maxO * E.fBound[node] * slack <= E.gThrs)) {
                        hit = true;
                        code = oe.code;
                        nc = coff[code + 1] - coff[code];
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            // exclusive scan of the candidate counts of the hit lanes
            int incl = hit ? nc : 0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            if (pass == 1 && hit) {
                const int pos = nbrBase + run + __popc(bal & ((1u << lane) - 1u));
                NbrEntry e;
                e.fslot = node;
                e.code = code;
                e.g = i;
                e.candBase = (int)(candBase + candRun + (unsigned)(incl - nc));
                E.nbr[pos] = e;
                if (nd < td) E.pending[atomicAdd(&E.cnt->nPending, 1)] = pos;
            }
            run += __popc(bal);
            candRun += (unsigned)__shfl_sync(0xffffffffu, incl, 31);
        }
        if (pass == 0) {
            total = run;
            if (lane == 0) {
                nbrBase = atomicAdd(&E.cnt->nNbr, run);
                candBase = atomicAdd(&E.cnt->nCand, candRun);
            }
            nbrBase = __shfl_sync(0xffffffffu, nbrBase, 0);
            candBase = __shfl_sync(0xffffffffu, candBase, 0);
            if (total == 0) break;
        }
    }
    d.nbrOff = nbrBase;
    d.nbrCnt = total;
    if (lane == 0) E.gdesc[i] = d;
}

__global__ void __launch_bounds__(256) resolve_kernel(EnumParams E, int nPending) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPending) return;
    const int idx = E.pending[p];
    if (idx < 0) return; // resolved in an earlier round
    NbrEntry e = E.nbr[idx];
    const int4 gn = E.gNodes[e.g];
    const int W = E.depthInfo[gn.x].W, cube = 2 * W + 1;
    const int lx = gn.y + e.code % cube - W, ly = gn.z + (e.code / cube) % cube - W, lz = gn.w + e.code / (cube * cube) - W;
    const int td = gn.x + E.depthShift;
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

} // namespace

void launch_enum(const EnumParams &E, cudaStream_t st) {
    if (E.nG <= 0) return;
    enum_kernel<<<(E.nG + 7) / 8, 256, 0, st>>>(E);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
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
