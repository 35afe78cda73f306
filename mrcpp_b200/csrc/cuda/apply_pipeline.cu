// Work-list pipeline of the operator application for k = 7 (K = 8): the inner loops of
// ConvolutionCalculator::calcNode / applyOperComp / applyOperator / tensorApplyOperComp
// (src/treebuilders/ConvolutionCalculator.cpp:224-382) split into five kernels so that the FP64
// tensor pipe never waits for screening, barriers or an unbalanced output node:
//
//   screen   warp per (output node, input node) pair: every candidate term is screened with the reference's
//            predicate ((1.0*|O_x|)*|O_y|)*|O_z| * (bandSizeFactor*|f_ft|) > gThrs in the reference's FP64
//            operation order (:283-288, :320-328); result = 64-bit (gt,ft) mask per candidate + tuple counts per
//            (pair, gt, ft).
//   scan     exclusive scans: (pair, ft) segments inside each output block (g, gt), blocks inside the
//            iteration's tuple list, and the split of every block into units of U tuples.
//   fill     warp per pair: surviving tuples are written as 16-byte records (source block, three operator
//            blocks), sorted per output block by (input node, ft, term) -- the reference's accumulation
//            order (calcNode :252-267 loops input node, ft, gt, term) -- so that a source block is
//            fetched once for all of its terms.
//   contract persistent warps pull units from a queue; per tuple 48 DMMA.8x8x4 (three 1-D contractions,
//            6 K^4 = 24 576 flop), accumulators in registers, one partial block per unit.
//   reduce   warp per output block: partial blocks summed in unit order (fixed order: results are run-to-run
//            identical), written to the node store together with the component norm (calcNorms, :270-272).
//
// Fragment algebra of the contraction (lane = 4 r + q): see apply_kernels.cu (same chaining: the D fragment
// of stage 1 is the B fragment of stage 2 under the index permutation sigma; stage 3 contracts the
// tile index after one trip through a padded warp-private shared-memory tile).
#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

constexpr double kMachineZero = 1.0e-14;
constexpr int kTileSi = 18;  // i2 stride (doubles): 144 B == 16 B mod 128 B
constexpr int kTileSm = 152; // m1 stride (doubles): 1216 B == 64 B mod 128 B
constexpr int kTileDoubles = 8 * kTileSm;
constexpr int kContractWarps = 4;     // per CTA
constexpr int kContractCtasPerSm = 3; // 12 warps / SM, <= 168 registers per thread

__device__ __forceinline__ void decode_delta(int code, int W, int d[3]) {
    const int cube = 2 * W + 1;
    d[0] = code % cube - W;
    d[1] = (code / cube) % cube - W;
    d[2] = code / (cube * cube) - W;
}

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
    unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
    unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

// ------------------------------------------------------------------------------------------------ screen
__global__ void __launch_bounds__(256) pipe_screen_kernel(ApplyParams P, PipeBuffers B, int nNbr) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= nNbr) return;
    const NbrEntry nb = P.nbr[w];
    const GDesc g = P.gdesc[nb.g];
    const DepthInfo di = P.depthInfo[g.depth];
    const int *coff = P.candOff + di.cubeOff;
    const int e0 = coff[nb.code];
    const int nc = coff[nb.code + 1] - e0;
    // f component norms; generated nodes carry scaling only (MWNode.cpp:644)
    double fn[8];
    if (nb.fslot < P.nRealF) {
#pragma unroll
        for (int ft = 0; ft < 8; ft++) fn[ft] = P.fNorms[(size_t)nb.fslot * 8 + ft];
    } else {
        fn[0] = P.fGenNorms[nb.fslot - P.nRealF];
#pragma unroll
        for (int ft = 1; ft < 8; ft++) fn[ft] = 0.0;
    }
    unsigned long long fmask = 0ull; // (gt,ft) bits whose f component is not negligible (:254-255)
    double fnMax = 0.0;
#pragma unroll
    for (int ft = 0; ft < 8; ft++) {
        if (!(fn[ft] < kMachineZero)) fmask |= 0x0101010101010101ull << ft;
        fnMax = fmax(fnMax, fn[ft]);
    }
    int d[3];
    decode_delta(nb.code, di.W, d);
    // screening threshold of this output node: one value per apply iteration, or scaled per node (apply with precTrees):
    // gThrs = prec * precFac * sqrt(|g|^2 / M) in the reference's operation order (ConvolutionCalculator.cpp:241-248)
    const double gThrs = (P.precFac != nullptr && P.sqrtTerm >= 0.0) ? P.prec * P.precFac[nb.g] * P.sqrtTerm : P.gThrs;
    int cnt0 = 0, cnt1 = 0; // lane l counts bits l and l + 32
    for (int base = 0; base < nc; base += 32) {
        const int c = base + lane;
        unsigned long long pass = 0ull;
        if (c < nc) {
            const int term = P.candTerm[e0 + c];
            unsigned long long todo = P.candMask[e0 + c] & fmask;
            if (P.derivDir >= 0) {
                pass = todo; // derivative apply has no norm screening (DerivativeCalculator.cpp:211-249)
            } else if (todo) {
                const int nbase = P.nodeBase[(size_t)term * P.DM + g.depth];
                const double4 v0 = *reinterpret_cast<const double4 *>(P.onorms + (size_t)(nbase + d[0]) * 4);
                const double4 v1 = *reinterpret_cast<const double4 *>(P.onorms + (size_t)(nbase + d[1]) * 4);
                const double4 v2 = *reinterpret_cast<const double4 *>(P.onorms + (size_t)(nbase + d[2]) * 4);
                const double n0[4] = {v0.x, v0.y, v0.z, v0.w};
                const double n1[4] = {v1.x, v1.y, v1.z, v1.w};
                const double n2[4] = {v2.x, v2.y, v2.z, v2.w};
                // band size factor of (gt, ft): 64 * prod_d (nodes of component c_d along d), one 16-byte load per candidate
                // instead of 64 scattered table reads
                const int4 s4 = P.bsfSep[(size_t)term * P.DM + g.depth];
                const int sep[4] = {s4.x, s4.y, s4.z, s4.w};
                // whole-candidate early-out: rounded products of non-negative numbers are monotone, so the same expression
                // evaluated on the component-wise maxima bounds every (gt, ft) combination from above -- exact, no slack
                {
                    const double m0 = fmax(fmax(n0[0], n0[1]), fmax(n0[2], n0[3]));
                    const double m1 = fmax(fmax(n1[0], n1[1]), fmax(n1[2], n1[3]));
                    const double m2 = fmax(fmax(n2[0], n2[1]), fmax(n2[2], n2[3]));
                    const int sm = max(max(sep[0], sep[1]), max(sep[2], sep[3]));
                    double oMax = 1.0;
                    oMax *= m0;
                    oMax *= m1;
                    oMax *= m2;
                    const double tMax = (sm * sm * sm * 64) * fnMax;
                    if (!(oMax * tMax > gThrs)) todo = 0ull;
                }
                if (todo)
#pragma unroll
                for (int gt = 0; gt < 8; gt++) {
#pragma unroll
                    for (int ft = 0; ft < 8; ft++) {
                        const int b = gt * 8 + ft;
                        if ((todo >> b) & 1ull) {
                            double oNorm = 1.0;
                            oNorm *= n0[2 * (gt & 1) + (ft & 1)];
                            oNorm *= n1[2 * ((gt >> 1) & 1) + ((ft >> 1) & 1)];
                            oNorm *= n2[2 * ((gt >> 2) & 1) + ((ft >> 2) & 1)];
                            const int bsI = sep[2 * (gt & 1) + (ft & 1)] * sep[2 * ((gt >> 1) & 1) + ((ft >> 1) & 1)] *
                                            sep[2 * ((gt >> 2) & 1) + ((ft >> 2) & 1)] * 64;
                            const double fThreshold = bsI * fn[ft];
                            const double upperBound = oNorm * fThreshold;
                            if (upperBound > gThrs) pass |= 1ull << b;
                        }
                    }
                }
            }
            B.masks[(size_t)nb.candBase + c] = pass;
        }
        unsigned long long uni = warp_or64(pass);
        while (uni) {
            const int b = __ffsll((long long)uni) - 1;
            uni &= uni - 1;
            const int n = __popc(__ballot_sync(0xffffffffu, (pass >> b) & 1ull));
            if (lane == (b & 31)) {
                if (b < 32) cnt0 += n;
                else cnt1 += n;
            }
        }
    }
    B.cnt64[(size_t)w * 64 + lane] = (unsigned short)cnt0;
    B.cnt64[(size_t)w * 64 + 32 + lane] = (unsigned short)cnt1;
}

// ------------------------------------------------------------------------------------------------ scan
// warp per output block (g, gt): offsets of the (input node, ft) segments, input node major
__global__ void __launch_bounds__(256) pipe_segscan_kernel(ApplyParams P, PipeBuffers B, int nBlocks) {
    const int lane = threadIdx.x & 31;
    const int blk = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= nBlocks) return;
    const GDesc g = P.gdesc[blk >> 3];
    const int gt = blk & 7;
    int run = 0;
    for (int n0 = 0; n0 < g.nbrCnt; n0 += 4) {
        const int n = n0 + (lane >> 3), ft = lane & 7;
        const size_t idx = (size_t)(g.nbrOff + n) * 64 + gt * 8 + ft;
        const int v = (n < g.nbrCnt) ? (int)B.cnt64[idx] : 0;
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (n < g.nbrCnt) B.segOff[idx] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) B.blockCnt[blk] = run;
}

// one CTA: tuple offsets of the blocks, unit size, unit offsets of the blocks
__device__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long *sm, unsigned long long &total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long s = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0ull;
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

// Offsets of the output blocks in the tuple list and in the unit list. Two levels: tiles of kBlockTile blocks are scanned
// by one CTA each (coalesced), one CTA scans the tile totals and writes the header; pipe_units_kernel adds the tile bases.
// Unit size: fixed, so that the summation order of a block does not depend on how much other work the launch holds
// (results are bit-identical for any number of ranks sharing the work vector).
constexpr int kBlockTile = 2048; // 256 threads x 8 blocks
static_assert(kBlockTile == kSubRangeBlocks, "sub-range boundaries are block-scan tiles");
constexpr int kTupBits = 36;      // packed tile sums: low 36 bits tuples (an iteration holds < 2^32), high 28 bits units

__global__ void __launch_bounds__(256) pipe_blockscan_tiles_kernel(PipeBuffers B, int nBlocks, int U) {
    __shared__ unsigned long long sm[33];
    const int base = blockIdx.x * kBlockTile + threadIdx.x * 8;
    // low kTupBits bits: tuples, high bits: units of the tile (2048 blocks x <= 17 K units: far below 2^28)
    unsigned long long v[8], local = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const unsigned long long c = (base + i < nBlocks) ? (unsigned long long)B.blockCnt[base + i] : 0ull;
        v[i] = c | (((c + U - 1) / U) << kTupBits);
        local += v[i];
    }
    unsigned long long total;
    unsigned long long run = block_excl_scan(local, sm, total);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (base + i < nBlocks) {
            B.blockTupOff[base + i] = (unsigned)(run & ((1ull << kTupBits) - 1));
            B.blockUnitOff[base + i] = (int)(run >> kTupBits);
        }
        run += v[i];
    }
    if (threadIdx.x == 0) B.tileTotal[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) pipe_blockscan_top_kernel(PipeBuffers B, int nBlocks, int nTiles, int U) {
    __shared__ unsigned long long sm[33];
    const int tid = threadIdx.x;
    const int per = (nTiles + 1023) / 1024;
    const int b0 = min(tid * per, nTiles), b1 = min(b0 + per, nTiles);
    unsigned long long locT = 0, locU = 0;
    for (int b = b0; b < b1; b++) {
        const unsigned long long t = B.tileTotal[b];
        locT += t & ((1ull << kTupBits) - 1);
        locU += t >> kTupBits;
    }
    unsigned long long totT, totU;
    unsigned long long runT = block_excl_scan(locT, sm, totT);
    unsigned long long runU = block_excl_scan(locU, sm, totU);
    for (int b = b0; b < b1; b++) {
        const unsigned long long t = B.tileTotal[b];
        B.tileBaseTup[b] = runT;
        B.tileBaseUnit[b] = (int)runU;
        runT += t & ((1ull << kTupBits) - 1);
        runU += t >> kTupBits;
    }
    __syncthreads(); // tileBaseUnit is complete
    if (tid == 0) {
        B.blockTupOff[nBlocks] = (unsigned)totT;
        B.blockUnitOff[nBlocks] = (int)totU;
        B.header->totalTuples = totT;
        B.header->nUnits = (int)totU;
        B.header->U = U;
        // contraction in sub-ranges (host mirror: an iteration's first nodes go down while its last ones are still contracted):
        // first unit of the block tiles the host picked as boundaries
        for (int s = 0; s < kMaxSubRanges; s++)
            B.header->subUnit[s] = (s < B.nSub && B.subTile[s] < nTiles) ? B.tileBaseUnit[B.subTile[s]] : (int)totU;
    }
}

__global__ void __launch_bounds__(256) pipe_units_kernel(PipeBuffers B, int nBlocks) {
    const int lane = threadIdx.x & 31;
    const int blk = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= nBlocks) return;
    const int U = B.header->U;
    const int cnt = B.blockCnt[blk];
    // tile-local offsets -> global offsets (each entry is read and rewritten by its own warp only)
    const unsigned t0 = B.blockTupOff[blk] + (unsigned)B.tileBaseTup[blk / kBlockTile];
    const int u0 = B.blockUnitOff[blk] + B.tileBaseUnit[blk / kBlockTile];
    __syncwarp();
    if (lane == 0) {
        B.blockTupOff[blk] = t0;
        B.blockUnitOff[blk] = u0;
    }
    const int nu = (cnt + U - 1) / U;
    for (int u = lane; u < nu; u += 32) {
        UnitDesc d;
        d.t0 = t0 + (unsigned)u * U;
        d.cnt = min(U, cnt - u * U);
        B.units[u0 + u] = d;
    }
}

// ------------------------------------------------------------------------------------------------ fill
__global__ void __launch_bounds__(256) pipe_fill_kernel(ApplyParams P, PipeBuffers B, int nNbr) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= nNbr) return;
    const NbrEntry nb = P.nbr[w];
    if (nb.g < B.fillLo || nb.g >= B.fillHi) return; // this pass fills (and queues the gather of) the nodes [fillLo, fillHi)
    const GDesc g = P.gdesc[nb.g];
    const DepthInfo di = P.depthInfo[g.depth];
    const int *coff = P.candOff + di.cubeOff;
    const int e0 = coff[nb.code];
    const int nc = coff[nb.code + 1] - e0;
    int d[3];
    decode_delta(nb.code, di.W, d);
    // lane l keeps the running write position of bits l and l + 32
    unsigned pos0 = B.blockTupOff[nb.g * 8 + (lane >> 3)] + (unsigned)B.segOff[(size_t)w * 64 + lane];
    unsigned pos1 = B.blockTupOff[nb.g * 8 + 4 + (lane >> 3)] + (unsigned)B.segOff[(size_t)w * 64 + 32 + lane];
    const int fbase = (nb.fslot < P.nRealF) ? nb.fslot * 8 : P.nRealF * 8 + (nb.fslot - P.nRealF);
    const bool gen = nb.fslot >= P.nRealF;
    unsigned ftUsed = 0; // ft blocks of the input node that surviving tuples of this lane read
    for (int base = 0; base < nc; base += 32) {
        const int c = base + lane;
        unsigned long long pass = 0ull;
        int oi0 = 0, oi1 = 0, oi2 = 0;
        if (c < nc) {
            pass = B.masks[(size_t)nb.candBase + c];
            {
                unsigned long long m = pass; // fold the gt index away: bit ft of the low byte = some (gt, ft) survives
                m |= m >> 32;
                m |= m >> 16;
                m |= m >> 8;
                ftUsed |= (unsigned)(m & 0xFFull);
            }
            if (pass) {
                const int nbase = P.nodeBase[(size_t)P.candTerm[e0 + c] * P.DM + g.depth];
                oi0 = (nbase + d[0]) * 4;
                oi1 = (nbase + d[1]) * 4;
                oi2 = (nbase + d[2]) * 4;
            }
        }
        unsigned long long uni = warp_or64(pass);
        while (uni) {
            const int b = __ffsll((long long)uni) - 1;
            uni &= uni - 1;
            const bool mine = (pass >> b) & 1ull;
            const unsigned bal = __ballot_sync(0xffffffffu, mine);
            const unsigned start = __shfl_sync(0xffffffffu, (b < 32) ? pos0 : pos1, b & 31);
            if (mine) {
                const int gt = b >> 3, ft = b & 7;
                TupleRec r;
                r.fblk = gen ? fbase : fbase + ft;
                r.o0 = oi0 + 2 * (gt & 1) + (ft & 1);
                r.o1 = oi1 + 2 * ((gt >> 1) & 1) + ((ft >> 1) & 1);
                r.o2 = oi2 + 2 * ((gt >> 2) & 1) + ((ft >> 2) & 1);
                if (P.derivDir >= 0) {
                    // DerivativeCalculator::tensorApplyOperComp (:253-275): only dimension derivDir carries an operator
                    // block, the others are pure index rotations = contraction with the identity (exact in FP64)
                    if (P.derivDir != 0) r.o0 = P.identIdx;
                    if (P.derivDir != 1) r.o1 = P.identIdx;
                    if (P.derivDir != 2) r.o2 = P.identIdx;
                }
                *reinterpret_cast<int4 *>(B.tuples + start + __popc(bal & ((1u << lane) - 1u))) =
                    make_int4(r.fblk, r.o0, r.o1, r.o2);
            }
            if (lane == (b & 31)) {
                if (b < 32) pos0 += __popc(bal);
                else pos1 += __popc(bal);
            }
        }
    }
    // lazy residency, per coefficient BLOCK: the contraction will read these (input node, ft) blocks; queue the ones that are
    // not in HBM yet (a smooth region contributes its scaling block and few wavelet blocks: whole nodes need not cross PCIe)
    if (B.resident != nullptr && !gen) {
        const unsigned used = __reduce_or_sync(0xffffffffu, ftUsed);
        if (lane < 8 && ((used >> lane) & 1u)) {
            const int blk = nb.fslot * 8 + lane;
            if (atomicCAS(&B.resident[blk], 0, 1) == 0) B.fetchList[atomicAdd(B.fetchCnt, 1)] = blk;
        }
    }
}

// ------------------------------------------------------------------------------------------------ contract
struct OpFrag {
    double a00, a01, a10, a11, b20, b21;
};

__device__ __forceinline__ OpFrag load_frags(const double *__restrict__ mats, const int4 rec, int fo) {
    OpFrag f;
    const double *o0 = mats + (size_t)rec.y * 64 + fo;
    const double *o1 = mats + (size_t)rec.z * 64 + fo;
    const double *o2 = mats + (size_t)rec.w * 64 + fo;
    f.a00 = __ldg(o0);
    f.a01 = __ldg(o0 + 4);
    f.a10 = __ldg(o1);
    f.a11 = __ldg(o1 + 4);
    f.b20 = __ldg(o2);
    f.b21 = __ldg(o2 + 4);
    return f;
}

__global__ void __launch_bounds__(kContractWarps * 32, kContractCtasPerSm) pipe_contract_kernel(ApplyParams P, PipeBuffers B, int uBase, int nUnits) {
    extern __shared__ __align__(16) double tiles[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, q = lane & 3;
    const int sig = (r >> 1) + 4 * (r & 1); // sigma(r)
    double *T = tiles + warp * kTileDoubles;
    const int fo = q + 8 * r;         // operator fragment offset: element (q + 4s) + 8 r
    const int bo = q + 8 * sig;       // source fragment offset: f[i0 = q + 4s, i1 = sigma(r), i2 = j]
    const int nReal8 = P.nRealF * 8;
    const int4 *recs = reinterpret_cast<const int4 *>(B.tuples);

    for (;;) {
        int u = 0;
        if (lane == 0) u = uBase + atomicAdd(B.queue, 1); // units [uBase, nUnits) of the iteration
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nUnits) break;
        const UnitDesc ud = B.units[u];
        double acc[8][2];
#pragma unroll
        for (int t = 0; t < 8; t++) acc[t][0] = acc[t][1] = 0.0;
        double bf[8][2];
        int curF = -1;
        int4 nrec = __ldg(recs + ud.t0);
        OpFrag nf = load_frags(P.mats, nrec, fo);
        for (int t = 0; t < ud.cnt; t++) {
            const int4 rec = nrec;
            const OpFrag of = nf;
            if (t + 1 < ud.cnt) {
                nrec = __ldg(recs + ud.t0 + t + 1);
                nf = load_frags(P.mats, nrec, fo);
            }
            if (rec.x != curF) {
                curF = rec.x;
                const double *fblk = (curF < nReal8) ? P.fReal + (size_t)curF * 512 : P.fGen + (size_t)(curF - nReal8) * 512;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    bf[j][0] = __ldg(fblk + bo + 64 * j);
                    bf[j][1] = __ldg(fblk + bo + 4 + 64 * j);
                }
            }
            // stage 1: X1[m0=r][i1=q+4e][i2=j]; stage 2: X2[m1=r][m0=2q+e][i2=j]; then T[m1][i2][m0]
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double d10 = 0.0, d11 = 0.0;
                dmma884(d10, d11, of.a00, bf[j][0]);
                dmma884(d10, d11, of.a01, bf[j][1]);
                double d20 = 0.0, d21 = 0.0;
                dmma884(d20, d21, of.a10, d10);
                dmma884(d20, d21, of.a11, d11);
                *reinterpret_cast<double2 *>(T + 2 * q + kTileSi * j + kTileSm * r) = make_double2(d20, d21);
            }
            __syncwarp();
            // stage 3: g[m0=t][m1=r][m2=2q+e] += sum_i2 X2 * O2[i2][m2]
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double2 x0 = *reinterpret_cast<const double2 *>(T + 2 * v + kTileSi * q + kTileSm * r);
                const double2 x1 = *reinterpret_cast<const double2 *>(T + 2 * v + kTileSi * (q + 4) + kTileSm * r);
                dmma884(acc[2 * v][0], acc[2 * v][1], x0.x, of.b20);
                dmma884(acc[2 * v][0], acc[2 * v][1], x1.x, of.b21);
                dmma884(acc[2 * v + 1][0], acc[2 * v + 1][1], x0.y, of.b20);
                dmma884(acc[2 * v + 1][0], acc[2 * v + 1][1], x1.y, of.b21);
            }
            __syncwarp();
        }
        // partial block of this unit: element m0 + 8 m1 + 64 m2
        double *pb = B.partials + (size_t)u * 512 + 8 * r + 64 * (2 * q);
#pragma unroll
        for (int v = 0; v < 4; v++) {
            *reinterpret_cast<double2 *>(pb + 2 * v) = make_double2(acc[2 * v][0], acc[2 * v + 1][0]);
            *reinterpret_cast<double2 *>(pb + 64 + 2 * v) = make_double2(acc[2 * v][1], acc[2 * v + 1][1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ contract, other orders
// Even K != 8 (k = 3, 5, 9, 11): FP64 FMA path on the same work lists. DMMA tiles are 8 x 8 x 4, so K = 10 / 12 would
// waste 50 % / 25 % of the (shared) FP64 unit on padding, while the DFMA rate of sm_100a equals the DMMA rate
// (profiles/r01_dmma_dfma_probe.txt); a register-tiled FMA contraction is the faster choice there.
// One warp per unit. A stage computes out[r + K^2 c] = sum_t in[K r + t] * op[t + K c] (tensorApplyOperComp,
// ConvolutionCalculator.cpp:363-379: g = f^T O): lane owns rows r = lane + 32 i, keeps two rows (2 K inputs) in
// registers and streams the operator block through uniform (broadcast) loads, so one 16-byte operator load feeds
// four FMAs. Stages 1 and 2 write warp-private shared scratch; stage 3 accumulates into registers that stay live
// for the whole unit.
template <int K, int R, bool ACC>
__device__ __forceinline__ void fma_stage(const double *__restrict__ in, const double *__restrict__ op, double *__restrict__ out, int lane) {
    constexpr int K2 = K * K;
#pragma unroll 1
    for (int i0 = 0; i0 < R; i0 += 2) { // not unrolled: the operator block is re-streamed per row pair instead of pinned in registers
        const int r0 = lane + 32 * i0, r1 = r0 + 32;
        const bool v0 = r0 < K2, v1 = (i0 + 1 < R) && (r1 < K2);
        double x0[K], x1[K];
#pragma unroll
        for (int t = 0; t < K; t += 2) {
            const double2 a = v0 ? *reinterpret_cast<const double2 *>(in + K * r0 + t) : make_double2(0.0, 0.0);
            const double2 b = v1 ? *reinterpret_cast<const double2 *>(in + K * r1 + t) : make_double2(0.0, 0.0);
            x0[t] = a.x;
            x0[t + 1] = a.y;
            x1[t] = b.x;
            x1[t + 1] = b.y;
        }
        // K independent accumulator pairs: the operator block streams through uniform 16-byte loads, each feeding 4 FMAs
        double a0[K], a1[K];
#pragma unroll
        for (int c = 0; c < K; c++) a0[c] = a1[c] = 0.0;
#pragma unroll
        for (int t = 0; t < K; t += 2) {
#pragma unroll
            for (int c = 0; c < K; c++) {
                const double2 w = __ldg(reinterpret_cast<const double2 *>(op + K * c + t));
                a0[c] = fma(x0[t], w.x, a0[c]);
                a1[c] = fma(x1[t], w.x, a1[c]);
                a0[c] = fma(x0[t + 1], w.y, a0[c]);
                a1[c] = fma(x1[t + 1], w.y, a1[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < K; c++) {
            if (ACC) {
                if (v0) out[r0 + K2 * c] += a0[c];
                if (v1) out[r1 + K2 * c] += a1[c];
            } else {
                if (v0) out[r0 + K2 * c] = a0[c];
                if (v1) out[r1 + K2 * c] = a1[c];
            }
        }
    }
}

template <int K> __global__ void __launch_bounds__(256, 1) pipe_contract_fma_kernel(ApplyParams P, PipeBuffers B, int nUnits, int nWarps) {
    extern __shared__ __align__(16) double scratch[];
    constexpr int K2 = K * K, Kd = K2 * K, R = (K2 + 31) / 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nWarps) return;
    double *S1 = scratch + (size_t)warp * 3 * Kd;
    double *S2 = S1 + Kd;
    double *S3 = S2 + Kd; // accumulators of the unit's output block (lane-private rows: no synchronisation needed)
    const long long nReal8 = (long long)P.nRealF * 8;
    const int4 *recs = reinterpret_cast<const int4 *>(B.tuples);
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(B.queue, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nUnits) break;
        const UnitDesc ud = B.units[u];
        for (int e = lane; e < Kd; e += 32) S3[e] = 0.0;
        __syncwarp();
        for (int t = 0; t < ud.cnt; t++) {
            const int4 rec = __ldg(recs + ud.t0 + t);
            const double *fblk = (rec.x < nReal8) ? P.fReal + (size_t)rec.x * Kd : P.fGen + (size_t)(rec.x - nReal8) * Kd;
            fma_stage<K, R, false>(fblk, P.mats + (size_t)rec.y * K2, S1, lane);
            __syncwarp();
            fma_stage<K, R, false>(S1, P.mats + (size_t)rec.z * K2, S2, lane);
            __syncwarp();
            fma_stage<K, R, true>(S2, P.mats + (size_t)rec.w * K2, S3, lane);
            __syncwarp();
        }
        double *pb = B.partials + (size_t)u * Kd;
        for (int e = lane; e < Kd; e += 32) pb[e] = S3[e];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ contract, other orders (DMMA, padded)
// The same three contractions on the FP64 tensor cores for K != 8. A stage is the GEMM D[c][col] = sum_t A[c][t] B[t][col]
// with A = O^T (K x K, rows padded to 8 MT, inner dimension to 4 KS with zeros) and B the K x K^2 view of the source
// (K^2 columns in NT tiles of 8). Useful fraction of the issued DMMA work: K = 12: 75 %, K = 10: 50 %, K = 6: 50 %.
// Unlike the FMA variant above the operator costs MT x KS registers per stage instead of K^2 / 2 uniform loads.
// Stage outputs live in warp-private shared memory as out[col + CP c] (CP = K^2 padded to 8 mod 16 doubles: the 16-byte
// stores of a quarter warp then hit 8 distinct bank groups); the next stage reads them back as B fragments
// (t = col % K, row = col / K + K c). The third stage accumulates in registers for the whole unit.
template <int K> struct PadDims {
    static constexpr int K2 = K * K, Kd = K2 * K;
    static constexpr int MT = (K + 7) / 8, KS = (K + 3) / 4, NT = (K2 + 7) / 8;
    static constexpr int CP = ((K2 + 7) / 16) * 16 + 8; // >= K2, == 8 (mod 16)
};

template <int K, bool DENSE, bool ACC>
__device__ __forceinline__ void dmma_stage(const double *__restrict__ in, const double *__restrict__ op, double *__restrict__ out,
                                           double (&acc)[PadDims<K>::NT][PadDims<K>::MT][2], int rr, int q) {
    using D = PadDims<K>;
    // operator fragments: A[c = rr + 8 mt][t = q + 4 s] = op[t + K c]
    double a[D::MT][D::KS];
#pragma unroll
    for (int mt = 0; mt < D::MT; mt++)
#pragma unroll
        for (int s = 0; s < D::KS; s++) {
            const int c = rr + 8 * mt, t = q + 4 * s;
            a[mt][s] = (c < K && t < K) ? __ldg(op + t + K * c) : 0.0;
        }
#pragma unroll
    for (int n0 = 0; n0 < D::NT; n0++) {
        const int r = 8 * n0 + rr; // source row = column of the GEMM
        const int rowAddr = DENSE ? K * r : K * (r % K) + D::CP * (r / K);
        double b[D::KS];
#pragma unroll
        for (int s = 0; s < D::KS; s++) {
            const int t = q + 4 * s;
            b[s] = (r < D::K2 && t < K) ? in[rowAddr + t] : 0.0;
        }
#pragma unroll
        for (int mt = 0; mt < D::MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
            if (ACC) {
                d0 = acc[n0][mt][0];
                d1 = acc[n0][mt][1];
            }
#pragma unroll
            for (int s = 0; s < D::KS; s++) dmma884(d0, d1, a[mt][s], b[s]);
            if (ACC) {
                acc[n0][mt][0] = d0;
                acc[n0][mt][1] = d1;
            } else {
                const int c = rr + 8 * mt, col = 8 * n0 + 2 * q;
                if (c < K && col < D::K2) *reinterpret_cast<double2 *>(out + col + D::CP * c) = make_double2(d0, d1);
            }
        }
    }
}

template <int K> __global__ void __launch_bounds__(256, 1) pipe_contract_pad_kernel(ApplyParams P, PipeBuffers B, int nUnits, int nWarps) {
    using D = PadDims<K>;
    extern __shared__ __align__(16) double scratch[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= nWarps) return;
    const int rr = lane >> 2, q = lane & 3;
    double *S1 = scratch + (size_t)warp * 2 * K * D::CP;
    double *S2 = S1 + K * D::CP;
    const long long nReal8 = (long long)P.nRealF * 8;
    const int4 *recs = reinterpret_cast<const int4 *>(B.tuples);
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(B.queue, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= nUnits) break;
        const UnitDesc ud = B.units[u];
        double acc[D::NT][D::MT][2];
#pragma unroll
        for (int n0 = 0; n0 < D::NT; n0++)
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) acc[n0][mt][0] = acc[n0][mt][1] = 0.0;
        for (int t = 0; t < ud.cnt; t++) {
            const int4 rec = __ldg(recs + ud.t0 + t);
            const double *fblk = (rec.x < nReal8) ? P.fReal + (size_t)rec.x * D::Kd : P.fGen + (size_t)(rec.x - nReal8) * D::Kd;
            dmma_stage<K, true, false>(fblk, P.mats + (size_t)rec.y * D::K2, S1, acc, rr, q);
            __syncwarp();
            dmma_stage<K, false, false>(S1, P.mats + (size_t)rec.z * D::K2, S2, acc, rr, q);
            __syncwarp();
            dmma_stage<K, false, true>(S2, P.mats + (size_t)rec.w * D::K2, nullptr, acc, rr, q);
            __syncwarp();
        }
        // partial block of this unit, dense: element col + K^2 c (col = m0 + K m1, c = m2)
        double *pb = B.partials + (size_t)u * D::Kd;
#pragma unroll
        for (int n0 = 0; n0 < D::NT; n0++)
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) {
                const int c = rr + 8 * mt, col = 8 * n0 + 2 * q;
                if (c < K && col < D::K2)
                    *reinterpret_cast<double2 *>(pb + col + D::K2 * c) = make_double2(acc[n0][mt][0], acc[n0][mt][1]);
            }
    }
}

// The same padded-DMMA contraction with a CTA of four warps working on ONE unit: the n-tiles (columns) of every stage are
// dealt out to the warps, the stage outputs live in CTA-shared memory. A unit then needs 2 K CP doubles of shared memory per
// CTA instead of per warp, so 5-7 CTAs (20-28 warps) fit on an SM instead of 8 warps, and the DMMA latency is hidden by
// other warps instead of by instruction-level parallelism alone. Summation order per output element is unchanged (tuples in
// list order), so results are bit-identical to the warp-private kernel.
template <int K, bool DENSE, bool ACC, int NTW>
__device__ __forceinline__ void dmma_stage_coop(const double *__restrict__ in, const double *__restrict__ op, double *__restrict__ out,
                                                double (&acc)[NTW][PadDims<K>::MT][2], int rr, int q, int warp) {
    using D = PadDims<K>;
    double a[D::MT][D::KS];
#pragma unroll
    for (int mt = 0; mt < D::MT; mt++)
#pragma unroll
        for (int s = 0; s < D::KS; s++) {
            const int c = rr + 8 * mt, t = q + 4 * s;
            a[mt][s] = (c < K && t < K) ? __ldg(op + t + K * c) : 0.0;
        }
#pragma unroll
    for (int i = 0; i < NTW; i++) {
        const int n0 = warp + 4 * i;
        if (n0 < D::NT) {
            const int r = 8 * n0 + rr;
            const int rowAddr = DENSE ? K * r : K * (r % K) + D::CP * (r / K);
            double b[D::KS];
#pragma unroll
            for (int s = 0; s < D::KS; s++) {
                const int t = q + 4 * s;
                b[s] = (r < D::K2 && t < K) ? in[rowAddr + t] : 0.0;
            }
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) {
                double d0 = 0.0, d1 = 0.0;
                if (ACC) {
                    d0 = acc[i][mt][0];
                    d1 = acc[i][mt][1];
                }
#pragma unroll
                for (int s = 0; s < D::KS; s++) dmma884(d0, d1, a[mt][s], b[s]);
                if (ACC) {
                    acc[i][mt][0] = d0;
                    acc[i][mt][1] = d1;
                } else {
                    const int c = rr + 8 * mt, col = 8 * n0 + 2 * q;
                    if (c < K && col < D::K2) *reinterpret_cast<double2 *>(out + col + D::CP * c) = make_double2(d0, d1);
                }
            }
        }
    }
}

template <int K> __global__ void __launch_bounds__(128) pipe_contract_coop_kernel(ApplyParams P, PipeBuffers B, int nUnits) {
    using D = PadDims<K>;
    constexpr int NTW = (D::NT + 3) / 4;
    extern __shared__ __align__(16) double scratch[];
    __shared__ int sUnit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = lane >> 2, q = lane & 3;
    double *S1 = scratch;
    double *S2 = S1 + K * D::CP;
    const long long nReal8 = (long long)P.nRealF * 8;
    const int4 *recs = reinterpret_cast<const int4 *>(B.tuples);
    for (;;) {
        if (threadIdx.x == 0) sUnit = atomicAdd(B.queue, 1);
        __syncthreads();
        const int u = sUnit;
        __syncthreads();
        if (u >= nUnits) break;
        const UnitDesc ud = B.units[u];
        double acc[NTW][D::MT][2];
#pragma unroll
        for (int i = 0; i < NTW; i++)
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) acc[i][mt][0] = acc[i][mt][1] = 0.0;
        for (int t = 0; t < ud.cnt; t++) {
            const int4 rec = __ldg(recs + ud.t0 + t);
            const double *fblk = (rec.x < nReal8) ? P.fReal + (size_t)rec.x * D::Kd : P.fGen + (size_t)(rec.x - nReal8) * D::Kd;
            dmma_stage_coop<K, true, false, NTW>(fblk, P.mats + (size_t)rec.y * D::K2, S1, acc, rr, q, warp);
            __syncthreads();
            dmma_stage_coop<K, false, false, NTW>(S1, P.mats + (size_t)rec.z * D::K2, S2, acc, rr, q, warp);
            __syncthreads();
            dmma_stage_coop<K, false, true, NTW>(S2, P.mats + (size_t)rec.w * D::K2, nullptr, acc, rr, q, warp);
            __syncthreads(); // the next tuple's second stage overwrites S2
        }
        double *pb = B.partials + (size_t)u * D::Kd;
#pragma unroll
        for (int i = 0; i < NTW; i++) {
            const int n0 = warp + 4 * i;
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) {
                const int c = rr + 8 * mt, col = 8 * n0 + 2 * q;
                if (n0 < D::NT && c < K && col < D::K2) {
                    if (K & 1) { // odd K: rows of the dense block are not 16-byte aligned, K^2 is odd
                        pb[col + D::K2 * c] = acc[i][mt][0];
                        if (col + 1 < D::K2) pb[col + 1 + D::K2 * c] = acc[i][mt][1];
                    } else {
                        *reinterpret_cast<double2 *>(pb + col + D::K2 * c) = make_double2(acc[i][mt][0], acc[i][mt][1]);
                    }
                }
            }
        }
    }
}

// Stage 1 stacked along M. Consecutive tuples of a unit are sorted by (input node, ft, term), so runs of tuples share their
// SOURCE block and differ only in the operator: the first contraction of up to G such tuples is ONE GEMM
//   [O_x^(1)^T; ...; O_x^(G)^T] (G K x K)  .  f (K x K^2)
// whose G K rows fill whole 8-row tiles (K = 12, G = 2: 24 rows = 3 tiles instead of 2 x 2 padded ones; K = 10, G = 4: 40 rows =
// 5 tiles instead of 4 x 2; K = 6, G = 4: 24 rows = 3 tiles instead of 4 x 1) and reads the source fragments once. Stages 2 and 3
// then run per tuple as in pipe_contract_coop_kernel. Every output element is the same sequence of DMMA k-steps over the same
// operands as in the unstacked kernel (padding contributes exact zeros), and tuples are accumulated in list order: results
// are bit-identical to it.
template <int K, int G, int MINB> __global__ void __launch_bounds__(128, MINB) pipe_contract_stack_kernel(ApplyParams P, PipeBuffers B, int nUnits) {
    using D = PadDims<K>;
    constexpr int NTW = (D::NT + 3) / 4;
    constexpr int GMT = (G * K + 7) / 8; // m-tiles of the stacked first stage
    extern __shared__ __align__(16) double scratch[];
    __shared__ int sUnit;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rr = lane >> 2, q = lane & 3;
    double *S1 = scratch;                 // G stage-1 outputs, K * CP doubles each
    double *S2 = S1 + G * K * D::CP;      // stage-2 output of the tuple in flight
    const long long nReal8 = (long long)P.nRealF * 8;
    const int4 *recs = reinterpret_cast<const int4 *>(B.tuples);
    for (;;) {
        if (threadIdx.x == 0) sUnit = atomicAdd(B.queue, 1);
        __syncthreads();
        const int u = sUnit;
        __syncthreads();
        if (u >= nUnits) break;
        const UnitDesc ud = B.units[u];
        double acc[NTW][D::MT][2];
#pragma unroll
        for (int i = 0; i < NTW; i++)
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) acc[i][mt][0] = acc[i][mt][1] = 0.0;
        for (int t = 0; t < ud.cnt;) {
            // ---- group: up to G consecutive tuples reading the same source block (uniform over the CTA)
            int4 rec[G];
            rec[0] = __ldg(recs + ud.t0 + t);
            int g = 1;
#pragma unroll
            for (int i = 1; i < G; i++) {
                rec[i] = rec[0];
                if (g == i && t + i < ud.cnt) {
                    const int4 r2 = __ldg(recs + ud.t0 + t + i);
                    if (r2.x == rec[0].x) {
                        rec[i] = r2;
                        g = i + 1;
                    }
                }
            }
            const double *fblk = (rec[0].x < nReal8) ? P.fReal + (size_t)rec[0].x * D::Kd : P.fGen + (size_t)(rec[0].x - nReal8) * D::Kd;
            // ---- stage 1, stacked: A[row = rr + 8 mt][t = q + 4 s] = O_x^(row / K)[t + K (row % K)]
            {
                double a[GMT][D::KS];
#pragma unroll
                for (int mt = 0; mt < GMT; mt++) {
                    const int row = rr + 8 * mt, gi = row / K, c = row - gi * K;
                    int oy = rec[0].y;
#pragma unroll
                    for (int i = 1; i < G; i++)
                        if (gi == i) oy = rec[i].y;
                    const double *op = P.mats + (size_t)oy * D::K2;
#pragma unroll
                    for (int s = 0; s < D::KS; s++) {
                        const int tt = q + 4 * s;
                        a[mt][s] = (gi < g && tt < K) ? __ldg(op + tt + K * c) : 0.0;
                    }
                }
#pragma unroll
                for (int i = 0; i < NTW; i++) {
                    const int n0 = warp + 4 * i;
                    if (n0 < D::NT) {
                        const int r = 8 * n0 + rr;
                        double b[D::KS];
#pragma unroll
                        for (int s = 0; s < D::KS; s++) {
                            const int tt = q + 4 * s;
                            b[s] = (r < D::K2 && tt < K) ? fblk[K * r + tt] : 0.0;
                        }
#pragma unroll
                        for (int mt = 0; mt < GMT; mt++) {
                            if (8 * mt < g * K) {
                                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                                for (int s = 0; s < D::KS; s++) dmma884(d0, d1, a[mt][s], b[s]);
                                const int row = rr + 8 * mt, gi = row / K, c = row - gi * K, col = 8 * n0 + 2 * q;
                                if (gi < g && col < D::K2)
                                    *reinterpret_cast<double2 *>(S1 + (size_t)gi * K * D::CP + col + D::CP * c) = make_double2(d0, d1);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- stages 2 and 3 per tuple of the group, list order
#pragma unroll
            for (int gi = 0; gi < G; gi++) {
                if (gi < g) {
                    dmma_stage_coop<K, false, false, NTW>(S1 + (size_t)gi * K * D::CP, P.mats + (size_t)rec[gi].z * D::K2, S2, acc, rr, q, warp);
                    __syncthreads();
                    dmma_stage_coop<K, false, true, NTW>(S2, P.mats + (size_t)rec[gi].w * D::K2, nullptr, acc, rr, q, warp);
                    __syncthreads(); // the next tuple's second stage overwrites S2; the next group's first stage overwrites S1
                }
            }
            t += g;
        }
        double *pb = B.partials + (size_t)u * D::Kd;
#pragma unroll
        for (int i = 0; i < NTW; i++) {
            const int n0 = warp + 4 * i;
#pragma unroll
            for (int mt = 0; mt < D::MT; mt++) {
                const int c = rr + 8 * mt, col = 8 * n0 + 2 * q;
                if (n0 < D::NT && c < K && col < D::K2) *reinterpret_cast<double2 *>(pb + col + D::K2 * c) = make_double2(acc[i][mt][0], acc[i][mt][1]);
            }
        }
    }
}

template <int K, int G, int MINB = 0> void launch_contract_stack(const ApplyParams &P, const PipeBuffers &B, int nUnits, cudaStream_t st) {
    using D = PadDims<K>;
    static_assert((K & 1) == 0, "stacked contraction: even K only (16-byte aligned dense blocks)");
    static int grid = 0;
    const size_t bytes = (size_t)(G + 1) * K * D::CP * sizeof(double);
    if (!grid) {
        int dev = 0, sms = 0, perSm = 0;
        MRX_CUDA(cudaGetDevice(&dev));
        MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MRX_CUDA(cudaFuncSetAttribute(pipe_contract_stack_kernel<K, G, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        MRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, pipe_contract_stack_kernel<K, G, MINB>, 128, bytes));
        grid = sms * std::max(perSm, 1);
    }
    pipe_contract_stack_kernel<K, G, MINB><<<std::min(grid, nUnits), 128, bytes, st>>>(P, B, nUnits);
}

template <int K> void launch_contract_coop(const ApplyParams &P, const PipeBuffers &B, int nUnits, cudaStream_t st) {
    using D = PadDims<K>;
    static int grid = 0;
    const size_t bytes = (size_t)2 * K * D::CP * sizeof(double);
    if (!grid) {
        int dev = 0, sms = 0, perSm = 0;
        MRX_CUDA(cudaGetDevice(&dev));
        MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MRX_CUDA(cudaFuncSetAttribute(pipe_contract_coop_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        MRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, pipe_contract_coop_kernel<K>, 128, bytes));
        grid = sms * std::max(perSm, 1);
    }
    pipe_contract_coop_kernel<K><<<std::min(grid, nUnits), 128, bytes, st>>>(P, B, nUnits);
}

// ------------------------------------------------------------------------------------------------ reduce
// stageRows != nullptr (sharded apply): the block goes to row (blk >> 3) of the rank's segment of the exchange staging buffer
// instead of the node store; the whole iteration is unpacked into the node store after the peers' rows have arrived
__global__ void __launch_bounds__(256) pipe_reduce_kernel(PipeBuffers B, const int *__restrict__ gslots, double *__restrict__ gCoefs,
                                                          double *__restrict__ gNorms, double *__restrict__ gNormsW, int blkBase, int nBlocks,
                                                          int Kd, double *__restrict__ stageRows) {
    const int lane = threadIdx.x & 31;
    const int blk = blkBase + blockIdx.x * 8 + (threadIdx.x >> 5); // blocks [blkBase, nBlocks) of the rank's items
    if (blk >= nBlocks) return;
    const int u0 = B.blockUnitOff[blk], u1 = B.blockUnitOff[blk + 1];
    const int slot = gslots[blk >> 3], gt = blk & 7;
    double *oS = stageRows ? stageRows + (size_t)blk * Kd : gCoefs + ((size_t)slot * 8 + gt) * Kd;
    double2 *o = reinterpret_cast<double2 *>(oS);
    double n2 = 0.0;
    if (Kd & 1) {
        // odd K: blocks are not 16-byte aligned; element-wise, 8 elements per lane and trip (same unit order)
        for (int base = 0; base < Kd; base += 256) {
            double s1[8];
#pragma unroll
            for (int i = 0; i < 8; i++) s1[i] = 0.0;
            for (int u = u0; u < u1; u++) {
                const double *p = B.partials + (size_t)u * Kd;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int e = base + i * 32 + lane;
                    if (e < Kd) s1[i] += p[e];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int e = base + i * 32 + lane;
                if (e < Kd) {
                    oS[e] = s1[i];
                    n2 = fma(s1[i], s1[i], n2);
                }
            }
        }
    } else
    // even K: a block is a whole number of 16-byte pairs; 8 pairs per lane and trip
    for (int base = 0; base < Kd / 2; base += 256) {
        double2 s[8];
#pragma unroll
        for (int i = 0; i < 8; i++) s[i] = make_double2(0.0, 0.0);
        for (int u = u0; u < u1; u++) { // unit order = tuple order: fixed summation order
            const double2 *p = reinterpret_cast<const double2 *>(B.partials + (size_t)u * Kd);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int e = base + i * 32 + lane;
                if (e < Kd / 2) {
                    const double2 v = p[e];
                    s[i].x += v.x;
                    s[i].y += v.y;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = base + i * 32 + lane;
            if (e < Kd / 2) {
                o[e] = s[i];
                n2 = fma(s[i].x, s[i].x, n2);
                n2 = fma(s[i].y, s[i].y, n2);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
    if (lane == 0) {
        const double nrm = sqrt(n2);
        gNorms[(size_t)slot * 8 + gt] = nrm;
        gNormsW[blk] = nrm; // work-vector order: what the host's norm bookkeeping and the rank exchange read
    }
}

// sharded apply: one iteration's output nodes from the rank-major staging buffer (row r * rowsPerRank + j holds work-vector
// item shard_item(r, j), the block-cyclic distribution of the work vector) into the node store
__global__ void __launch_bounds__(256) unpack_nodes_kernel(double *__restrict__ coefs, const double *__restrict__ stage,
                                                           const int *__restrict__ gslotsAll, int nG, int world, int rowsPerRank,
                                                           int shardB, int ncoef, const double *__restrict__ normRows, double *__restrict__ gNorms) {
    const int r = blockIdx.x / rowsPerRank, j = blockIdx.x - r * rowsPerRank;
    const int i = shard_item(r, j, world, shardB);
    if (i >= nG) return;
    const int slot = gslotsAll[i];
    const double2 *src = reinterpret_cast<const double2 *>(stage + (size_t)blockIdx.x * ncoef);
    double2 *dst = reinterpret_cast<double2 *>(coefs + (size_t)slot * ncoef);
    for (int e = threadIdx.x; e < ncoef / 2; e += 256) dst[e] = src[e];
    // component norms of the node (all-gathered in the same rank-major layout) into the node store
    if (threadIdx.x < 8) gNorms[(size_t)slot * 8 + threadIdx.x] = normRows[(size_t)blockIdx.x * 8 + threadIdx.x];
}

} // namespace

void launch_pipe_screen(const ApplyParams &P, const PipeBuffers &B, int nNbr, cudaStream_t st) {
    if (nNbr <= 0) return;
    pipe_screen_kernel<<<(nNbr + 7) / 8, 256, 0, st>>>(P, B, nNbr);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_pipe_scan(const ApplyParams &P, const PipeBuffers &B, int nG, int unitHint, cudaStream_t st) {
    const int nBlocks = nG * 8;
    const int nTiles = (nBlocks + kBlockTile - 1) / kBlockTile;
    if (nBlocks > 0) {
        pipe_segscan_kernel<<<(nBlocks + 7) / 8, 256, 0, st>>>(P, B, nBlocks);
        pipe_blockscan_tiles_kernel<<<nTiles, 256, 0, st>>>(B, nBlocks, unitHint);
        launch_counter() += 2;
    }
    pipe_blockscan_top_kernel<<<1, 1024, 0, st>>>(B, nBlocks, nTiles, unitHint);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_pipe_units(const PipeBuffers &B, int nG, cudaStream_t st) {
    const int nBlocks = nG * 8;
    if (nBlocks > 0) {
        pipe_units_kernel<<<(nBlocks + 7) / 8, 256, 0, st>>>(B, nBlocks);
        launch_counter()++;
    }
    MRX_CUDA(cudaGetLastError());
}

void launch_pipe_fill(const ApplyParams &P, const PipeBuffers &B, int nNbr, cudaStream_t st) {
    if (nNbr > 0) {
        pipe_fill_kernel<<<(nNbr + 7) / 8, 256, 0, st>>>(P, B, nNbr);
        launch_counter()++;
    }
    MRX_CUDA(cudaGetLastError());
}

int pipe_contract_grid() {
    static int grid = 0;
    if (!grid) {
        int dev = 0, sms = 0;
        MRX_CUDA(cudaGetDevice(&dev));
        MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MRX_CUDA(cudaFuncSetAttribute(pipe_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(kContractWarps * kTileDoubles * sizeof(double))));
        grid = sms * kContractCtasPerSm;
    }
    return grid;
}

int pipe_contract_warps() { return pipe_contract_grid() * kContractWarps; }

template <int K> void launch_contract_fma(const ApplyParams &P, const PipeBuffers &B, int nUnits, cudaStream_t st) {
    constexpr int Kd = K * K * K;
    static int sms = 0, nWarps = 0;
    static size_t bytes = 0;
    if (!sms) {
        int dev = 0;
        MRX_CUDA(cudaGetDevice(&dev));
        MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        nWarps = 8;
        while (nWarps > 1 && (size_t)nWarps * 3 * Kd * sizeof(double) > 220 * 1024) nWarps--;
        bytes = (size_t)nWarps * 3 * Kd * sizeof(double);
        MRX_CUDA(cudaFuncSetAttribute(pipe_contract_fma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    const int grid = std::min(sms, (nUnits + nWarps - 1) / nWarps);
    pipe_contract_fma_kernel<K><<<grid, 256, bytes, st>>>(P, B, nUnits, nWarps);
}

template <int K> void launch_contract_pad(const ApplyParams &P, const PipeBuffers &B, int nUnits, cudaStream_t st) {
    using D = PadDims<K>;
    static int sms = 0, nWarps = 0;
    static size_t bytes = 0;
    if (!sms) {
        int dev = 0;
        MRX_CUDA(cudaGetDevice(&dev));
        MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        nWarps = 8;
        const size_t perWarp = (size_t)2 * K * D::CP * sizeof(double);
        while (nWarps > 1 && (size_t)nWarps * perWarp > 220 * 1024) nWarps--;
        bytes = (size_t)nWarps * perWarp;
        MRX_CUDA(cudaFuncSetAttribute(pipe_contract_pad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    const int grid = std::min(sms, (nUnits + nWarps - 1) / nWarps);
    pipe_contract_pad_kernel<K><<<grid, 256, bytes, st>>>(P, B, nUnits, nWarps);
}

bool pipe_supports_order(int K) { return K >= 4 && K <= 12; }

void launch_pipe_contract(const ApplyParams &P, const PipeBuffers &B, int nUnits, cudaStream_t st, int uBase) {
    if (nUnits - uBase <= 0) return;
    if (uBase != 0 && P.K != 8) MRX_ABORT("pipeline contraction: unit sub-ranges are implemented for the k = 7 kernel only");
    MRX_CUDA(cudaMemsetAsync(B.queue, 0, sizeof(int), st));
    static const bool useFma = getenv("MRX_FMA") != nullptr; // development switch: FMA instead of padded-DMMA contraction
    static const bool warpPrivate = getenv("MRX_PAD_WARP") != nullptr; // development switch: one warp per unit instead of one CTA
    static const bool noStack = getenv("MRX_NO_STACK") != nullptr;     // development switch: unstacked first stage
    if (P.K == 8) {
        const int grid = std::min(pipe_contract_grid(), (nUnits - uBase + kContractWarps - 1) / kContractWarps);
        pipe_contract_kernel<<<grid, kContractWarps * 32, kContractWarps * kTileDoubles * sizeof(double), st>>>(P, B, uBase, nUnits);
    } else if (P.K == 4) {
        launch_contract_fma<4>(P, B, nUnits, st);
    } else if (P.K == 6) {
        if (useFma) launch_contract_fma<6>(P, B, nUnits, st);
        else if (warpPrivate) launch_contract_pad<6>(P, B, nUnits, st);
        else if (noStack) launch_contract_coop<6>(P, B, nUnits, st);
        else launch_contract_stack<6, 4>(P, B, nUnits, st);
    } else if (P.K == 10) {
        if (useFma) launch_contract_fma<10>(P, B, nUnits, st);
        else if (warpPrivate) launch_contract_pad<10>(P, B, nUnits, st);
        else if (noStack) launch_contract_coop<10>(P, B, nUnits, st);
        else {
            static const int variant = getenv("MRX_STACK_VARIANT") ? atoi(getenv("MRX_STACK_VARIANT")) : 0; // development switch
            if (variant == 1) launch_contract_stack<10, 4, 5>(P, B, nUnits, st);
            else if (variant == 2) launch_contract_stack<10, 2, 0>(P, B, nUnits, st);
            else if (variant == 3) launch_contract_stack<10, 2, 6>(P, B, nUnits, st);
            else launch_contract_stack<10, 4, 0>(P, B, nUnits, st);
        }
    } else if (P.K == 12) {
        if (useFma) launch_contract_fma<12>(P, B, nUnits, st);
        else if (warpPrivate) launch_contract_pad<12>(P, B, nUnits, st);
        else if (noStack) launch_contract_coop<12>(P, B, nUnits, st);
        else launch_contract_stack<12, 2>(P, B, nUnits, st);
    } else if (P.K == 5) { // odd K (even polynomial orders): same padded-DMMA kernel, element-wise partial blocks
        launch_contract_coop<5>(P, B, nUnits, st);
    } else if (P.K == 7) {
        launch_contract_coop<7>(P, B, nUnits, st);
    } else if (P.K == 9) {
        launch_contract_coop<9>(P, B, nUnits, st);
    } else if (P.K == 11) {
        launch_contract_coop<11>(P, B, nUnits, st);
    } else {
        MRX_ABORT("pipeline contraction: unsupported order");
    }
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_unpack_nodes(double *coefs, const double *stage, const int *gslotsAll, int nG, int world, int rowsPerRank, int shardB, int ncoef,
                         const double *normRows, double *gNorms, cudaStream_t st) {
    if (nG <= 0) return;
    unpack_nodes_kernel<<<world * rowsPerRank, 256, 0, st>>>(coefs, stage, gslotsAll, nG, world, rowsPerRank, shardB, ncoef, normRows, gNorms);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_pipe_reduce(const ApplyParams &P, const PipeBuffers &B, const int *gslots, double *gNorms, double *gNormsW, int nG,
                        cudaStream_t st, double *stageRows, int gBase) {
    const int nBlocks = nG * 8, blkBase = gBase * 8;
    if (nBlocks - blkBase <= 0) return;
    pipe_reduce_kernel<<<(nBlocks - blkBase + 7) / 8, 256, 0, st>>>(B, gslots, P.gCoefs, gNorms, gNormsW, blkBase, nBlocks, P.K * P.K * P.K,
                                                                     stageRows);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
