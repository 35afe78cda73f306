// Operator-application kernels (the hot loop of ConvolutionCalculator::calcNode,
// src/treebuilders/ConvolutionCalculator.cpp:224-382, and DerivativeCalculator::calcNode,
// src/treebuilders/DerivativeCalculator.cpp:115-275).
//
// One CTA per output node g (8 warps; warp w owns output component gt = w). The node's band is
// processed in batches:
//   phase A (all threads): every (input node, candidate term) pair of the batch is screened with the
//     reference's predicate. The integer band tests were hoisted into per-depth tables on the host
//     (DepthInfo); what remains is ((1.0*|O_x|)*|O_y|)*|O_z| * (bandSizeFactor*|f_ft|) > gThrs evaluated
//     in the reference's FP64 operation order (applyOperComp :277-288, applyOperator :297-329), for the
//     band-allowed (gt,ft) combinations only. Survivors are compacted, order preserving, into a
//     shared-memory item list (input slot, term, offset, 64-bit (gt,ft) mask).
//   phase B (per warp): for every item with bits for this warp's gt, and every ft, the three 1-D
//     contractions of tensorApplyOperComp (:333-382). A (g, gt) block is accumulated by exactly one
//     warp in a fixed order (input node, ft, term ascending), so results are run-to-run identical.
//
// k = 7 (K = 8): FP64 tensor-core path. An 8x8x8 block is eight 8x8 tiles; each contraction is
//   16 DMMA.8x8x4. Stages 1 and 2 chain register-to-register: the D fragment of stage 1 (row = lane/4,
//   columns 2q,2q+1 with q = lane%4) is exactly a B fragment of stage 2 when the free input index is
//   enumerated in the order sigma(c) = c/2 + 4*(c%2). Stage 3 contracts the tile index, which lives
//   inside a thread, so the 512 intermediates take one trip through a padded, conflict-free
//   warp-private shared-memory tile (8 x 16-byte stores + 8 x 16-byte loads per thread).
// other orders: generic FP64 FMA path through warp-private shared-memory scratch.
//
// Algorithmic work per surviving tuple: 6 K^4 flop (3 x 2 K^4); the source block
// (8 K^3 bytes) is reused across terms, the three operator blocks (3 x 8 K^2 bytes) come from L1/L2.
#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

constexpr double kMachineZero = 1.0e-14;
constexpr int kApplyThreads = 256;
constexpr int kCap = 1024;    // candidates / items per batch
constexpr int kMaxNbrBatch = 256;

struct Item {
    int fslot;
    int term;
    int code; // offset code inside the depth's band cube
    int pad;
    unsigned long long mask;
};

struct BatchSmem {
    Item items[kCap];
    int prefix[kMaxNbrBatch + 1]; // candidate prefix over the neighbours of the batch
    int warpTot[8];
    int nItems;
    int nNbrBatch;
    int total;
};

__device__ __forceinline__ void decode_delta(int code, int W, int d[3]) {
    const int cube = 2 * W + 1;
    d[0] = code % cube - W;
    d[1] = (code / cube) % cube - W;
    d[2] = code / (cube * cube) - W;
}

// Phase A for neighbours [p0, ...) of g. Returns the index of the first neighbour NOT processed.
__device__ int screen_batch(const ApplyParams &P, const GDesc &g, const DepthInfo di, int p0, BatchSmem &S,
                            unsigned long long &applied) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int *coff = P.candOff + di.cubeOff;
    // ---- candidate counts of up to 256 neighbours, block-wide inclusive scan
    const int p = p0 + tid;
    int cnt = 0;
    if (p < g.nbrCnt) {
        const int code = P.nbr[g.nbrOff + p].code;
        cnt = coff[code + 1] - coff[code];
    }
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) S.warpTot[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; w++) wbase += S.warpTot[w];
    incl += wbase;
    S.prefix[tid + 1] = incl;
    if (tid == 0) S.prefix[0] = 0;
    __syncthreads();
    // number of neighbours whose candidates fit into the batch
    if (tid == 0) {
        int n = 0;
        const int lim = min(kMaxNbrBatch, g.nbrCnt - p0);
        while (n < lim && S.prefix[n + 1] <= kCap) n++;
        S.nNbrBatch = n; // >= 1 because a single neighbour has at most M <= kCap candidates
        S.total = S.prefix[n];
        S.nItems = 0;
    }
    __syncthreads();
    const int nNb = S.nNbrBatch, total = S.total;
    const bool deriv = P.derivDir >= 0;
    // ---- screen candidates, order-preserving compaction
    for (int base = 0; base < total; base += kApplyThreads) {
        const int c = base + tid;
        unsigned long long pass = 0ull;
        Item it;
        it.fslot = 0;
        it.term = 0;
        it.code = 0;
        it.pad = 0;
        if (c < total) {
            // neighbour of candidate c: binary search in prefix
            int lo = 0, hi = nNb;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (S.prefix[mid] <= c) lo = mid;
                else hi = mid;
            }
            const NbrEntry nb = P.nbr[g.nbrOff + p0 + lo];
            const int e = coff[nb.code] + (c - S.prefix[lo]);
            const int term = P.candTerm[e];
            const unsigned long long band = P.candMask[e];
            it.fslot = nb.fslot;
            it.term = term;
            it.code = nb.code;
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
#pragma unroll
            for (int ft = 0; ft < 8; ft++)
                if (!(fn[ft] < kMachineZero)) fmask |= 0x0101010101010101ull << ft;
            unsigned long long todo = band & fmask;
            if (deriv) {
                pass = todo; // derivative apply has no norm screening (DerivativeCalculator.cpp:211-249)
            } else if (todo) {
                int d[3];
                decode_delta(nb.code, di.W, d);
                const int nbase = P.nodeBase[(size_t)term * P.DM + g.depth];
                double nr[3][4];
#pragma unroll
                for (int dd = 0; dd < 3; dd++) {
                    const double4 v = *reinterpret_cast<const double4 *>(P.onorms + (size_t)(nbase + d[dd]) * 4);
                    nr[dd][0] = v.x;
                    nr[dd][1] = v.y;
                    nr[dd][2] = v.z;
                    nr[dd][3] = v.w;
                }
                const int *bs = P.bsf + ((size_t)term * P.DM + g.depth) * 64;
                while (todo) {
                    const int b = __ffsll((long long)todo) - 1;
                    todo &= todo - 1;
                    const int gt = b >> 3, ft = b & 7;
                    const int c0 = 2 * (gt & 1) + (ft & 1);
                    const int c1 = 2 * ((gt >> 1) & 1) + ((ft >> 1) & 1);
                    const int c2 = 2 * ((gt >> 2) & 1) + ((ft >> 2) & 1);
                    double oNorm = 1.0;
                    oNorm *= nr[0][c0];
                    oNorm *= nr[1][c1];
                    oNorm *= nr[2][c2];
                    const double fThreshold = bs[b] * fn[ft];
                    const double upperBound = oNorm * fThreshold;
                    if (upperBound > P.gThrs) pass |= 1ull << b;
                }
            }
        }
        it.mask = pass;
        applied += __popcll(pass);
        const bool keep = pass != 0ull;
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        const int within = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) S.warpTot[warp] = __popc(bal);
        __syncthreads();
        int off = S.nItems;
        for (int w = 0; w < warp; w++) off += S.warpTot[w];
        if (keep) S.items[off + within] = it;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += S.warpTot[w];
            S.nItems += t;
        }
        __syncthreads();
    }
    return p0 + nNb;
}

// ------------------------------------------------------------------------------------------------
// generic order: warp-private scratch S1,S2 (K^3 doubles each), result accumulated into global g
__device__ __forceinline__ void contract_generic(int K, const double *__restrict__ f, const double *o0, const double *o1,
                                                 const double *o2, double *S1, double *S2, double *g, int lane) {
    const int K2 = K * K, Kd = K2 * K;
    const double *ops[3] = {o0, o1, o2};
    const double *in = f;
    double *outs[3] = {S1, S2, g};
    for (int st = 0; st < 3; st++) {
        const double *op = ops[st];
        double *out = outs[st];
        for (int o = lane; o < Kd; o += 32) {
            const int c = o / K2, r = o - c * K2;
            double acc;
            if (op) {
                acc = 0.0;
                const double *fr = in + K * r;
                const double *oc = op + K * c;
                for (int t = 0; t < K; t++) acc = fma(fr[t], oc[t], acc);
            } else {
                acc = in[c + K * r]; // identity: pure transpose
            }
            if (st == 2) out[o] += acc;
            else out[o] = acc;
        }
        __syncwarp();
        in = out;
    }
}

__global__ void __launch_bounds__(kApplyThreads) apply_generic_kernel(ApplyParams P, int NW) {
    extern __shared__ __align__(16) unsigned char smraw[];
    BatchSmem &S = *reinterpret_cast<BatchSmem *>(smraw);
    double *scratch = reinterpret_cast<double *>(smraw + ((sizeof(BatchSmem) + 15) & ~(size_t)15));
    const GDesc g = P.gdesc[blockIdx.x];
    const int K = P.K, K2 = K * K, Kd = K2 * K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool deriv = P.derivDir >= 0;

    double *gNode = (g.partial < 0) ? P.gCoefs + (size_t)g.slot * 8 * Kd : P.partials + (size_t)g.partial * 8 * Kd;
    for (int i = tid; i < 8 * Kd; i += kApplyThreads) gNode[i] = 0.0; // gNode.zeroCoefs()
    __syncthreads();
    if (g.nbrCnt == 0) return;
    const DepthInfo di = P.depthInfo[g.depth];

    unsigned long long applied = 0;
    int p0 = 0;
    while (p0 < g.nbrCnt) {
        const int p1 = screen_batch(P, g, di, p0, S, applied);
        const int nItems = S.nItems;
        if (warp < NW) {
            double *S1 = scratch + (size_t)warp * 2 * Kd;
            double *S2 = S1 + Kd;
            for (int gt = warp; gt < 8; gt += NW) {
                int i0 = 0;
                while (i0 < nItems) {
                    const int fslot = S.items[i0].fslot;
                    int i1 = i0 + 1;
                    while (i1 < nItems && S.items[i1].fslot == fslot) i1++;
                    const bool gen = fslot >= P.nRealF;
                    for (int ft = 0; ft < (gen ? 1 : 8); ft++) {
                        const double *fblk =
                            gen ? P.fGen + (size_t)(fslot - P.nRealF) * Kd : P.fReal + ((size_t)fslot * 8 + ft) * Kd;
                        const int c0 = 2 * (gt & 1) + (ft & 1);
                        const int c1 = 2 * ((gt >> 1) & 1) + ((ft >> 1) & 1);
                        const int c2 = 2 * ((gt >> 2) & 1) + ((ft >> 2) & 1);
                        for (int i = i0; i < i1; i++) {
                            if (!((S.items[i].mask >> (gt * 8 + ft)) & 1ull)) continue;
                            int d[3];
                            decode_delta(S.items[i].code, di.W, d);
                            const int nbase = P.nodeBase[(size_t)S.items[i].term * P.DM + g.depth];
                            const double *o0 = P.mats + ((size_t)(nbase + d[0]) * 4 + c0) * K2;
                            const double *o1 = P.mats + ((size_t)(nbase + d[1]) * 4 + c1) * K2;
                            const double *o2 = P.mats + ((size_t)(nbase + d[2]) * 4 + c2) * K2;
                            if (deriv) {
                                if (P.derivDir != 0) o0 = nullptr;
                                if (P.derivDir != 1) o1 = nullptr;
                                if (P.derivDir != 2) o2 = nullptr;
                            }
                            contract_generic(K, fblk, o0, o1, o2, S1, S2, gNode + (size_t)gt * Kd, lane);
                        }
                    }
                    i0 = i1;
                }
            }
        }
        __syncthreads();
        p0 = p1;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, off);
    if (lane == 0 && applied) atomicAdd(P.counters, applied);
}

// ------------------------------------------------------------------------------------------------
// K = 8: FP64 tensor cores
constexpr int kTileSi = 18;   // i2 stride (doubles): 144 B == 16 B mod 128 B
constexpr int kTileSm = 152;  // m1 stride (doubles): 1216 B == 64 B mod 128 B
constexpr int kTileDoubles = 8 * kTileSm;

__global__ void __launch_bounds__(kApplyThreads, 1) apply_dmma8_kernel(ApplyParams P) {
    extern __shared__ __align__(16) unsigned char smraw[];
    constexpr int K2 = 64, Kd = 512;
    double *tiles = reinterpret_cast<double *>(smraw); // 8 warps x kTileDoubles
    BatchSmem &S = *reinterpret_cast<BatchSmem *>(smraw + (size_t)8 * kTileDoubles * sizeof(double));
    const GDesc g = P.gdesc[blockIdx.x];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int sig = (r >> 1) + 4 * (r & 1); // sigma(r)
    double *T = tiles + warp * kTileDoubles;
    const int gt = warp;
    double *gblk = ((g.partial < 0) ? P.gCoefs + (size_t)g.slot * 8 * Kd : P.partials + (size_t)g.partial * 8 * Kd) + (size_t)gt * Kd;

    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; t++) acc[t][0] = acc[t][1] = 0.0;

    unsigned long long applied = 0;
    if (g.nbrCnt > 0) {
        const DepthInfo di = P.depthInfo[g.depth];
        int p0 = 0;
        while (p0 < g.nbrCnt) {
            const int p1 = screen_batch(P, g, di, p0, S, applied);
            const int nItems = S.nItems;
            int i0 = 0;
            while (i0 < nItems) {
                const int fslot = S.items[i0].fslot;
                int i1 = i0 + 1;
                unsigned long long runMask = S.items[i0].mask;
                while (i1 < nItems && S.items[i1].fslot == fslot) {
                    runMask |= S.items[i1].mask;
                    i1++;
                }
                const uint32_t myFt = (uint32_t)((runMask >> (gt * 8)) & 0xffull);
                const bool gen = fslot >= P.nRealF;
                for (int ft = 0; ft < 8; ft++) {
                    if (!((myFt >> ft) & 1u)) continue;
                    const double *fblk =
                        gen ? P.fGen + (size_t)(fslot - P.nRealF) * Kd : P.fReal + ((size_t)fslot * 8 + ft) * Kd;
                    // B fragments of stage 1: f[i0 = q+4s, i1 = sigma(r), i2 = j]
                    double bf[8][2];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        bf[j][0] = __ldg(fblk + q + 8 * sig + 64 * j);
                        bf[j][1] = __ldg(fblk + q + 4 + 8 * sig + 64 * j);
                    }
                    const int c0 = 2 * (gt & 1) + (ft & 1);
                    const int c1 = 2 * ((gt >> 1) & 1) + ((ft >> 1) & 1);
                    const int c2 = 2 * ((gt >> 2) & 1) + ((ft >> 2) & 1);
                    for (int i = i0; i < i1; i++) {
                        if (!((S.items[i].mask >> (gt * 8 + ft)) & 1ull)) continue;
                        int d[3];
                        decode_delta(S.items[i].code, di.W, d);
                        const int nbase = P.nodeBase[(size_t)S.items[i].term * P.DM + g.depth];
                        const double *o0 = P.mats + ((size_t)(nbase + d[0]) * 4 + c0) * K2;
                        const double *o1 = P.mats + ((size_t)(nbase + d[1]) * 4 + c1) * K2;
                        const double *o2 = P.mats + ((size_t)(nbase + d[2]) * 4 + c2) * K2;
                        // operator fragments: element (q + 4s) + 8 r of each block
                        const double a00 = __ldg(o0 + q + 8 * r), a01 = __ldg(o0 + q + 4 + 8 * r);
                        const double a10 = __ldg(o1 + q + 8 * r), a11 = __ldg(o1 + q + 4 + 8 * r);
                        const double b20 = __ldg(o2 + q + 8 * r), b21 = __ldg(o2 + q + 4 + 8 * r);
                        // stage 1: X1[m0=r][i1=q+4e][i2=j]
                        double d1[8][2];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            d1[j][0] = d1[j][1] = 0.0;
                            dmma884(d1[j][0], d1[j][1], a00, bf[j][0]);
                            dmma884(d1[j][0], d1[j][1], a01, bf[j][1]);
                        }
                        // stage 2: X2[m1=r][m0=2q+e][i2=j]
                        double d2[8][2];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            d2[j][0] = d2[j][1] = 0.0;
                            dmma884(d2[j][0], d2[j][1], a10, d1[j][0]);
                            dmma884(d2[j][0], d2[j][1], a11, d1[j][1]);
                        }
                        // transpose through the warp-private tile: T[m1][i2][m0]
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            *reinterpret_cast<double2 *>(T + 2 * q + kTileSi * j + kTileSm * r) =
                                make_double2(d2[j][0], d2[j][1]);
                        __syncwarp();
                        // stage 3: g[m0=t][m1=r][m2=2q+e] += sum_i2 X2 * O2[i2][m2]
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const double2 x0 = *reinterpret_cast<const double2 *>(T + 2 * u + kTileSi * q + kTileSm * r);
                            const double2 x1 =
                                *reinterpret_cast<const double2 *>(T + 2 * u + kTileSi * (q + 4) + kTileSm * r);
                            dmma884(acc[2 * u][0], acc[2 * u][1], x0.x, b20);
                            dmma884(acc[2 * u][0], acc[2 * u][1], x1.x, b21);
                            dmma884(acc[2 * u + 1][0], acc[2 * u + 1][1], x0.y, b20);
                            dmma884(acc[2 * u + 1][0], acc[2 * u + 1][1], x1.y, b21);
                        }
                        __syncwarp();
                    }
                }
                i0 = i1;
            }
            __syncthreads();
            p0 = p1;
        }
    }
    // g block gt: element m0 + 8 m1 + 64 m2
#pragma unroll
    for (int t = 0; t < 8; t++) {
        gblk[t + 8 * r + 64 * (2 * q)] = acc[t][0];
        gblk[t + 8 * r + 64 * (2 * q + 1)] = acc[t][1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) applied += __shfl_xor_sync(0xffffffffu, applied, off);
    if (lane == 0 && applied) atomicAdd(P.counters, applied);
}

template <typename Kern> void ensure_smem(Kern kern, size_t bytes, size_t &configured) {
    if (bytes > configured) {
        MRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
}

} // namespace

void launch_apply(const ApplyParams &P, int nG, cudaStream_t st) {
    if (nG <= 0) return;
    if (P.M > kCap) MRX_ABORT("separation rank exceeds the screening batch capacity");
    const bool deriv = P.derivDir >= 0;
    const char *force = getenv("MRX_FORCE_GENERIC");
    if (P.K == 8 && !deriv && !(force && force[0] == '1')) {
        size_t bytes = (size_t)8 * kTileDoubles * 8 + sizeof(BatchSmem);
        static size_t conf = 0;
        ensure_smem(apply_dmma8_kernel, bytes, conf);
        apply_dmma8_kernel<<<nG, kApplyThreads, bytes, st>>>(P);
    } else {
        const int Kd = P.K * P.K * P.K;
        size_t head = (sizeof(BatchSmem) + 15) & ~(size_t)15;
        int NW = 8;
        while (NW > 1 && head + (size_t)NW * 2 * Kd * 8 > 200 * 1024) NW >>= 1;
        size_t bytes = head + (size_t)NW * 2 * Kd * 8;
        static size_t conf = 0;
        ensure_smem(apply_generic_kernel, bytes, conf);
        apply_generic_kernel<<<nG, kApplyThreads, bytes, st>>>(P, NW);
    }
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
