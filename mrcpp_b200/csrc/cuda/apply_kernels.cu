// Operator-application kernels (the hot loop of ConvolutionCalculator::calcNode,
// src/treebuilders/ConvolutionCalculator.cpp:224-382, and DerivativeCalculator::calcNode,
// src/treebuilders/DerivativeCalculator.cpp:115-275).
//
// One CTA per output node g. For every input node f of g's band:
//   phase A (all threads, one thread per separation term): the reference's screening, bit for bit --
//     per-term max-width test, per-dimension band test per component, then
//     ((1.0*|O_x|)*|O_y|)*|O_z| * (bandSizeFactor*|f_ft|) > gThrs in the same FP64 operation order
//     (applyOperComp :277-288, applyOperator :297-329). The outcome is a (gt,ft) x term bit matrix in
//     shared memory, built with warp ballots.
//   phase B (warp w owns output component gt = w): for every surviving (ft, term) the three 1-D
//     contractions of tensorApplyOperComp (:333-382). The accumulator of a (g, gt) block is owned by
//     exactly one warp, so the summation order is fixed (f ascending, ft ascending, term ascending).
//
// k = 7 (K = 8): FP64 tensor-core path. An 8x8x8 block is eight 8x8 tiles; each contraction is
//   16 DMMA.8x8x4. Stage 1 and 2 chain register-to-register: the D fragment of stage 1 (row = lane/4,
//   columns 2q,2q+1 with q = lane%4) is exactly a B fragment of stage 2 when the free input index is
//   enumerated in the order sigma(c) = c/2 + 4*(c%2). Stage 3 contracts the tile index, which lives
//   inside a thread, so the 512 intermediates take one trip through a padded, conflict-free
//   warp-private shared-memory tile (8 x 16-byte stores + 8 x 16-byte loads per thread).
// other orders: generic FP64 FMA path through warp-private shared-memory scratch.
//
// Algorithmic work per surviving tuple: 6 K^4 flop (3 x 2 K^4); the source block
// (8 K^3 bytes) is reused by up to 8 M tuples, the three operator blocks (3 x 8 K^2 bytes) come from L1/L2.
#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

constexpr double kMachineZero = 1.0e-14;
constexpr int kApplyThreads = 256;

struct NbrInfo {
    int fslot;
    int d[3];
};

__device__ __forceinline__ NbrInfo decode_nbr(const ApplyParams &P, const GDesc &g, int p) {
    NbrInfo n;
    n.fslot = P.nbr[g.nbrOff + p];
    int x = p % g.nb[0];
    int yz = p / g.nb[0];
    int y = yz % g.nb[1];
    int z = yz / g.nb[1];
    n.d[0] = g.s[0] + x - g.l[0];
    n.d[1] = g.s[1] + y - g.l[1];
    n.d[2] = g.s[2] + z - g.l[2];
    return n;
}

// Phase A. termMask[c * MW + w] = bit t set iff tuple (term 32w+t, gt = c/8, ft = c%8) passes.
// nodeIdx[3*term + d] = global operator-node index for translation d (valid when any bit of the term is set).
__device__ void screen_neighbour(const ApplyParams &P, const GDesc &g, const NbrInfo &nb, uint32_t *termMask, int *nodeIdx,
                                 int MW, bool deriv) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthreads = blockDim.x;
    double fn[8];
    if (nb.fslot < P.nRealF) {
#pragma unroll
        for (int ft = 0; ft < 8; ft++) fn[ft] = P.fNorms[(size_t)nb.fslot * 8 + ft];
    } else {
        fn[0] = P.fGenNorms[nb.fslot - P.nRealF];
#pragma unroll
        for (int ft = 1; ft < 8; ft++) fn[ft] = 0.0; // generated nodes carry scaling only (MWNode.cpp:644)
    }
    const int a0 = abs(nb.d[0]), a1 = abs(nb.d[1]), a2 = abs(nb.d[2]);
    const int maxDelta = max(a0, max(a1, a2));
    const bool depth0 = (g.depth == 0);
    for (int base = 0; base < P.M; base += nthreads) {
        const int term = base + tid;
        unsigned long long bits = 0ull;
        if (term < P.M) {
            const int *bwp = P.bw + ((size_t)term * P.DM + g.depth) * 5;
            const int mt = P.maxTransl[(size_t)term * P.DM + g.depth];
            if (maxDelta <= bwp[4] && maxDelta <= mt) {
                const int off = P.nodeOff[(size_t)term * P.DM + g.depth];
                int nidx[3];
                double nr[3][4];
                bool ib[3][4];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    nidx[d] = off + nb.d[d] + mt;
                    nodeIdx[3 * term + d] = nidx[d];
                    const int ad = abs(nb.d[d]);
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        ib[d][c] = ad <= bwp[c];
                        nr[d][c] = ib[d][c] ? P.onorms[(size_t)nidx[d] * 4 + c] : 0.0;
                    }
                }
                const int *bs = P.bsf + ((size_t)term * P.DM + g.depth) * 64;
                for (int gt = 0; gt < 8; gt++) {
#pragma unroll
                    for (int ft = 0; ft < 8; ft++) {
                        if (fn[ft] < kMachineZero) continue;
                        const int c0 = 2 * (gt & 1) + (ft & 1);
                        const int c1 = 2 * ((gt >> 1) & 1) + ((ft >> 1) & 1);
                        const int c2 = 2 * ((gt >> 2) & 1) + ((ft >> 2) & 1);
                        if (!(ib[0][c0] && ib[1][c1] && ib[2][c2])) continue;
                        bool pass;
                        if (!deriv) {
                            if (!depth0 && gt == 0 && ft == 0) continue; // T block only at operator depth 0 (:261)
                            double oNorm = 1.0;
                            oNorm *= nr[0][c0];
                            oNorm *= nr[1][c1];
                            oNorm *= nr[2][c2];
                            const double fThreshold = bs[gt * 8 + ft] * fn[ft];
                            const double upperBound = oNorm * fThreshold;
                            pass = upperBound > P.gThrs;
                        } else {
                            // DerivativeCalculator::applyOperator (:211-249): operator only along derivDir,
                            // identity (same node, T or A component) in the other directions
                            pass = true;
#pragma unroll
                            for (int d = 0; d < 3; d++) {
                                if (d == P.derivDir) continue;
                                const int cd = (d == 0) ? c0 : (d == 1 ? c1 : c2);
                                if (!(nb.d[d] == 0 && (cd == 0 || cd == 3))) pass = false;
                            }
                        }
                        if (pass) bits |= 1ull << (gt * 8 + ft);
                    }
                }
            }
        }
        if (base + warp * 32 < P.M) {
            const int w = (base >> 5) + warp;
            for (int c = 0; c < 64; c++) {
                uint32_t word = __ballot_sync(0xffffffffu, (bits >> c) & 1ull);
                if (lane == 0) termMask[c * MW + w] = word;
            }
        }
    }
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

template <bool DERIV>
__global__ void __launch_bounds__(kApplyThreads) apply_generic_kernel(ApplyParams P, int MW, int NW) {
    extern __shared__ unsigned char smraw[];
    const GDesc g = P.gdesc[blockIdx.x];
    const int K = P.K, K2 = K * K, Kd = K2 * K;
    uint32_t *termMask = reinterpret_cast<uint32_t *>(smraw);
    int *nodeIdx = reinterpret_cast<int *>(termMask + 64 * MW);
    size_t off = (size_t)(64 * MW + 3 * P.M) * 4;
    off = (off + 15) & ~(size_t)15;
    double *scratch = reinterpret_cast<double *>(smraw + off);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    double *gNode = P.gCoefs + (size_t)g.slot * 8 * Kd;
    for (int i = tid; i < 8 * Kd; i += kApplyThreads) gNode[i] = 0.0; // gNode.zeroCoefs()
    __syncthreads();

    unsigned long long applied = 0;
    const int nNbr = g.nb[0] * g.nb[1] * g.nb[2];
    for (int p = 0; p < nNbr; p++) {
        const NbrInfo nb = decode_nbr(P, g, p);
        screen_neighbour(P, g, nb, termMask, nodeIdx, MW, DERIV);
        __syncthreads();
        if (warp < NW) {
            double *S1 = scratch + (size_t)warp * 2 * Kd;
            double *S2 = S1 + Kd;
            const bool gen = nb.fslot >= P.nRealF;
            for (int gt = warp; gt < 8; gt += NW) {
                for (int ft = 0; ft < (gen ? 1 : 8); ft++) {
                    const double *fblk = gen ? P.fGen + (size_t)(nb.fslot - P.nRealF) * Kd
                                             : P.fReal + ((size_t)nb.fslot * 8 + ft) * Kd;
                    const int c0 = 2 * (gt & 1) + (ft & 1);
                    const int c1 = 2 * ((gt >> 1) & 1) + ((ft >> 1) & 1);
                    const int c2 = 2 * ((gt >> 2) & 1) + ((ft >> 2) & 1);
                    for (int w = 0; w < MW; w++) {
                        uint32_t word = termMask[(gt * 8 + ft) * MW + w];
                        while (word) {
                            const int b = __ffs(word) - 1;
                            word &= word - 1;
                            const int term = w * 32 + b;
                            const double *o0 = P.mats + ((size_t)nodeIdx[3 * term + 0] * 4 + c0) * K2;
                            const double *o1 = P.mats + ((size_t)nodeIdx[3 * term + 1] * 4 + c1) * K2;
                            const double *o2 = P.mats + ((size_t)nodeIdx[3 * term + 2] * 4 + c2) * K2;
                            if (DERIV) {
                                if (P.derivDir != 0) o0 = nullptr;
                                if (P.derivDir != 1) o1 = nullptr;
                                if (P.derivDir != 2) o2 = nullptr;
                            }
                            contract_generic(K, fblk, o0, o1, o2, S1, S2, gNode + (size_t)gt * Kd, lane);
                            applied++;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (lane == 0 && applied) atomicAdd(P.counters, applied);
}

// ------------------------------------------------------------------------------------------------
// K = 8: FP64 tensor cores
constexpr int kTileSi = 18;   // i2 stride (doubles): 144 B == 16 B mod 128 B
constexpr int kTileSm = 152;  // m1 stride (doubles): 1216 B == 64 B mod 128 B
constexpr int kTileDoubles = 8 * kTileSm;

__global__ void __launch_bounds__(kApplyThreads, 1) apply_dmma8_kernel(ApplyParams P, int MW) {
    extern __shared__ unsigned char smraw[];
    const GDesc g = P.gdesc[blockIdx.x];
    constexpr int K2 = 64, Kd = 512;
    double *tiles = reinterpret_cast<double *>(smraw); // 8 warps x kTileDoubles
    uint32_t *termMask = reinterpret_cast<uint32_t *>(tiles + 8 * kTileDoubles);
    int *nodeIdx = reinterpret_cast<int *>(termMask + 64 * MW);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    const int sig = (r >> 1) + 4 * (r & 1); // sigma(r)
    double *T = tiles + warp * kTileDoubles;
    const int gt = warp;

    double acc[8][2];
#pragma unroll
    for (int t = 0; t < 8; t++) acc[t][0] = acc[t][1] = 0.0;

    unsigned long long applied = 0;
    const int nNbr = g.nb[0] * g.nb[1] * g.nb[2];
    for (int p = 0; p < nNbr; p++) {
        const NbrInfo nb = decode_nbr(P, g, p);
        screen_neighbour(P, g, nb, termMask, nodeIdx, MW, false);
        __syncthreads();
        const bool gen = nb.fslot >= P.nRealF;
        for (int ft = 0; ft < (gen ? 1 : 8); ft++) {
            // any work for (gt, ft)?
            uint32_t any = 0;
            for (int w = 0; w < MW; w++) any |= termMask[(gt * 8 + ft) * MW + w];
            if (!any) continue;
            const double *fblk =
                gen ? P.fGen + (size_t)(nb.fslot - P.nRealF) * Kd : P.fReal + ((size_t)nb.fslot * 8 + ft) * Kd;
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
            for (int w = 0; w < MW; w++) {
                uint32_t word = termMask[(gt * 8 + ft) * MW + w];
                while (word) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1;
                    const int term = w * 32 + b;
                    const double *o0 = P.mats + ((size_t)nodeIdx[3 * term + 0] * 4 + c0) * K2;
                    const double *o1 = P.mats + ((size_t)nodeIdx[3 * term + 1] * 4 + c1) * K2;
                    const double *o2 = P.mats + ((size_t)nodeIdx[3 * term + 2] * 4 + c2) * K2;
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
                        *reinterpret_cast<double2 *>(T + 2 * q + kTileSi * j + kTileSm * r) = make_double2(d2[j][0], d2[j][1]);
                    __syncwarp();
                    // stage 3: g[m0=t][m1=r][m2=2q+e] += sum_i2 X2 * O2[i2][m2]
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double2 x0 = *reinterpret_cast<const double2 *>(T + 2 * u + kTileSi * q + kTileSm * r);
                        const double2 x1 = *reinterpret_cast<const double2 *>(T + 2 * u + kTileSi * (q + 4) + kTileSm * r);
                        dmma884(acc[2 * u][0], acc[2 * u][1], x0.x, b20);
                        dmma884(acc[2 * u][0], acc[2 * u][1], x1.x, b21);
                        dmma884(acc[2 * u + 1][0], acc[2 * u + 1][1], x0.y, b20);
                        dmma884(acc[2 * u + 1][0], acc[2 * u + 1][1], x1.y, b21);
                    }
                    __syncwarp();
                    applied++;
                }
            }
        }
        __syncthreads();
    }
    // g block gt: element m0 + 8 m1 + 64 m2
    double *gblk = P.gCoefs + ((size_t)g.slot * 8 + gt) * Kd;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        gblk[t + 8 * r + 64 * (2 * q)] = acc[t][0];
        gblk[t + 8 * r + 64 * (2 * q + 1)] = acc[t][1];
    }
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
    const int MW = (P.M + 31) / 32;
    const bool deriv = P.derivDir >= 0;
    const char *force = getenv("MRX_FORCE_GENERIC");
    if (P.K == 8 && !deriv && !(force && force[0] == '1')) {
        size_t bytes = (size_t)8 * kTileDoubles * 8 + (size_t)(64 * MW + 3 * P.M) * 4;
        static size_t conf = 0;
        ensure_smem(apply_dmma8_kernel, bytes, conf);
        apply_dmma8_kernel<<<nG, kApplyThreads, bytes, st>>>(P, MW);
    } else {
        const int Kd = P.K * P.K * P.K;
        size_t head = ((size_t)(64 * MW + 3 * P.M) * 4 + 15) & ~(size_t)15;
        int NW = 8;
        while (NW > 1 && head + (size_t)NW * 2 * Kd * 8 > 200 * 1024) NW >>= 1;
        size_t bytes = head + (size_t)NW * 2 * Kd * 8;
        if (deriv) {
            static size_t conf = 0;
            ensure_smem(apply_generic_kernel<true>, bytes, conf);
            apply_generic_kernel<true><<<nG, kApplyThreads, bytes, st>>>(P, MW, NW);
        } else {
            static size_t conf = 0;
            ensure_smem(apply_generic_kernel<false>, bytes, conf);
            apply_generic_kernel<false><<<nG, kApplyThreads, bytes, st>>>(P, MW, NW);
        }
    }
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
