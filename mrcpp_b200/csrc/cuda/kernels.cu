// Tree-maintenance kernels: two-scale filter passes (generic order), generated children, norms, dot.
//
// One CTA per parent node. A node is 8 blocks of K^3 doubles; each of the three passes applies the
// 2K x 2K two-scale filter along one dimension (math_utils::apply_filter, math_utils.cpp:175-194:
// out(K^2 x K) (+)= in(K x K^2)^T F), contracting the fastest index and making it the slowest, so
// after three passes the layout is restored. Blocks are staged in shared memory with one pad word
// per K elements so that the strided reads of a pass are bank-conflict free.
//
// Algorithmic traffic per node: read 8 K^3 + write 8 K^3 doubles = 128 K^3 bytes.
#include "../engine.hpp"
#include "common.cuh"
#include "kernels.cuh"

namespace mrx {

namespace {

constexpr int kTransformThreads = 256;

// MODE 0: TopDown (parent -> children scaling, = or +=); 1: BottomUp (children scaling -> parent);
// MODE 2: generated children of an input-tree node (scaling only, separate pool) + their norms;
// MODE 3: in-node compression (MWNode::mwTransform(Compression), MWNode.cpp:557-594): the node's own 8 blocks hold the
//         scaling blocks of its children (projection, ProjectionCalculator.cpp:34-51) and are replaced by (s, d).
// MODE 4: in-node reconstruction (MWNode::mwTransform(Reconstruction)): (s, d) replaced by the children's scaling blocks.
template <int MODE>
__global__ void __launch_bounds__(kTransformThreads)
transform_kernel(double *__restrict__ coefs, const double *__restrict__ realCoefs, double *__restrict__ genCoefs,
                 double *__restrict__ genNorms, int nReal, const int *__restrict__ pairs, int K, int padOn,
                 const double *__restrict__ filters, int overwrite) {
    extern __shared__ double sm[];
    const int K2 = K * K, Kd = K2 * K, ncoef = 8 * Kd;
    const int KdP = padOn ? (Kd + K2) : Kd;
    const int rowS = padOn ? (K + 1) : K;
    double *A = sm;
    double *B = sm + 8 * KdP;
    double *F = B + 8 * KdP;
    const int parent = pairs[2 * blockIdx.x];
    const int child0 = pairs[2 * blockIdx.x + 1];
    const int tid = threadIdx.x;

    const int op = (MODE == 1 || MODE == 3) ? 0 : 1; // Compression : Reconstruction
    for (int i = tid; i < 4 * K2; i += kTransformThreads) F[i] = filters[(size_t)op * 4 * K2 + i];

    for (int o = tid; o < 8 * Kd; o += kTransformThreads) {
        int t = o / Kd, rem = o - t * Kd;
        double v;
        if (MODE == 0 || MODE == 3 || MODE == 4) {
            v = coefs[(size_t)parent * ncoef + o];
        } else if (MODE == 1) {
            v = coefs[(size_t)(child0 + t) * ncoef + rem];
        } else {
            if (parent < nReal) v = realCoefs[(size_t)parent * ncoef + o];
            else v = (t == 0) ? genCoefs[(size_t)(parent - nReal) * Kd + rem] : 0.0;
        }
        A[t * KdP + (padOn ? rem + rem / K : rem)] = v;
    }
    __syncthreads();

    double *in = A, *out = B;
    for (int pass = 0; pass < 3; pass++) {
        for (int o = tid; o < 8 * Kd; o += kTransformThreads) {
            int gt = o / Kd, rem = o - gt * Kd;
            int j = rem / K2, m = rem - j * K2;
            int gbit = (gt >> pass) & 1;
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < 2; b++) {
                int ft = (gt & ~(1 << pass)) | (b << pass);
                const double *inb = in + ft * KdP + m * rowS;
                const double *Fm = F + (2 * gbit + b) * K2 + j;
                for (int t = 0; t < K; t++) acc = fma(inb[t], Fm[t * K], acc);
            }
            if (pass < 2) {
                out[gt * KdP + (padOn ? rem + rem / K : rem)] = acc;
            } else if (MODE == 0) {
                double *dst = coefs + (size_t)(child0 + gt) * ncoef + rem;
                if (overwrite) *dst = acc;
                else *dst += acc;
            } else if (MODE == 1 || MODE == 3 || MODE == 4) {
                coefs[(size_t)parent * ncoef + o] = acc;
            } else {
                genCoefs[(size_t)(child0 - nReal + gt) * Kd + rem] = acc;
                out[gt * KdP + (padOn ? rem + rem / K : rem)] = acc;
            }
        }
        __syncthreads();
        double *tmp = in;
        in = out;
        out = tmp;
    }
    if (MODE == 0 && overwrite) {
        // giveChildrenCoefs(overwrite=true) zeroes the children first (MWNode.cpp:317-319)
        for (int o = tid; o < 8 * 7 * Kd; o += kTransformThreads) {
            int c = o / (7 * Kd), rem = o - c * 7 * Kd;
            coefs[(size_t)(child0 + c) * ncoef + Kd + rem] = 0.0;
        }
    }
    if (MODE == 2) {
        // norms of the 8 generated scaling blocks: warp w reduces child w (`in` holds the last pass)
        int w = tid >> 5, lane = tid & 31;
        double s = 0.0;
        for (int e = lane; e < Kd; e += 32) {
            double v = in[w * KdP + (padOn ? e + e / K : e)];
            s = fma(v, v, s);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) genNorms[child0 - nReal + w] = sqrt(s);
    }
}

// ------------------------------------------------------------------------------------------------
// k = 7 (K = 8): two-scale transform on the FP64 tensor cores. One CTA (4 warps) per parent node.
// The three filter passes are the same algebra as the three contractions of the operator application
// (apply_pipeline.cu): pass p contracts dimension p with the 2K x 2K filter, i.e. for output block gt
//   out_gt = sum_{b} F[2 gbit + b]^T . in_{ft(b)}        (math_utils::apply_filter, math_utils.cpp:175-194)
// Warp w = (g0, f2) runs passes 0 and 1 register-to-register (D fragment -> B fragment under sigma) for
// its four input blocks (f0, f1), writes P1(g0, g1, f2) into a padded shared tile, and after one barrier
// warp w = (g0, g1) contracts z for g2 = 0, 1 out of the tiles. 768 DMMA.8x8x4 per node = 96 K^4 flop.
// padded tile of the z pass: element (m1 = r, i2 = j, m0 pair) at 2 q + kT8Si j + kT8Sm r. kT8Si / 2 = 5 (mod 8) and
// kT8Sm / 2 = 4 (mod 8) keep the 16-byte stores of a quarter warp and both 16-byte fragment reads of the z pass on eight
// distinct bank groups with only 2 pad doubles per row of 8 (704 doubles per tile instead of 1216).
constexpr int kT8Si = 10, kT8Sm = 88, kT8Doubles = 8 * kT8Sm;
constexpr size_t kT8Bytes = (size_t)8 * kT8Doubles * sizeof(double); // 45 KB: the TMA-staged node (32 KB) aliases the tiles

template <int MODE> // 0: TopDown (parent 8 blocks -> children scaling, = or +=); 1: BottomUp (children scaling -> parent); 3: in-node compression
__global__ void __launch_bounds__(128) transform8_kernel(double *__restrict__ coefs, const int *__restrict__ pairs,
                                                         const double *__restrict__ filters, int overwrite, double *__restrict__ norms) {
    extern __shared__ __align__(128) double tiles8[]; // 8 padded tiles (kT8Doubles each); the TMA-staged node (8 x 512) lives in the same bytes until its fragments are in registers
    constexpr int Kd = 512, ncoef = 8 * Kd;
    const int parent = pairs[2 * blockIdx.x];
    const int child0 = pairs[2 * blockIdx.x + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = lane >> 2, q = lane & 3;
    const int sig = (r >> 1) + 4 * (r & 1);
    const int bo = q + 8 * sig;
    // filter fragments: F[op][2 gbit + b][t * 8 + j] -> element t = q + 4 s, j = r (A operand of passes 0/1, B operand of pass 2)
    const double *F = filters + (size_t)((MODE == 1 || MODE == 3) ? 0 : 1) * 4 * 64;
    double fa[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        fa[i][0] = F[i * 64 + q * 8 + r];
        fa[i][1] = F[i * 64 + (q + 4) * 8 + r];
    }
    // ---- the node's 8 source blocks (32 KB) are staged by the TMA engine: one bulk copy for a parent node (contiguous),
    //      eight 4 KB copies for the scaling blocks of the children; all bytes in flight at once
    double *inbuf = tiles8;
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(ncoef * sizeof(double)));
        if (MODE == 0 || MODE == 3) {
            bulk_g2s(inbuf, coefs + (size_t)parent * ncoef, (uint32_t)(ncoef * sizeof(double)), &bar);
        } else {
#pragma unroll
            for (int ft = 0; ft < 8; ft++)
                bulk_g2s(inbuf + ft * Kd, coefs + (size_t)(child0 + ft) * ncoef, (uint32_t)(Kd * sizeof(double)), &bar);
        }
    }
    mbar_wait(&bar, 0);
    {
        const int g0 = warp & 1, f2 = warp >> 1;
        // filter pair of pass 0 for this warp's g0 (selected once: keeps the fragment table in registers)
        double fg[2][2];
#pragma unroll
        for (int f0 = 0; f0 < 2; f0++) {
            fg[f0][0] = g0 ? fa[2 + f0][0] : fa[f0][0];
            fg[f0][1] = g0 ? fa[2 + f0][1] : fa[f0][1];
        }
        // source fragments out of the TMA-staged node: f[i0 = q + 4 s, i1 = sigma(r), i2 = j]
        double bf[4][8][2];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const double *blk = inbuf + ((b & 1) | ((b >> 1) << 1) | (f2 << 2)) * Kd;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                bf[b][j][0] = blk[bo + 64 * j];
                bf[b][j][1] = blk[bo + 4 + 64 * j];
            }
        }
        __syncthreads(); // every warp holds its source fragments: the staging bytes become the tiles
        double p0[2][8][2];
#pragma unroll
        for (int f1 = 0; f1 < 2; f1++) {
#pragma unroll
            for (int j = 0; j < 8; j++) p0[f1][j][0] = p0[f1][j][1] = 0.0;
#pragma unroll
            for (int f0 = 0; f0 < 2; f0++) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    dmma884(p0[f1][j][0], p0[f1][j][1], fg[f0][0], bf[f0 | (f1 << 1)][j][0]);
                    dmma884(p0[f1][j][0], p0[f1][j][1], fg[f0][1], bf[f0 | (f1 << 1)][j][1]);
                }
            }
        }
#pragma unroll
        for (int g1 = 0; g1 < 2; g1++) {
            double *T = tiles8 + (g0 | (g1 << 1) | (f2 << 2)) * kT8Doubles;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int f1 = 0; f1 < 2; f1++) {
                    dmma884(d0, d1, fa[2 * g1 + f1][0], p0[f1][j][0]);
                    dmma884(d0, d1, fa[2 * g1 + f1][1], p0[f1][j][1]);
                }
                *reinterpret_cast<double2 *>(T + 2 * q + kT8Si * j + kT8Sm * r) = make_double2(d0, d1);
            }
        }
    }
    __syncthreads();
    {
        const int g01 = warp; // (g0, g1)
#pragma unroll
        for (int g2 = 0; g2 < 2; g2++) {
            double acc[8][2];
#pragma unroll
            for (int t = 0; t < 8; t++) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
            for (int f2 = 0; f2 < 2; f2++) {
                const double *T = tiles8 + (g01 | (f2 << 2)) * kT8Doubles;
                const double b0 = fa[2 * g2 + f2][0], b1 = fa[2 * g2 + f2][1];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    const double2 x0 = *reinterpret_cast<const double2 *>(T + 2 * v + kT8Si * q + kT8Sm * r);
                    const double2 x1 = *reinterpret_cast<const double2 *>(T + 2 * v + kT8Si * (q + 4) + kT8Sm * r);
                    dmma884(acc[2 * v][0], acc[2 * v][1], x0.x, b0);
                    dmma884(acc[2 * v][0], acc[2 * v][1], x1.x, b1);
                    dmma884(acc[2 * v + 1][0], acc[2 * v + 1][1], x0.y, b0);
                    dmma884(acc[2 * v + 1][0], acc[2 * v + 1][1], x1.y, b1);
                }
            }
            const int gt = g01 | (g2 << 2);
            double n2 = 0.0;
            double *dst = ((MODE == 0) ? coefs + (size_t)(child0 + gt) * ncoef : coefs + (size_t)parent * ncoef + (size_t)gt * Kd) +
                          8 * r + 64 * (2 * q);
#pragma unroll
            for (int v = 0; v < 4; v++) {
                double2 lo = make_double2(acc[2 * v][0], acc[2 * v + 1][0]);
                double2 hi = make_double2(acc[2 * v][1], acc[2 * v + 1][1]);
                if (MODE == 0 && !overwrite) {
                    const double2 a = *reinterpret_cast<const double2 *>(dst + 2 * v);
                    const double2 b = *reinterpret_cast<const double2 *>(dst + 64 + 2 * v);
                    lo.x += a.x;
                    lo.y += a.y;
                    hi.x += b.x;
                    hi.y += b.y;
                }
                *reinterpret_cast<double2 *>(dst + 2 * v) = lo;
                *reinterpret_cast<double2 *>(dst + 64 + 2 * v) = hi;
                n2 = fma(lo.x, lo.x, n2);
                n2 = fma(lo.y, lo.y, n2);
                n2 = fma(hi.x, hi.x, n2);
                n2 = fma(hi.y, hi.y, n2);
            }
            // component norm of the block just written (MWNode::calcNorms): scaling block of child gt (TopDown) or block gt of
            // the parent (BottomUp / in-node compression); fixed shuffle tree -> deterministic
            if (norms) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, off);
                if (lane == 0) norms[(MODE == 0) ? (size_t)(child0 + gt) * 8 : (size_t)parent * 8 + gt] = sqrt(n2);
            }
        }
    }
    if (MODE == 0 && overwrite) {
        // giveChildrenCoefs(overwrite=true) zeroes the children first (MWNode.cpp:317-319)
        for (int o = threadIdx.x; o < 8 * 7 * Kd / 2; o += 128) {
            const int c = o / (7 * Kd / 2), rem = o - c * (7 * Kd / 2);
            reinterpret_cast<double2 *>(coefs + (size_t)(child0 + c) * ncoef + Kd)[rem] = make_double2(0.0, 0.0);
        }
    }
}

// Tried and dropped (round 2, profiles/r02o_transform8_persistent_variant.txt): a PERSISTENT variant with its own 32 KB staging
// buffer next to the tiles and the next node's bulk copy issued as soon as the source fragments are in registers. It needs 77 KB
// of shared memory and 164 registers (the fragment tables stay live across the node loop), i.e. 2 CTAs = 8 warps per SM instead
// of 5 x 4, and ran SLOWER: BottomUp 0.708 vs 0.648 ms per pass on the 256 K-node C2 tree, TopDown(+=) 1.09 vs 1.00 ms on the
// bench tree. With five resident CTAs the hardware already overlaps the load of one node with the passes of another.

// ------------------------------------------------------------------------------------------------
// Even orders K = 4, 6, 8, 10, 12: two-scale transform on the FP64 tensor cores, one CTA (8 warps) per node.
// A pass contracts dimension p together with bit p of the block index: with the 2K x 2K two-scale matrix
//   F2[(b, t)][(gbit, j)] = F[2 gbit + b](t, j)                      (math_utils::apply_filter, math_utils.cpp:175-194)
// it is the GEMM  out[(gbit, j)][col] = sum_(b,t) F2[(b,t)][(gbit,j)] in[(b,t)][col]  over the 4 K^2 columns (the two other
// indices x the two other block bits). Treating the four K x K filter blocks as ONE 2K x 2K operand removes most of the tile
// padding: M = 2K rows in tiles of 8 (K = 12: 100 %, 10: 83 %, 6: 75 % useful), inner dimension 2K = K/2 steps of 4, exact.
// The node lives in shared memory for the three passes, which run IN PLACE (a column's outputs only replace that column's
// inputs, and a warp owns whole columns); the last pass writes its D fragments straight to global memory.
// Shared layout: element (blk, i0, i1, i2) at SB blk + i0 + S1 i1 + S2 i2 with S1 = 12 (mod 16), S2 = 4 or 12 (mod 16),
// SB = 2 (mod 16): with the column tiles chosen below (an index pair x two block bits) the fragment loads of all three
// passes and the fragment stores of passes 0 and 1 are free of bank conflicts.
constexpr int next_mod16(int x, int m) { return x + ((m - x % 16) + 16) % 16; }
constexpr int cmin(int a, int b) { return a < b ? a : b; }
template <int K> struct TLayout {
    static constexpr int K2 = K * K, Kd = K2 * K;
    static constexpr int S1 = (K == 4) ? 4 : 12;
    static constexpr int S2 = cmin(next_mod16((K - 1) * S1 + K, 4), next_mod16((K - 1) * S1 + K, 12));
    static constexpr int SB = next_mod16((K - 1) * S2 + (K - 1) * S1 + K, 2);
    static constexpr int MT = (2 * K + 7) / 8, KS = K / 2, NT = K2 / 2;
    static constexpr size_t bytes = (size_t)8 * SB * sizeof(double);
};

// MODE 0: TopDown; 1: BottomUp; 2: generated children (+ norms); 3: in-node compression
template <int K, int MODE>
__global__ void __launch_bounds__(256) transformK_kernel(double *__restrict__ coefs, const double *__restrict__ realCoefs,
                                                         double *__restrict__ genCoefs, double *__restrict__ genNorms, int nReal,
                                                         const int *__restrict__ pairs, const double *__restrict__ filters, int overwrite,
                                                         double *__restrict__ norms) {
    using L = TLayout<K>;
    constexpr int K2 = L::K2, Kd = L::Kd, ncoef = 8 * Kd, S1 = L::S1, S2 = L::S2, SB = L::SB, MT = L::MT, KS = L::KS, NT = L::NT;
    extern __shared__ __align__(16) double smK[];
    __shared__ double sN[8][8];
    const int parent = pairs[2 * blockIdx.x];
    const int child0 = pairs[2 * blockIdx.x + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    // ---- filter fragments: A[row = (gbit, j)][kc = (b, t)] = F[2 gbit + b](t, j)
    const double *F = filters + (size_t)((MODE == 1 || MODE == 3) ? 0 : 1) * 4 * K2;
    double a[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int row = 8 * mt + r, kc = 4 * s + q;
            a[mt][s] = (row < 2 * K) ? F[(2 * (row / K) + kc / K) * K2 + (kc % K) * K + row % K] : 0.0;
        }
    // ---- stage the node: 16-byte coalesced loads into the padded layout
    for (int e = tid; e < 4 * Kd; e += 256) {
        const int blk = e / (Kd / 2), rem = 2 * (e - blk * (Kd / 2));
        const int i0 = rem % K, i1 = (rem / K) % K, i2 = rem / K2;
        double2 v;
        if (MODE == 0 || MODE == 3) {
            v = *reinterpret_cast<const double2 *>(coefs + (size_t)parent * ncoef + (size_t)blk * Kd + rem);
        } else if (MODE == 1) {
            v = *reinterpret_cast<const double2 *>(coefs + (size_t)(child0 + blk) * ncoef + rem);
        } else {
            if (parent < nReal) v = *reinterpret_cast<const double2 *>(realCoefs + (size_t)parent * ncoef + (size_t)blk * Kd + rem);
            else v = (blk == 0) ? *reinterpret_cast<const double2 *>(genCoefs + (size_t)(parent - nReal) * Kd + rem) : make_double2(0.0, 0.0);
        }
        *reinterpret_cast<double2 *>(smK + SB * blk + i0 + S1 * i1 + S2 * i2) = v;
    }
    __syncthreads();
    // ---- pass 0: contract i0 / block bit 0. Tile = (i1 pair, fixed i2) x (bit 1, bit 2); column c = bit1 + 2 bit2 + 4 e1
    for (int tile = warp; tile < NT; tile += 8) {
        const int pa = tile % (K / 2), i2 = tile / (K / 2);
        const int cb = S1 * (2 * pa + (r >> 2)) + S2 * i2 + SB * (2 * (r & 1) + 4 * ((r >> 1) & 1));
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int kc = 4 * s + q;
            bf[s] = smK[cb + kc % K + SB * (kc / K)];
        }
        // D columns 2q, 2q+1: bit1 = 0 / 1, bit2 = q & 1, e1 = q >> 1
        const int ob = S1 * (2 * pa + (q >> 1)) + S2 * i2 + SB * (4 * (q & 1));
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < KS; s++) dmma884(d0, d1, a[mt][s], bf[s]);
            const int row = 8 * mt + r;
            if (row < 2 * K) {
                const int o = ob + row % K + SB * (row / K);
                smK[o] = d0;
                smK[o + 2 * SB] = d1;
            }
        }
    }
    __syncthreads();
    // ---- pass 1: contract i1 / block bit 1. Tile = (i0 pair, fixed i2) x (bit 0, bit 2); column c = e0 + 2 bit0 + 4 bit2
    for (int tile = warp; tile < NT; tile += 8) {
        const int pa = tile % (K / 2), i2 = tile / (K / 2);
        const int cb = 2 * pa + (r & 1) + S2 * i2 + SB * (((r >> 1) & 1) + 4 * (r >> 2));
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int kc = 4 * s + q;
            bf[s] = smK[cb + S1 * (kc % K) + 2 * SB * (kc / K)];
        }
        const int ob = 2 * pa + S2 * i2 + SB * ((q & 1) + 4 * (q >> 1)); // columns 2q, 2q+1 = the i0 pair
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < KS; s++) dmma884(d0, d1, a[mt][s], bf[s]);
            const int row = 8 * mt + r;
            if (row < 2 * K) *reinterpret_cast<double2 *>(smK + ob + S1 * (row % K) + 2 * SB * (row / K)) = make_double2(d0, d1);
        }
    }
    __syncthreads();
    // ---- pass 2: contract i2 / block bit 2, results to global. Tile = (i0 pair, fixed i1) x (bit 0, bit 1)
    double n2a = 0.0, n2b = 0.0; // square norm contributions to the blocks gt = q (gbit 0) and q + 4 (gbit 1)
    for (int tile = warp; tile < NT; tile += 8) {
        const int pa = tile % (K / 2), i1 = tile / (K / 2);
        const int cb = 2 * pa + (r & 1) + S1 * i1 + SB * (((r >> 1) & 1) + 2 * (r >> 2));
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int kc = 4 * s + q;
            bf[s] = smK[cb + S2 * (kc % K) + 4 * SB * (kc / K)];
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < KS; s++) dmma884(d0, d1, a[mt][s], bf[s]);
            const int row = 8 * mt + r;
            if (row < 2 * K) {
                const int gbit = row / K, j = row % K;
                const int gt = q + 4 * gbit; // bit0 = q & 1, bit1 = q >> 1
                const int elem = 2 * pa + K * i1 + K2 * j;
                double *dst;
                if (MODE == 0) dst = coefs + (size_t)(child0 + gt) * ncoef + elem;
                else if (MODE == 2) dst = genCoefs + (size_t)(child0 - nReal + gt) * Kd + elem;
                else dst = coefs + (size_t)parent * ncoef + (size_t)gt * Kd + elem;
                if (MODE == 0 && !overwrite) {
                    const double2 o = *reinterpret_cast<const double2 *>(dst);
                    d0 += o.x;
                    d1 += o.y;
                }
                *reinterpret_cast<double2 *>(dst) = make_double2(d0, d1);
                const double c2 = fma(d1, d1, d0 * d0);
                if (gbit) n2b += c2;
                else n2a += c2;
            }
        }
    }
    // ---- norms of the eight blocks written (fixed reduction shape: deterministic)
    if (norms != nullptr || MODE == 2) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
            double v = g ? n2b : n2a;
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (r == 0) sN[warp][q + 4 * g] = v;
        }
        __syncthreads();
        if (tid < 8) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) v += sN[w][tid];
            const double nrm = sqrt(v);
            if (MODE == 2) genNorms[child0 - nReal + tid] = nrm;
            else if (MODE == 0) norms[(size_t)(child0 + tid) * 8] = nrm;
            else norms[(size_t)parent * 8 + tid] = nrm;
        }
    }
    if (MODE == 0 && overwrite) {
        // giveChildrenCoefs(overwrite=true) zeroes the children first (MWNode.cpp:317-319)
        for (int o = tid; o < 8 * 7 * Kd / 2; o += 256) {
            const int c = o / (7 * Kd / 2), rem = o - c * (7 * Kd / 2);
            reinterpret_cast<double2 *>(coefs + (size_t)(child0 + c) * ncoef + Kd)[rem] = make_double2(0.0, 0.0);
        }
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The same kernel as a PERSISTENT CTA with the next node's blocks in flight while the current node is transformed: the staging
// loads are cp.async (LDGSTS) copies into the second of two shared-memory buffers, committed before the three passes of the
// current node start and waited for after they end. ncu of the one-node-per-CTA kernel (profiles/r02b_ncu_transform_kernels.txt)
// showed neither roof near (K = 10: DMMA sub-pipe 38-45 %, DRAM 16-22 %, largest stall long scoreboard on the staging loads,
// 2 CTAs per SM): load -> passes -> store were serial inside a CTA. Here the load of node i + 1 overlaps the passes of node i.
// Used where two node buffers fit into shared memory (K <= 10); arithmetic and summation order are those of transformK_kernel.
template <int K, int MODE, int NW>
__global__ void __launch_bounds__(32 * NW) transformK_pipe_kernel(double *__restrict__ coefs, const double *__restrict__ realCoefs,
                                                         double *__restrict__ genCoefs, double *__restrict__ genNorms, int nReal,
                                                         const int *__restrict__ pairs, const double *__restrict__ filters, int overwrite,
                                                         double *__restrict__ norms, int cnt) {
    using L = TLayout<K>;
    constexpr int K2 = L::K2, Kd = L::Kd, ncoef = 8 * Kd, S1 = L::S1, S2 = L::S2, SB = L::SB, MT = L::MT, KS = L::KS, NT = L::NT;
    extern __shared__ __align__(16) double smAll[]; // two node buffers of 8 SB doubles
    __shared__ double sN[NW][8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = lane >> 2, q = lane & 3;
    // ---- filter fragments: A[row = (gbit, j)][kc = (b, t)] = F[2 gbit + b](t, j)
    const double *F = filters + (size_t)((MODE == 1 || MODE == 3) ? 0 : 1) * 4 * K2;
    double a[MT][KS];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int row = 8 * mt + r, kc = 4 * s + q;
            a[mt][s] = (row < 2 * K) ? F[(2 * (row / K) + kc / K) * K2 + (kc % K) * K + row % K] : 0.0;
        }
    // ---- staging of one node into a buffer: 16-byte asynchronous copies into the padded layout
    auto stage = [&](int item, double *buf) {
        const int sp = pairs[2 * item], sc0 = pairs[2 * item + 1];
        for (int e = tid; e < 4 * Kd; e += 32 * NW) {
            const int blk = e / (Kd / 2), rem = 2 * (e - blk * (Kd / 2));
            const int i0 = rem % K, i1 = (rem / K) % K, i2 = rem / K2;
            double *dst = buf + SB * blk + i0 + S1 * i1 + S2 * i2;
            const double *src = nullptr;
            if (MODE == 0 || MODE == 3) {
                src = coefs + (size_t)sp * ncoef + (size_t)blk * Kd + rem;
            } else if (MODE == 1) {
                src = coefs + (size_t)(sc0 + blk) * ncoef + rem;
            } else {
                if (sp < nReal) src = realCoefs + (size_t)sp * ncoef + (size_t)blk * Kd + rem;
                else if (blk == 0) src = genCoefs + (size_t)(sp - nReal) * Kd + rem;
            }
            if (src) cp_async16(dst, src);
            else *reinterpret_cast<double2 *>(dst) = make_double2(0.0, 0.0);
        }
    };
    int item = blockIdx.x, nIt = 0;
    if (item < cnt) stage(item, smAll);
    cp_async_commit();
    for (; item < cnt; item += gridDim.x, nIt++) {
    double *smK = smAll + (size_t)(nIt & 1) * 8 * SB;
    const int parent = pairs[2 * item];
    const int child0 = pairs[2 * item + 1];
    if (item + (int)gridDim.x < cnt) stage(item + gridDim.x, smAll + (size_t)((nIt + 1) & 1) * 8 * SB);
    cp_async_commit();
    cp_async_wait<1>(); // everything but the group just committed has landed: the current node is in its buffer
    __syncthreads();
    // ---- pass 0: contract i0 / block bit 0. Tile = (i1 pair, fixed i2) x (bit 1, bit 2); column c = bit1 + 2 bit2 + 4 e1
    for (int tile = warp; tile < NT; tile += NW) {
        const int pa = tile % (K / 2), i2 = tile / (K / 2);
        const int cb = S1 * (2 * pa + (r >> 2)) + S2 * i2 + SB * (2 * (r & 1) + 4 * ((r >> 1) & 1));
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int kc = 4 * s + q;
            bf[s] = smK[cb + kc % K + SB * (kc / K)];
        }
        // D columns 2q, 2q+1: bit1 = 0 / 1, bit2 = q & 1, e1 = q >> 1
        const int ob = S1 * (2 * pa + (q >> 1)) + S2 * i2 + SB * (4 * (q & 1));
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < KS; s++) dmma884(d0, d1, a[mt][s], bf[s]);
            const int row = 8 * mt + r;
            if (row < 2 * K) {
                const int o = ob + row % K + SB * (row / K);
                smK[o] = d0;
                smK[o + 2 * SB] = d1;
            }
        }
    }
    __syncthreads();
    // ---- pass 1: contract i1 / block bit 1. Tile = (i0 pair, fixed i2) x (bit 0, bit 2); column c = e0 + 2 bit0 + 4 bit2
    for (int tile = warp; tile < NT; tile += NW) {
        const int pa = tile % (K / 2), i2 = tile / (K / 2);
        const int cb = 2 * pa + (r & 1) + S2 * i2 + SB * (((r >> 1) & 1) + 4 * (r >> 2));
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int kc = 4 * s + q;
            bf[s] = smK[cb + S1 * (kc % K) + 2 * SB * (kc / K)];
        }
        const int ob = 2 * pa + S2 * i2 + SB * ((q & 1) + 4 * (q >> 1)); // columns 2q, 2q+1 = the i0 pair
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < KS; s++) dmma884(d0, d1, a[mt][s], bf[s]);
            const int row = 8 * mt + r;
            if (row < 2 * K) *reinterpret_cast<double2 *>(smK + ob + S1 * (row % K) + 2 * SB * (row / K)) = make_double2(d0, d1);
        }
    }
    __syncthreads();
    // ---- pass 2: contract i2 / block bit 2, results to global. Tile = (i0 pair, fixed i1) x (bit 0, bit 1)
    double n2a = 0.0, n2b = 0.0; // square norm contributions to the blocks gt = q (gbit 0) and q + 4 (gbit 1)
    for (int tile = warp; tile < NT; tile += NW) {
        const int pa = tile % (K / 2), i1 = tile / (K / 2);
        const int cb = 2 * pa + (r & 1) + S1 * i1 + SB * (((r >> 1) & 1) + 2 * (r >> 2));
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) {
            const int kc = 4 * s + q;
            bf[s] = smK[cb + S2 * (kc % K) + 4 * SB * (kc / K)];
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int s = 0; s < KS; s++) dmma884(d0, d1, a[mt][s], bf[s]);
            const int row = 8 * mt + r;
            if (row < 2 * K) {
                const int gbit = row / K, j = row % K;
                const int gt = q + 4 * gbit; // bit0 = q & 1, bit1 = q >> 1
                const int elem = 2 * pa + K * i1 + K2 * j;
                double *dst;
                if (MODE == 0) dst = coefs + (size_t)(child0 + gt) * ncoef + elem;
                else if (MODE == 2) dst = genCoefs + (size_t)(child0 - nReal + gt) * Kd + elem;
                else dst = coefs + (size_t)parent * ncoef + (size_t)gt * Kd + elem;
                if (MODE == 0 && !overwrite) {
                    const double2 o = *reinterpret_cast<const double2 *>(dst);
                    d0 += o.x;
                    d1 += o.y;
                }
                *reinterpret_cast<double2 *>(dst) = make_double2(d0, d1);
                const double c2 = fma(d1, d1, d0 * d0);
                if (gbit) n2b += c2;
                else n2a += c2;
            }
        }
    }
    // ---- norms of the eight blocks written (fixed reduction shape: deterministic)
    if (norms != nullptr || MODE == 2) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
            double v = g ? n2b : n2a;
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (r == 0) sN[warp][q + 4 * g] = v;
        }
        __syncthreads();
        if (tid < 8) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; w++) v += sN[w][tid];
            const double nrm = sqrt(v);
            if (MODE == 2) genNorms[child0 - nReal + tid] = nrm;
            else if (MODE == 0) norms[(size_t)(child0 + tid) * 8] = nrm;
            else norms[(size_t)parent * 8 + tid] = nrm;
        }
    }
    if (MODE == 0 && overwrite) {
        // giveChildrenCoefs(overwrite=true) zeroes the children first (MWNode.cpp:317-319)
        for (int o = tid; o < 8 * 7 * Kd / 2; o += 32 * NW) {
            const int c = o / (7 * Kd / 2), rem = o - c * (7 * Kd / 2);
            reinterpret_cast<double2 *>(coefs + (size_t)(child0 + c) * ncoef + Kd)[rem] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads(); // this node's buffer is refilled by the staging of the node after next; sN is rewritten
    } // nodes of this CTA
}

// persistent variant with NW warps per CTA: grid (0 where two node buffers do not fit or MRX_NO_TPIPE=1), configured once
template <int K, int MODE, int NW> int transformK_pipe_grid() {
    static int grid = -1;
    if (grid < 0) {
        grid = 0;
        if (2 * TLayout<K>::bytes <= 220 * 1024 && !getenv("MRX_NO_TPIPE")) {
            int dev = 0, sms = 0, perSm = 0;
            MRX_CUDA(cudaGetDevice(&dev));
            MRX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            MRX_CUDA(cudaFuncSetAttribute(transformK_pipe_kernel<K, MODE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * TLayout<K>::bytes)));
            MRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, transformK_pipe_kernel<K, MODE, NW>, 32 * NW, 2 * TLayout<K>::bytes));
            grid = sms * std::max(perSm, 1);
        }
    }
    return grid;
}

template <int K, int MODE, int NW>
bool try_transformK_pipe(double *coefs, const double *realCoefs, double *genCoefs, double *genNorms, int nReal, const int *pairs, int cnt,
                         const double *filters, int overwrite, double *norms, cudaStream_t st) {
    const int pipeGrid = transformK_pipe_grid<K, MODE, NW>();
    // enough nodes per CTA for the prefetch to pay; TopDown(overwrite) at small K is dominated by zeroing the children's wavelet
    // blocks, which a one-node CTA overlaps better with its neighbours on the SM
    if (pipeGrid > 0 && cnt >= 2 * pipeGrid && !(MODE == 0 && overwrite && K < 10)) {
        transformK_pipe_kernel<K, MODE, NW><<<pipeGrid, 32 * NW, 2 * TLayout<K>::bytes, st>>>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs,
                                                                                            filters, overwrite, norms, cnt);
        return true;
    }
    return false;
}

template <int K, int MODE>
void launch_transformK(double *coefs, const double *realCoefs, double *genCoefs, double *genNorms, int nReal, const int *pairs, int cnt,
                       const double *filters, int overwrite, double *norms, cudaStream_t st) {
    // warps of the persistent variant: K = 6 has 18 column tiles per pass, dealt evenly to 9 warps; K = 10 measured faster with 8
    // warps (7 rounds of tiles) than with 10 (5 exact rounds): 1.26 vs 1.32 ms per BottomUp pass on the C2 tree. ncu of the 8-warp
    // CTA at K = 10 (profiles/r02ag_ncu_transformK_pipe.txt): DMMA sub-pipe 50 % active, DRAM 18-25 %, 2 warps per scheduler, largest
    // stalls wait + math-pipe throttle: 16 warps per CTA at K = 10 (4 per scheduler): BottomUp 1.257 -> 1.185 ms; 17 warps (3 rounds of the 50 tiles) 1.231, 25 warps (2 exact rounds) 1.239 (profiles/r02ai). MRX_TPIPE_WARPS=8: the 8-warp CTA
    constexpr int NW = (K == 6) ? 9 : 8;
    static bool conf = false;
    static int wide = 0;
    if (!conf) {
        MRX_CUDA(cudaFuncSetAttribute(transformK_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TLayout<K>::bytes));
        wide = (K == 10) ? (getenv("MRX_TPIPE_WARPS") ? atoi(getenv("MRX_TPIPE_WARPS")) : 16) : 0;
        conf = true;
    }
    constexpr int W16 = (K == 10) ? 16 : NW;
    bool done;
    if (K == 10 && wide == 16) done = try_transformK_pipe<K, MODE, W16>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st);
    else done = try_transformK_pipe<K, MODE, NW>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st);
    if (done) return;
    transformK_kernel<K, MODE><<<cnt, 256, TLayout<K>::bytes, st>>>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, filters, overwrite, norms);
}

// even orders served by transformK_kernel (K = 8 keeps its register-chained kernel unless MRX_TK8 is set)
bool transformK_supports(int K) {
    static const bool k8 = getenv("MRX_TK8") != nullptr;
    return K == 4 || K == 6 || K == 10 || K == 12 || (K == 8 && k8);
}

template <int MODE>
void dispatch_transformK(int K, double *coefs, const double *realCoefs, double *genCoefs, double *genNorms, int nReal, const int *pairs,
                         int cnt, const double *filters, int overwrite, double *norms, cudaStream_t st) {
    switch (K) {
    case 4: launch_transformK<4, MODE>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st); break;
    case 6: launch_transformK<6, MODE>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st); break;
    case 8: launch_transformK<8, MODE>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st); break;
    case 10: launch_transformK<10, MODE>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st); break;
    case 12: launch_transformK<12, MODE>(coefs, realCoefs, genCoefs, genNorms, nReal, pairs, cnt, filters, overwrite, norms, st); break;
    default: MRX_ABORT("transformK: unsupported order");
    }
}

__global__ void __launch_bounds__(256) norms_kernel(const double *__restrict__ coefs, double *__restrict__ norms,
                                                    const int *__restrict__ slots, int Kd, double *__restrict__ normsW) {
    int node = slots ? slots[blockIdx.x] : blockIdx.x;
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double *b = coefs + ((size_t)node * 8 + w) * Kd;
    double s = 0.0;
    for (int e = lane; e < Kd; e += 32) {
        double v = b[e];
        s = fma(v, v, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        norms[(size_t)node * 8 + w] = sqrt(s);
        if (normsW) normsW[(size_t)blockIdx.x * 8 + w] = sqrt(s); // work-vector order for the host bookkeeping
    }
}

__global__ void __launch_bounds__(256) dot_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                  const int *__restrict__ pairs, double *__restrict__ res, int nRoots, int Kd) {
    __shared__ double part[8];
    int na = pairs[2 * blockIdx.x], nb = pairs[2 * blockIdx.x + 1];
    const double *pa = a + (size_t)na * 8 * Kd, *pb = b + (size_t)nb * 8 * Kd;
    int start = (na < nRoots) ? 0 : Kd; // scaling part only at the roots
    double s = 0.0;
    for (int e = start + threadIdx.x; e < 8 * Kd; e += 256) s = fma(pa[e], pb[e], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += part[i];
        res[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(double *__restrict__ coefs, const double *__restrict__ partials,
                                                              const int *__restrict__ items, int ncoef) {
    const int slot = items[3 * blockIdx.x], first = items[3 * blockIdx.x + 1], n = items[3 * blockIdx.x + 2];
    for (int e = threadIdx.x; e < ncoef; e += 256) {
        double s = 0.0;
        for (int c = 0; c < n; c++) s += partials[(size_t)(first + c) * ncoef + e]; // chunk order = neighbour order
        coefs[(size_t)slot * ncoef + e] = s;
    }
}

__global__ void scale_kernel(double *x, size_t n, double c) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) x[i] *= c;
}

// AdditionCalculator::calcNode (src/treebuilders/AdditionCalculator.h:42-66) for the nodes two trees share: out node o +=
// c * in node i. Wavelet blocks always; the scaling block only on root nodes -- every other scaling block of the sum is
// produced by the TopDown(+=) pass that follows (device_add), which also covers output nodes finer than the input tree.
// One CTA per pair; pairs of one launch have distinct output nodes, so there is no race and the summation order over the
// input trees is the launch order.
__global__ void __launch_bounds__(256) axpy_nodes_kernel(double *out, const double *in, const int *pairs, int nRoots, int Kd, double c) {
    const int o = pairs[2 * blockIdx.x], i = pairs[2 * blockIdx.x + 1];
    double *po = out + (size_t)o * 8 * Kd;
    const double *pi = in + (size_t)i * 8 * Kd;
    for (int j = (o < nRoots ? 0 : Kd) + threadIdx.x; j < 8 * Kd; j += 256) po[j] = __dadd_rn(po[j], __dmul_rn(c, pi[j]));
}

// MultiplicationCalculator::calcNode's element-wise part (MultiplicationCalculator.h:43-72) on scratch nodes: node j of the chunk
// has its eight reconstructed scaling blocks in block 0 of the scratch slots nC + 8 j + t (t = child). map = cvMap (Forward:
// sqrt(1 / w), InterpolatingBasis.cpp:115-124) or vcMap (Backward: sqrt(w)), applied along x, y, z like MWNode::cvTransform
// (MWNode.cpp:448-490), with the factor 2^(+-3 (n + 1) / 2) of the children's scale.
//   mode 0: P  = c * forward(S)      mode 1: P *= c * forward(S)      mode 2: P = backward(P)      mode 3: P = forward(S) ^ c
//   (mode 3: PowerCalculator.h:43-58)
__global__ void __launch_bounds__(256) product_values_kernel(double *P, const double *S, const int *scale, int nC, int K, const double *map,
                                                             double c, int mode) {
    const int j = blockIdx.x >> 3, t = blockIdx.x & 7;
    const int Kd = K * K * K;
    const size_t off = ((size_t)(nC + 8 * j + t) * 8) * Kd; // block 0 of the child slot
    const int np1 = scale[j] + 1;
    const double two_fac = mode == 2 ? sqrt(1.0 / exp2((double)(3 * np1))) : sqrt(exp2((double)(3 * np1)));
    for (int q = threadIdx.x; q < Kd; q += 256) {
        const int x = q % K, y = (q / K) % K, z = q / (K * K);
        if (mode == 2) {
            P[off + q] = two_fac * (((P[off + q] * map[x]) * map[y]) * map[z]);
        } else if (mode == 3) {
            P[off + q] = pow(two_fac * (((S[off + q] * map[x]) * map[y]) * map[z]), c);
        } else {
            const double v = c * (two_fac * (((S[off + q] * map[x]) * map[y]) * map[z]));
            P[off + q] = mode == 0 ? v : P[off + q] * v;
        }
    }
}

// MWNode::cvTransform (MWNode.cpp:448-490) as a standalone pass over a list of nodes: the node's 8 blocks hold the scaling
// coefficients of its children (0/1 representation, after mwTransform(Reconstruction)); Forward turns them into function values
// at the children's quadrature points, Backward back. For the interpolating basis the coefficient-value map is diagonal
// (InterpolatingBasis.cpp:115-124: sqrt(1 / w_j) forward, sqrt(w_j) backward), so the three apply_filter passes of the reference
// reduce to ((c m[x]) m[y]) m[z] -- same operation order -- times 2^(+-3 (n + 1) / 2) of the children's scale.
// Element-wise and HBM-bound: 8 K^3 doubles read and written per node (128 K^3 B), 16-byte accesses, one CTA per node.
__global__ void __launch_bounds__(256) cv_transform_kernel(double *__restrict__ coefs, const int *__restrict__ items, int K,
                                                           const double *__restrict__ map, int backward) {
    __shared__ double sMap[16];
    const int slot = items[2 * blockIdx.x], scale = items[2 * blockIdx.x + 1];
    const int Kd = K * K * K, ncoef = 8 * Kd;
    if (threadIdx.x < K) sMap[threadIdx.x] = map[threadIdx.x];
    __syncthreads();
    const int np1 = scale + 1;
    const double two_fac = backward ? sqrt(1.0 / exp2((double)(3 * np1))) : sqrt(exp2((double)(3 * np1)));
    double *c = coefs + (size_t)slot * ncoef;
    if ((Kd & 1) == 0) {
        double2 *c2 = reinterpret_cast<double2 *>(c);
        for (int o = threadIdx.x; o < ncoef / 2; o += 256) {
            const int q = (2 * o) % Kd; // K even: a pair never straddles a row
            const int x = q % K, y = (q / K) % K, z = q / (K * K);
            double2 v = c2[o];
            const double myz0 = sMap[y], mz = sMap[z];
            v.x = two_fac * (((v.x * sMap[x]) * myz0) * mz);
            v.y = two_fac * (((v.y * sMap[x + 1]) * myz0) * mz);
            c2[o] = v;
        }
    } else {
        for (int o = threadIdx.x; o < ncoef; o += 256) {
            const int q = o % Kd;
            const int x = q % K, y = (q / K) % K, z = q / (K * K);
            c[o] = two_fac * (((c[o] * sMap[x]) * sMap[y]) * sMap[z]);
        }
    }
}

size_t transform_smem(int K, int &padOn) {
    int K2 = K * K, Kd = K2 * K;
    padOn = 1;
    size_t bytes = ((size_t)16 * (Kd + K2) + 4 * K2) * sizeof(double);
    if (bytes > 227 * 1024) {
        padOn = 0;
        bytes = ((size_t)16 * Kd + 4 * K2) * sizeof(double);
    }
    if (bytes > 227 * 1024) MRX_ABORT("transform kernel: order too large for shared memory staging");
    return bytes;
}

template <int MODE> void set_smem_attr(size_t bytes) {
    static size_t configured = 0;
    if (bytes > configured) {
        MRX_CUDA(cudaFuncSetAttribute(transform_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
}

} // namespace

void launch_norms(const double *coefs, double *norms, const int *slots, int n, int Kd, cudaStream_t st, double *normsW) {
    if (n <= 0) return;
    norms_kernel<<<n, 256, 0, st>>>(coefs, norms, slots, Kd, normsW);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

bool transform_fuses_norms(int K) { return K == 8 || transformK_supports(K); }

void launch_transform(bool down, bool overwrite, double *coefs, const int *pairs, int cnt, int K, const double *filters,
                      cudaStream_t st, double *norms) {
    if (cnt <= 0) return;
    if (transformK_supports(K)) {
        if (down) dispatch_transformK<0>(K, coefs, nullptr, nullptr, nullptr, 0, pairs, cnt, filters, overwrite ? 1 : 0, norms, st);
        else dispatch_transformK<1>(K, coefs, nullptr, nullptr, nullptr, 0, pairs, cnt, filters, 1, norms, st);
        MRX_CUDA(cudaGetLastError());
        launch_counter()++;
        return;
    }
    if (K == 8) {
        constexpr size_t bytes8 = kT8Bytes;
        static bool conf = false;
        if (!conf) {
            MRX_CUDA(cudaFuncSetAttribute(transform8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes8));
            MRX_CUDA(cudaFuncSetAttribute(transform8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes8));
            conf = true;
        }
        if (down) transform8_kernel<0><<<cnt, 128, bytes8, st>>>(coefs, pairs, filters, overwrite ? 1 : 0, norms);
        else transform8_kernel<1><<<cnt, 128, bytes8, st>>>(coefs, pairs, filters, 1, norms);
        MRX_CUDA(cudaGetLastError());
        launch_counter()++;
        return;
    }
    int padOn;
    size_t bytes = transform_smem(K, padOn);
    if (down) {
        set_smem_attr<0>(bytes);
        transform_kernel<0><<<cnt, kTransformThreads, bytes, st>>>(coefs, nullptr, nullptr, nullptr, 0, pairs, K, padOn, filters,
                                                                  overwrite ? 1 : 0);
    } else {
        set_smem_attr<1>(bytes);
        transform_kernel<1><<<cnt, kTransformThreads, bytes, st>>>(coefs, nullptr, nullptr, nullptr, 0, pairs, K, padOn, filters, 1);
    }
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_compress_nodes(double *coefs, const int *pairs, int cnt, int K, const double *filters, cudaStream_t st, double *norms) {
    if (cnt <= 0) return;
    if (transformK_supports(K)) {
        dispatch_transformK<3>(K, coefs, nullptr, nullptr, nullptr, 0, pairs, cnt, filters, 1, norms, st);
        MRX_CUDA(cudaGetLastError());
        launch_counter()++;
        return;
    }
    if (K == 8) {
        constexpr size_t bytes8 = kT8Bytes;
        static bool conf = false;
        if (!conf) {
            MRX_CUDA(cudaFuncSetAttribute(transform8_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes8));
            conf = true;
        }
        transform8_kernel<3><<<cnt, 128, bytes8, st>>>(coefs, pairs, filters, 1, norms);
    } else {
        int padOn;
        size_t bytes = transform_smem(K, padOn);
        set_smem_attr<3>(bytes);
        transform_kernel<3><<<cnt, kTransformThreads, bytes, st>>>(coefs, nullptr, nullptr, nullptr, 0, pairs, K, padOn, filters, 1);
    }
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_reconstruct_nodes(double *coefs, const int *pairs, int cnt, int K, const double *filters, cudaStream_t st) {
    if (cnt <= 0) return;
    int padOn;
    size_t bytes = transform_smem(K, padOn);
    set_smem_attr<4>(bytes);
    transform_kernel<4><<<cnt, kTransformThreads, bytes, st>>>(coefs, nullptr, nullptr, nullptr, 0, pairs, K, padOn, filters, 1);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_cv_transform(double *coefs, const int *items, int cnt, int K, const double *map, bool backward, cudaStream_t st) {
    if (cnt <= 0) return;
    if (K > 16) MRX_ABORT("cvTransform: order too large");
    cv_transform_kernel<<<cnt, 256, 0, st>>>(coefs, items, K, map, backward ? 1 : 0);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_gen_children(const double *realCoefs, double *genCoefs, double *genNorms, int nReal, const int *items, int cnt,
                         int K, const double *filters, cudaStream_t st) {
    if (cnt <= 0) return;
    if (transformK_supports(K) || K == 8) {
        dispatch_transformK<2>(K, nullptr, realCoefs, genCoefs, genNorms, nReal, items, cnt, filters, 1, nullptr, st);
        MRX_CUDA(cudaGetLastError());
        launch_counter()++;
        return;
    }
    int padOn;
    size_t bytes = transform_smem(K, padOn);
    set_smem_attr<2>(bytes);
    transform_kernel<2><<<cnt, kTransformThreads, bytes, st>>>(nullptr, realCoefs, genCoefs, genNorms, nReal, items, K, padOn,
                                                              filters, 1);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_reduce_partials(double *coefs, const double *partials, const int *items, int cnt, int ncoef, cudaStream_t st) {
    if (cnt <= 0) return;
    reduce_partials_kernel<<<cnt, 256, 0, st>>>(coefs, partials, items, ncoef);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_dot(const double *a, const double *b, const int *pairs, double *res, int np, int nRoots, int Kd, cudaStream_t st) {
    if (np <= 0) return;
    dot_kernel<<<np, 256, 0, st>>>(a, b, pairs, res, nRoots, Kd);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_axpy_nodes(double *out, const double *in, const int *pairs, int np, int nRoots, int Kd, double c, cudaStream_t st) {
    if (np == 0) return;
    axpy_nodes_kernel<<<np, 256, 0, st>>>(out, in, pairs, nRoots, Kd, c);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_product_values(double *P, const double *S, const int *scale, int nC, int K, const double *map, double c, int mode,
                           cudaStream_t st) {
    if (nC == 0) return;
    product_values_kernel<<<8 * nC, 256, 0, st>>>(P, S, scale, nC, K, map, c, mode);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_scale(double *x, size_t n, double c, cudaStream_t st) {
    if (n == 0) return;
    scale_kernel<<<1184, 256, 0, st>>>(x, n, c);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
