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
        if (MODE == 0 || MODE == 3) {
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
            } else if (MODE == 1 || MODE == 3) {
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
constexpr int kT8Si = 18, kT8Sm = 152, kT8Doubles = 8 * kT8Sm; // same padded tile as the apply kernel

template <int MODE> // 0: TopDown (parent 8 blocks -> children scaling, = or +=); 1: BottomUp (children scaling -> parent); 3: in-node compression
__global__ void __launch_bounds__(128) transform8_kernel(double *__restrict__ coefs, const int *__restrict__ pairs,
                                                         const double *__restrict__ filters, int overwrite, double *__restrict__ norms) {
    extern __shared__ __align__(128) double tiles8[]; // 8 padded tiles (kT8Doubles each) + the TMA-staged node (8 x 512)
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
    double *inbuf = tiles8 + 8 * kT8Doubles;
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

bool transform_fuses_norms(int K) { return K == 8; }

void launch_transform(bool down, bool overwrite, double *coefs, const int *pairs, int cnt, int K, const double *filters,
                      cudaStream_t st, double *norms) {
    if (cnt <= 0) return;
    if (K == 8) {
        constexpr size_t bytes8 = (size_t)(8 * kT8Doubles + 8 * 512) * sizeof(double);
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
    if (K == 8) {
        constexpr size_t bytes8 = (size_t)(8 * kT8Doubles + 8 * 512) * sizeof(double);
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

void launch_gen_children(const double *realCoefs, double *genCoefs, double *genNorms, int nReal, const int *items, int cnt,
                         int K, const double *filters, cudaStream_t st) {
    if (cnt <= 0) return;
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

void launch_scale(double *x, size_t n, double c, cudaStream_t st) {
    if (n == 0) return;
    scale_kernel<<<1184, 256, 0, st>>>(x, n, c);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
