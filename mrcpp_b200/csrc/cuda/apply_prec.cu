// Locally scaled precision of mrcpp::apply(prec, out, oper, inp, precTrees, maxIter, absPrec)
// (src/treebuilders/apply.cpp:214-251) on the device.
//
// Reference: for every output node the precision is multiplied by
//     precFac(idx) = 1 / max_i sqrt( precTrees[i].getNode(idx).getMaxSquareNorm() )           apply.cpp:222-234
// after makeMaxSquareNorms() on every precision tree (MWTree.cpp:536-543, MWNode.cpp:1257-1269: largest scaled square norm
// 2^(3 n) |node|^2 among a node and its descendants). getNode(idx) GENERATES the node where the precision tree is coarser
// than the output grid (MWTree.cpp:340-352, FunctionNode::genChildren FunctionNode.cpp:293-328); a generated node carries no
// stored maximum and answers with its own scaled square norm (MWNode.h:84). The factor enters the screening threshold
// gThrs = prec * precFac * sqrt(|g|^2 / M) (ConvolutionCalculator.cpp:241-248) and the split threshold
// (WaveletAdaptor.h:51-54 -> tree_utils::split_check).
//
// Here: one CTA per item of the work vector. A thread descends each precision tree's child pointers to the deepest real node
// on the way to (scale, l). Real node at the target scale: the host-computed per-node value (stored maximum if positive, else
// the node's own scaled norm). Coarser: the CTA reconstructs the ONE descendant scaling block on the path level by level in
// shared memory -- the first step from the real leaf's eight blocks, every further step from a scaling block alone, exactly
// what genChildren produces (tree_utils::mw_transform, tree_utils.cpp:113-216) -- and takes its norm. Nothing is stored: the
// precision trees stay untouched (the reference's generated nodes are deleted again after the apply, apply.cpp:246).
#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

constexpr int kPrecThreads = 128;

// One filter pass of the reconstruction restricted to the child bit `gbit` along dimension `pass`:
//   out[blk'][j K^2 + m] = sum_b sum_t in[blk(b)][m K + t] F[2 gbit + b][t K + j]
// `nIn` live input blocks (8, 4, 2 for a real parent; 1 for a scaling-only parent, where b = 0 only), halved by the pass.
// Blocks are indexed by the remaining wavelet bits of the dimensions not contracted yet (bit d of the block index = wavelet
// along d), compacted: after pass p only dimensions > p keep a bit.
__device__ __forceinline__ void child_pass(const double *__restrict__ in, double *__restrict__ out, const double *__restrict__ F, int K, int gbit,
                                           int nIn) {
    const int K2 = K * K, Kd = K2 * K;
    const int nOut = nIn > 1 ? nIn / 2 : 1;
    for (int o = threadIdx.x; o < nOut * Kd; o += kPrecThreads) {
        const int blk = o / Kd, rem = o - blk * Kd;
        const int j = rem / K2, m = rem - j * K2;
        double acc = 0.0;
        const int nb = nIn > 1 ? 2 : 1;
        for (int b = 0; b < nb; b++) {
            // input block: bit 0 of the compacted index is the dimension being contracted
            const double *inb = in + (size_t)(2 * blk + b) * Kd * (nIn > 1 ? 1 : 0) + (size_t)m * K;
            const double *Fm = F + (2 * gbit + b) * K2 + j;
            for (int t = 0; t < K; t++) acc = fma(inb[t], Fm[t * K], acc);
        }
        out[o] = acc;
    }
}

__global__ void __launch_bounds__(kPrecThreads) prec_factor_kernel(PrecParams P) {
    extern __shared__ double psm[];
    const int K = P.K, K2 = K * K, Kd = K2 * K;
    double *A = psm;          // up to 8 blocks
    double *B = A + 8 * Kd;   // up to 4 blocks
    double *F = B + 4 * Kd;   // reconstruction filter, 4 K^2
    __shared__ int sNode, sDepth;
    __shared__ double sRed[kPrecThreads / 32];
    const int nG = P.nGptr ? *P.nGptr : P.nG;
    const int i = blockIdx.x;
    if (i >= nG) return;
    const int tid = threadIdx.x;
    for (int e = tid; e < 4 * K2; e += kPrecThreads) F[e] = P.filters[(size_t)4 * K2 + e]; // op 1 = Reconstruction
    const int4 gn = P.gNodesAll[i];
    const int td = gn.x + P.depthShift; // depth in the function trees
    const int lx = gn.y, ly = gn.z, lz = gn.w;
    double maxNorm = P.nTrees ? 0.0 : 1.0;
    for (int p = 0; p < P.nTrees; p++) {
        const PrecTreeDev T = P.trees[p];
        if (tid == 0) {
            int node = ((lx >> td) - P.corner[0]) + P.nboxes[0] * (((ly >> td) - P.corner[1]) + P.nboxes[1] * ((lz >> td) - P.corner[2]));
            int nd = 0;
            while (nd < td) {
                const int c0 = T.child0[node];
                if (c0 < 0) break;
                const int shift = td - nd - 1;
                node = c0 + (((lx >> shift) & 1) | (((ly >> shift) & 1) << 1) | (((lz >> shift) & 1) << 2));
                nd++;
            }
            sNode = node;
            sDepth = nd;
        }
        __syncthreads();
        const int node = sNode;
        int nd = sDepth;
        double v;
        if (nd == td) {
            v = T.vReal[node];
        } else {
            // generated descendant: first step from the real leaf's 8 blocks, then scaling-only steps
            for (int e = tid; e < 8 * Kd; e += kPrecThreads) A[e] = T.coefs[(size_t)node * 8 * Kd + e];
            __syncthreads();
            int nIn = 8;
            while (nd < td) {
                const int shift = td - nd - 1;
                const int c = ((lx >> shift) & 1) | (((ly >> shift) & 1) << 1) | (((lz >> shift) & 1) << 2);
                child_pass(A, B, F, K, c & 1, nIn);
                __syncthreads();
                nIn = nIn > 1 ? nIn / 2 : 1;
                child_pass(B, A, F, K, (c >> 1) & 1, nIn);
                __syncthreads();
                nIn = nIn > 1 ? nIn / 2 : 1;
                child_pass(A, B, F, K, (c >> 2) & 1, nIn);
                __syncthreads();
                for (int e = tid; e < Kd; e += kPrecThreads) A[e] = B[e];
                __syncthreads();
                nIn = 1;
                nd++;
            }
            double s = 0.0;
            for (int e = tid; e < Kd; e += kPrecThreads) s = fma(A[e], A[e], s);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if ((tid & 31) == 0) sRed[tid >> 5] = s;
            __syncthreads();
            s = 0.0;
            for (int w = 0; w < kPrecThreads / 32; w++) s += sRed[w];
            v = ldexp(s, 3 * (td + P.rootScale)); // scaled square norm 2^(3 n) |node|^2 (MWNode::calcScaledSquareNorm)
        }
        maxNorm = fmax(maxNorm, sqrt(v));
        __syncthreads(); // sNode / A are reused by the next tree
    }
    if (tid == 0) P.precFacAll[i] = 1.0 / maxNorm;
}

} // namespace

void launch_prec_factor(const PrecParams &P, int gridCap, cudaStream_t st) {
    if (gridCap <= 0) return;
    const size_t Kd = (size_t)P.K * P.K * P.K;
    const size_t bytes = (12 * Kd + 4 * (size_t)P.K * P.K) * sizeof(double);
    static size_t configured = 0;
    if (bytes > configured) {
        MRX_CUDA(cudaFuncSetAttribute(prec_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    prec_factor_kernel<<<gridCap, kPrecThreads, bytes, st>>>(P);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
