// Refinement step of TreeBuilder::build on the device: the norm bookkeeping, the split decisions and the next work
// vector, so that the host's critical path per iteration is three small read-backs instead of loops over the nodes.
//
// Reference logic restated (file:line relative to the MRCPP tree):
//   TreeBuilder::build, norm bookkeeping              src/treebuilders/TreeBuilder.cpp:56-66
//   TreeAdaptor::splitNodeVector                      src/treebuilders/TreeAdaptor.h:41-54
//   WaveletAdaptor::splitNode / tree_utils::split_check   WaveletAdaptor.h:51-54, src/utils/tree_utils.cpp:47-65
//   MWNode::getScalingNorm / getWaveletNorm           src/trees/MWNode.cpp:619-640
//   FunctionNode::createChildren (child order, 2 l + bit)   src/trees/FunctionNode.cpp:255-291, NodeIndex.h:49-53
//
// sNorm / wNorm are accumulated in work-vector order by ONE thread (the reference's summation order; the order defines
// the thresholds of every later iteration, and it must not depend on the number of ranks): 44 K additions per apply.
#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"

namespace mrx {

namespace {

__device__ __forceinline__ int cta_excl_scan_int(int v, int *sm, int &total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int s = sm[lane];
        int si = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, si, off);
            if (lane >= off) si += t;
        }
        sm[lane] = si - s;
        if (lane == 31) sm[32] = si;
    }
    __syncthreads();
    const int res = sm[warp] + incl - v;
    total = sm[32];
    __syncthreads();
    return res;
}

// one CTA of 1024 threads
__global__ void __launch_bounds__(1024) split_kernel(SplitParams S) {
    __shared__ double sW[2][1024], sS[2][1024];
    __shared__ int sm[33];
    __shared__ double shSq;
    const int tid = threadIdx.x;
    const int nG = S.nG;
    // ---- norm bookkeeping in work-vector order: ONE thread adds (the reference's summation order, TreeBuilder.cpp:56-66), the
    //      other threads stage the next 1024 items meanwhile (double buffer), so only the dependent additions are serial
    double sNorm = S.state[0], wNorm = S.state[1];
    if (S.iter == 0) sNorm = 0.0;
    auto stage = [&](int buf, int i) {
        if (i < nG) {
            const double *n = S.normRows + (size_t)shard_row(i, S.world, S.rows, S.shardB) * 8;
            double w = 0.0;
#pragma unroll
            for (int c = 1; c < 8; c++) w = fma(n[c], n[c], w); // MWNode::getWaveletNorm (squared), component order
            sW[buf][i & 1023] = w;
            sS[buf][i & 1023] = n[0] * n[0];
        }
    };
    stage(0, tid); // chunk 0
    const int nChunks = (nG + 1023) / 1024;
    for (int c = 0; c < nChunks; c++) {
        __syncthreads(); // chunk c is staged; thread 0 is done with chunk c - 1, whose buffer chunk c + 1 overwrites
        const int buf = c & 1, base = c * 1024;
        if (tid == 0) {
            const int m = min(1024, nG - base);
            if (S.iter == 0) {
                int k = 0;
                for (; k + 8 <= m; k += 8) {
                    double v[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) v[q] = sS[buf][k + q];
#pragma unroll
                    for (int q = 0; q < 8; q++) sNorm += v[q];
                }
                for (; k < m; k++) sNorm += sS[buf][k];
            }
            int k = 0;
            for (; k + 8 <= m; k += 8) {
                double v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) v[q] = sW[buf][k + q];
#pragma unroll
                for (int q = 0; q < 8; q++) wNorm += v[q];
            }
            for (; k < m; k++) wNorm += sW[buf][k];
        } else if (c + 1 < nChunks) {
            stage(buf ^ 1, base + 1024 + tid);
            if (tid == 1) stage(buf ^ 1, base + 1024); // thread 0's item
        }
    }
    __syncthreads();
    if (tid == 0) {
        const double sq = (sNorm < 0.0 || wNorm < 0.0) ? -1.0 : sNorm + wNorm;
        S.state[0] = sNorm;
        S.state[1] = wNorm;
        S.state[2] = sq;
        shSq = sq;
        S.res->squareNorm = sq;
        S.res->sNorm = sNorm;
        S.res->wNorm = wNorm;
    }
    __syncthreads();
    int nSplit = 0;
    if (S.doSplit) {
        // ---- split_check + positions of the children in the next work vector
        const double sq = shSq;
        double t_norm = 1.0;
        if (sq > 0.0 && !S.absPrec) t_norm = sqrt(sq);
        int run = 0;
        for (int base = 0; base < nG; base += 1024) {
            const int i = base + tid;
            int flag = 0;
            int4 gn = make_int4(0, 0, 0, 0);
            if (i < nG) {
                gn = S.gNodesAll[i];
                const int scale = gn.x + S.operRoot;
                const bool branch = S.isBranch ? (S.isBranch[i] != 0) : false;
                if (!branch && scale + 2 <= S.maxScale && S.prec > 0.0) {
                    const double *n = S.normRows + (size_t)shard_row(i, S.world, S.rows, S.shardB) * 8;
                    double w = 0.0;
#pragma unroll
                    for (int c = 1; c < 8; c++) w = fma(n[c], n[c], w);
                    // WaveletAdaptor::splitNode: prec * precFac (1.0 without precision trees), then split_check's product
                    const double precN = S.precFacAll ? S.prec * S.precFacAll[i] : S.prec;
                    const double thr = fmax(2.0 * 1.0e-15, precN * t_norm * S.scaleFac[scale - S.rootScale]);
                    if (sqrt(w) > thr) flag = 1;
                }
                S.flags[i] = (unsigned char)flag;
            }
            int total;
            const int pos = run + cta_excl_scan_int(flag, sm, total);
            if (flag) {
                if (S.pairsNext) {
                    S.pairsNext[2 * pos] = S.slotsCur[i];
                    S.pairsNext[2 * pos + 1] = S.slotBase + 8 * pos;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int q = 8 * pos + c;
                    S.gNodesNext[q] = make_int4(gn.x + 1, 2 * gn.y + (c & 1), 2 * gn.z + ((c >> 1) & 1), 2 * gn.w + ((c >> 2) & 1));
                    S.slotsNext[q] = S.slotBase + q;
                }
            }
            run += total;
        }
        nSplit = run;
    }
    if (tid == 0) {
        S.res->nSplit = nSplit;
        S.res->nNext = 8 * nSplit;
    }
}

// one CTA: the rank's share of a work vector (block-cyclic distribution) and the chunk table of its band enumeration
__global__ void __launch_bounds__(1024) prep_local_kernel(PrepParams P) {
    __shared__ int sm[33];
    const int tid = threadIdx.x;
    const int nG = (P.nG >= 0) ? P.nG : P.res->nNext;
    const int nL = shard_count(nG, P.world, P.rank, P.shardB);
    int run = 0;
    long long nbr = 0;
    for (int base = 0; base < nL; base += 1024) {
        const int j = base + tid;
        int cnt = 0;
        if (j < nL) {
            const int i = shard_item(P.rank, j, P.world, P.shardB);
            const int4 gn = P.gNodesAll[i];
            P.gNodesLoc[j] = gn;
            P.slotsLoc[j] = P.slotsAll[i];
            if (P.precAll) P.precLoc[j] = P.precAll[i];
            const int dep = gn.x;
            // deeper than every operator tree, or no band at that depth: empty band (ConvolutionCalculator.cpp:146-151)
            if (dep >= 0 && dep < P.DM && P.depthInfo[dep].W >= 0) {
                cnt = (P.offCount[dep] + 31) / 32;
                nbr += P.offCount[dep];
            }
        }
        int total;
        const int off = run + cta_excl_scan_int(cnt, sm, total);
        if (j < nL) P.chunkOffLoc[j] = off;
        run += total;
    }
    // total neighbour capacity: block reduction
    __shared__ long long red[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nbr += __shfl_xor_sync(0xffffffffu, nbr, off);
    if ((tid & 31) == 0) red[tid >> 5] = nbr;
    __syncthreads();
    if (tid == 0) {
        long long t = 0;
        for (int w = 0; w < 32; w++) t += red[w];
        P.chunkOffLoc[nL] = run;
        P.res->nChunksLoc = run;
        P.res->nbrCapLoc = t;
        P.res->nLoc = nL;
    }
}

__global__ void publish_kernel(const unsigned *__restrict__ src, volatile unsigned *hostDst, int words, volatile unsigned *hostFlag,
                               unsigned seq) {
    if (threadIdx.x != 0) return;
    for (int i = 0; i < words; i++) hostDst[i] = src[i]; // a few words: one thread, one ordered sequence of stores
    __threadfence_system();
    *hostFlag = seq;
}

} // namespace

void launch_publish(const void *src, void *hostDst, int words, unsigned *hostFlag, unsigned seq, cudaStream_t st) {
    publish_kernel<<<1, 32, 0, st>>>(static_cast<const unsigned *>(src), static_cast<volatile unsigned *>(hostDst), words, hostFlag, seq);
    MRX_CUDA(cudaGetLastError());
}

void launch_split(const SplitParams &S, cudaStream_t st) {
    split_kernel<<<1, 1024, 0, st>>>(S);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

void launch_prep_local(const PrepParams &P, cudaStream_t st) {
    prep_local_kernel<<<1, 1024, 0, st>>>(P);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}

} // namespace mrx
