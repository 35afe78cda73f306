// Projection of a Gaussian expansion on the device: the input generator of the apply path.
//
// Reference functions restated (file:line relative to the MRCPP tree):
//   project(prec, out, f, maxIter, absPrec)            src/treebuilders/project.cpp:85-104
//   TreeBuilder::build                                 src/treebuilders/TreeBuilder.cpp:38-86
//   ProjectionCalculator::calcNode                     src/treebuilders/ProjectionCalculator.cpp:34-51
//   MWNode::getExpandedChildPts                        src/trees/MWNode.cpp:903-925
//   GaussExp::evalf / GaussFunc::evalf                 src/functions/GaussExp.cpp, GaussFunc.cpp:47-66
//   MWNode::cvTransform(Backward) (interpolating)      src/trees/MWNode.cpp:448-490
//   MWNode::mwTransform(Compression)                   src/trees/MWNode.cpp:557-594
//   WaveletAdaptor::splitNode / split_check            src/treebuilders/WaveletAdaptor.h:51-54, tree_utils.cpp:47-65
//
// The host keeps the topology and the norm bookkeeping (work-vector order); every function value, the
// coefficient/value transform, the in-node compression and the norms are computed on the device, and the
// tree never has host coefficient storage unless somebody downloads it.
#include <cmath>
#include <cstring>
#include <map>

#include "../engine.hpp"
#include "common.cuh"
#include "kernels.cuh"

namespace mrx {

namespace {

#ifdef __CUDACC__ // the kernel and its launcher: CUDA compiler only (the driver below also builds for the CPU mock of tests/cpp/cuda_mock)
// One CTA per work node. Phase 1: ordered list of the terms that are not identically zero on the node's box
// (exp(-q2) underflows to exactly 0 for q2 > 746 in IEEE double, GaussFunc::evalf returns 0.0 * coef * p2 there: leaving
// such a term out of the sum changes nothing; same test as the host generator, tree.cpp project_gaussians).
// Phase 2: thread per quadrature point: sum of the active terms in expansion order, then sqrt(w) per dimension and the
// 2^{-3(n+1)/2} factor of cvTransform(Backward).
__global__ void __launch_bounds__(256) project_eval_kernel(double *__restrict__ coefs, const int *__restrict__ slots,
                                                           const int4 *__restrict__ nodeInfo, int K, GaussTable G,
                                                           const double *__restrict__ roots, const double *__restrict__ sqrtw) {
    extern __shared__ int active[]; // [G.n]
    __shared__ int nActive;
    __shared__ double sRoots[64], sSw[64];
    const int tid = threadIdx.x, lane = tid & 31;
    const int4 ni = nodeInfo[blockIdx.x];
    const int scale = ni.x;
    const int l[3] = {ni.y, ni.z, ni.w};
    const double len = ldexp(1.0, -scale);
    if (tid < K) {
        sRoots[tid] = roots[tid];
        sSw[tid] = sqrtw[tid];
    }
    if (tid < 32) {
        int cnt = 0;
        for (int base = 0; base < G.n; base += 32) {
            const int g = base + lane;
            bool on = false;
            if (g < G.n) {
                double minq2 = 0.0;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double lb = len * l[d], ub = len * (l[d] + 1);
                    const double p = G.pos[3 * g + d];
                    const double dist = (p < lb) ? lb - p : (p > ub ? p - ub : 0.0);
                    minq2 += G.alpha[g] * dist * dist;
                }
                on = !(minq2 > 747.0);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (on) active[cnt + __popc(bal & ((1u << lane) - 1u))] = g;
            cnt += __popc(bal);
        }
        if (lane == 0) nActive = cnt;
    }
    __syncthreads();
    const int nA = nActive;
    const int K2 = K * K, Kd = K2 * K;
    const double sFac = ldexp(1.0, -(scale + 1));
    const int np1 = scale + 1;
    const double two_fac = sqrt(1.0 / ldexp(1.0, 3 * np1));
    double *out = coefs + (size_t)slots[blockIdx.x] * 8 * Kd;
    for (int o = tid; o < 8 * Kd; o += 256) {
        const int tt = o / Kd, idx = o - tt * Kd;
        const int j0 = idx % K, j1 = (idx / K) % K, j2 = idx / K2;
        double r[3];
        r[0] = sFac * (sRoots[j0] + 2.0 * static_cast<double>(l[0]) + ((tt & 1) ? 1.0 : 0.0));
        r[1] = sFac * (sRoots[j1] + 2.0 * static_cast<double>(l[1]) + ((tt & 2) ? 1.0 : 0.0));
        r[2] = sFac * (sRoots[j2] + 2.0 * static_cast<double>(l[2]) + ((tt & 4) ? 1.0 : 0.0));
        double s = 0.0;
        for (int a = 0; a < nA; a++) {
            const int g = active[a];
            const double alpha = G.alpha[g], cf = G.coef[g];
            double q2 = 0.0, p2 = 1.0;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double q = r[d] - G.pos[3 * g + d];
                q2 += alpha * q * q;
                const int pw = G.power[3 * g + d];
                if (pw == 0) continue;
                if (pw == 1) p2 *= q;
                else p2 *= pow(q, (double)pw);
            }
            s += (q2 > 746.0) ? 0.0 * cf * p2 : cf * p2 * exp(-q2);
        }
        double v = s;
        v = v * sSw[j0];
        v = v * sSw[j1];
        v = v * sSw[j2];
        out[o] = two_fac * v;
    }
}

#endif // __CUDACC__

struct QuadDev {
    double *roots = nullptr, *sqrtw = nullptr;
};

const QuadDev &device_quadrature(int K) {
    static std::map<int, QuadDev> cache;
    auto it = cache.find(K);
    if (it != cache.end()) return it->second;
    const Quadrature &q = quadrature(K);
    std::vector<double> sw(K);
    for (int j = 0; j < K; j++) sw[j] = std::sqrt(q.weights[j]);
    QuadDev d;
    MRX_CUDA(cudaMalloc(&d.roots, sizeof(double) * K));
    MRX_CUDA(cudaMalloc(&d.sqrtw, sizeof(double) * K));
    MRX_CUDA(cudaMemcpy(d.roots, q.roots.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
    MRX_CUDA(cudaMemcpy(d.sqrtw, sw.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
    return cache[K] = d;
}

} // namespace

#ifdef __CUDACC__
void launch_project_eval(double *coefs, const int *slots, const int4 *nodeInfo, int cnt, int K, const GaussTable &g, const double *roots,
                         const double *sqrtw, cudaStream_t st) {
    if (cnt <= 0) return;
    if (K > 64) MRX_ABORT("projection kernel: order too large");
    project_eval_kernel<<<cnt, 256, sizeof(int) * std::max(g.n, 1), st>>>(coefs, slots, nodeInfo, K, g, roots, sqrtw);
    MRX_CUDA(cudaGetLastError());
    launch_counter()++;
}
#endif // __CUDACC__

/// project(prec, out, GaussExp) with the tree resident in HBM (the host holds topology + norms only)
void device_project_gaussians(mrx_tree &t, double prec, const GaussExp<3> &gexp, int maxIter, bool absPrec) {
    require_device("device_project_gaussians");
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    const int K = h.K, Kd = h.Kd, ncoef = h.ncoef;
    const int nGauss = (int)gexp.size();
    if ((size_t)nGauss * sizeof(int) > 200 * 1024) MRX_ABORT("device projection: more than 51200 Gaussian terms");
    h.allocCoefs = false;
    t.hostCoefsValid = false;
    // expansion tables
    std::vector<double> hc(nGauss), ha(nGauss), hp((size_t)3 * nGauss);
    std::vector<int> hw((size_t)3 * nGauss);
    for (int i = 0; i < nGauss; i++) {
        hc[i] = gexp[i].coef;
        ha[i] = gexp[i].alpha;
        for (int d = 0; d < 3; d++) {
            hp[3 * (size_t)i + d] = gexp[i].pos[d];
            hw[3 * (size_t)i + d] = gexp[i].power[d];
        }
    }
    DevBuf<double> dc, da, dp;
    DevBuf<int> dw, dSlots, dPairs;
    DevBuf<int4> dInfo;
    DevBuf<double> dNormsW;
    dc.reserve(std::max(nGauss, 1), false, st);
    da.reserve(std::max(nGauss, 1), false, st);
    dp.reserve(std::max(3 * nGauss, 1), false, st);
    dw.reserve(std::max(3 * nGauss, 1), false, st);
    if (nGauss > 0) {
        MRX_CUDA(cudaMemcpyAsync(dc.p, hc.data(), sizeof(double) * nGauss, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(da.p, ha.data(), sizeof(double) * nGauss, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(dp.p, hp.data(), sizeof(double) * 3 * nGauss, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(dw.p, hw.data(), sizeof(int) * 3 * nGauss, cudaMemcpyHostToDevice, st));
    }
    GaussTable G{dc.p, da.p, dp.p, dw.p, nGauss};
#ifdef __CUDACC__
    if (nGauss * (int)sizeof(int) > 48 * 1024) {
        static int configured = 0;
        if (nGauss > configured) {
            MRX_CUDA(cudaFuncSetAttribute(project_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nGauss * (int)sizeof(int)));
            configured = nGauss;
        }
    }
#endif

    const QuadDev &Q = device_quadrature(K);
    const double *filt = device_filters(h.k);

    // TreeBuilder::build (TreeBuilder.cpp:38-86): initial work vector = end nodes of the grid (TreeCalculator.h:37)
    std::vector<int> workVec;
    h.endNodeTable(workVec);
    // branch nodes of a pre-built grid are never computed by the projection: the closing BottomUp pass overwrites them, but
    // their storage must exist (and be defined) on the device
    t.dev.coefs.reserve((size_t)h.nReal * ncoef, false, st);
    t.dev.norms.reserve((size_t)h.nReal * 8, false, st);
    MRX_CUDA(cudaMemsetAsync(t.dev.coefs.p, 0, sizeof(double) * (size_t)h.nReal * ncoef, st));
    double sNorm = 0.0, wNorm = 0.0;
    int iter = 0;
    const int maxScale = h.mra.maxScale();
    std::vector<int4> info;
    std::vector<int> pairs;
    std::vector<double> nrm;
    while (!workVec.empty()) {
        const int nW = (int)workVec.size();
        info.resize(nW);
        pairs.resize((size_t)2 * nW);
        for (int i = 0; i < nW; i++) {
            const auto &nd = h.nodes[workVec[i]];
            info[i] = make_int4(nd.scale, nd.l[0], nd.l[1], nd.l[2]);
            pairs[2 * (size_t)i] = workVec[i];
            pairs[2 * (size_t)i + 1] = -1;
        }
        t.dev.coefs.reserve((size_t)h.nReal * ncoef, true, st);
        t.dev.norms.reserve((size_t)h.nReal * 8, true, st);
        dSlots.reserve(nW, false, st);
        dPairs.reserve((size_t)2 * nW, false, st);
        dInfo.reserve(nW, false, st);
        dNormsW.reserve((size_t)nW * 8, false, st);
        MRX_CUDA(cudaMemcpyAsync(dSlots.p, workVec.data(), sizeof(int) * nW, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(dPairs.p, pairs.data(), sizeof(int) * 2 * nW, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(dInfo.p, info.data(), sizeof(int4) * nW, cudaMemcpyHostToDevice, st));
        launch_project_eval(t.dev.coefs.p, dSlots.p, dInfo.p, nW, K, G, Q.roots, Q.sqrtw, st);
        launch_compress_nodes(t.dev.coefs.p, dPairs.p, nW, K, filt, st);
        launch_norms(t.dev.coefs.p, t.dev.norms.p, dSlots.p, nW, Kd, st, dNormsW.p);
        nrm.resize((size_t)nW * 8);
        MRX_CUDA(cudaMemcpyAsync(nrm.data(), dNormsW.p, sizeof(double) * nrm.size(), cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < nW; i++) {
            const int n = workVec[i];
            double sq = 0.0;
            for (int c = 0; c < 8; c++) {
                const double v = nrm[(size_t)i * 8 + c];
                h.cnorm[(size_t)n * 8 + c] = v;
                sq += v * v;
            }
            h.sqn[n] = sq;
            h.nodes[n].flags |= FlagHasCoefs;
        }
        if (iter == 0) {
            sNorm = 0.0;
            for (int n : workVec) sNorm += h.scalingNorm(n);
        }
        for (int n : workVec) wNorm += h.waveletNorm(n);
        if (sNorm < 0.0 or wNorm < 0.0) h.squareNorm = -1.0;
        else h.squareNorm = sNorm + wNorm;
        std::vector<int> newVec;
        if (iter >= maxIter and maxIter >= 0) workVec.clear();
        for (int n : workVec) {
            if (h.isBranch(n)) continue;
            if (h.nodes[n].scale + 2 > maxScale) continue;
            if (split_check(h, n, prec, 1.0, absPrec)) {
                const int c0 = h.createChildren(n, false);
                for (int c = 0; c < 8; c++) newVec.push_back(c0 + c);
            }
        }
        workVec.swap(newVec);
        iter++;
    }
    t.dev.nNodes = h.nReal;
    t.dev.nGen = 0;
    t.devValid = true;
    t.hostCoefsValid = false;
    // project.cpp:96-97: out.mwTransform(BottomUp); out.calcSquareNorm()
    device_mw_transform(t, MRX_BOTTOM_UP, true);
    h.calcSquareNorm();
}

} // namespace mrx
