// mrcpp::apply on the GPU: host-driven refinement loop + device kernels.
//
// Reference control flow replaced (file:line relative to the MRCPP tree):
//   apply<3,double>(prec,out,oper,inp,maxIter,absPrec)       src/treebuilders/apply.cpp:68-93
//   TreeBuilder::build                                        src/treebuilders/TreeBuilder.cpp:38-86
//   ConvolutionCalculator::{initBandSizes,makeOperBand,calcNode,applyOperComp,applyOperator,
//                           tensorApplyOperComp}              src/treebuilders/ConvolutionCalculator.cpp:105-382
//   WaveletAdaptor::splitNode / tree_utils::split_check       src/treebuilders/WaveletAdaptor.h:51-54, tree_utils.cpp:47-65
//   TreeAdaptor::splitNodeVector                              src/treebuilders/TreeAdaptor.h:41-54
//
// Division of labour. The host keeps only topology (which node exists where) and the scalar
// bookkeeping whose summation order defines the reference's thresholds (sNorm/wNorm in work-vector
// order, split decisions). Everything that touches coefficients runs on the device:
//   1. generated input nodes are materialised level by level (kernels.cu, MODE 2);
//   2. one CTA per output node screens every (input node, term) pair with the reference's exact
//      predicate (integer band tests + the FP64 norm product in the reference's operation order) and
//      contracts the surviving (ft, gt, term) tuples; warp w owns output component gt = w, so the
//      accumulation order is fixed and no atomics are needed;
//   3. component norms come back (8 doubles per node) for the split decision.
#include <algorithm>
#include <chrono>
#include <cstring>

#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace mrx {

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ConvolutionCalculator::initBandSizes / calcBandSizeFactor (ConvolutionCalculator.cpp:105-139), including the
// quirk that a negative width does not prevent the final assignment (the `continue` only skips the product).
void band_size_factors(const Operator &op, int DM, std::vector<int> &bsf, std::vector<int> &bw) {
    const int M = op.size();
    bsf.assign((size_t)M * DM * 64, 0);
    bw.assign((size_t)M * DM * 5, -1);
    for (int i = 0; i < M; i++) {
        const OperTerm &ot = op.terms[i];
        for (int depth = 0; depth < DM; depth++) {
            for (int c = 0; c < 5; c++) bw[((size_t)i * DM + depth) * 5 + c] = ot.width(depth, c);
            for (int gt = 0; gt < 8; gt++)
                for (int ft = 0; ft < 8; ft++) {
                    int totNodes = 1;
                    for (int d = 0; d < 3; d++) {
                        int oIdx = 2 * ((gt >> d) & 1) + ((ft >> d) & 1);
                        int width = ot.width(depth, oIdx);
                        if (width < 0) continue;
                        totNodes *= 2 * width + 1;
                    }
                    bsf[((size_t)i * DM + depth) * 64 + gt * 8 + ft] = totNodes * 64;
                }
        }
    }
}

struct Scratch {
    DevBuf<GDesc> gdesc;
    DevBuf<int> nbr;
    DevBuf<int> genItems;
    DevBuf<int> gslots;
    DevBuf<unsigned long long> counters;
};

} // namespace

void device_apply(double prec, mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int maxIter, bool absPrec,
                  mrx_apply_stats *stats) {
    require_device("device_apply");
    cudaStream_t st = stream();
    mrx_apply_stats S{};
    long long launches0 = launch_counter();
    double t0 = now_ms();

    // ---- residency: input tree + operator tables in HBM
    if (!inp.devValid) tree_upload(inp);
    oper_upload(oper);
    Operator &op = oper.op;
    op.calcBandWidths(prec);
    const int M = op.size(), DM = oper.dev.DM;
    {
        std::vector<int> bsf, bw;
        band_size_factors(op, DM, bsf, bw);
        oper.dev.bsf.reserve(bsf.size(), false, st);
        oper.dev.bw.reserve(bw.size(), false, st);
        MRX_CUDA(cudaMemcpyAsync(oper.dev.bsf.p, bsf.data(), sizeof(int) * bsf.size(), cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(oper.dev.bw.p, bw.data(), sizeof(int) * bw.size(), cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaStreamSynchronize(st));
    }
    S.ms_upload = now_ms() - t0;

    Tree<3> &g = out.host;
    Tree<3> &f = inp.host;
    const int K = g.K, Kd = g.Kd, ncoef = g.ncoef;
    const int maxScale = g.mra.maxScale();
    g.allocCoefs = false; // the output lives in HBM until somebody asks for it
    out.hostCoefsValid = false;
    out.devValid = true;
    const double *filt = device_filters(g.k);

    Scratch scr;
    scr.counters.reserve(4, false, st);
    MRX_CUDA(cudaMemsetAsync(scr.counters.p, 0, 4 * sizeof(unsigned long long), st));

    cudaEvent_t ev0, ev1;
    MRX_CUDA(cudaEventCreate(&ev0));
    MRX_CUDA(cudaEventCreate(&ev1));
    float kernel_ms = 0.f;

    double tb = now_ms();
    std::vector<int> workVec;
    g.nodeTable(workVec); // getInitialWorkVector: ALL nodes of `out` (ConvolutionCalculator.cpp:400-405)
    double sNorm = 0.0, wNorm = 0.0;
    int iter = 0;
    const int fRealN = f.nReal;
    std::vector<GDesc> gdesc;
    std::vector<int> nbr;
    std::vector<int> newParents;
    std::vector<double> normsHost;
    while (!workVec.empty()) {
        const int nG = (int)workVec.size();
        // ---- band enumeration (makeOperBand/fillOperBand, :142-222, non-periodic) on the topology
        gdesc.resize(nG);
        nbr.clear();
        newParents.clear();
        for (int i = 0; i < nG; i++) {
            const auto &nd = g.nodes[workVec[i]];
            GDesc &d = gdesc[i];
            d.slot = workVec[i];
            d.depth = nd.scale - op.operRoot;
            d.nbrOff = (int)nbr.size();
            for (int x = 0; x < 3; x++) {
                d.l[x] = nd.l[x];
                d.s[x] = 0;
                d.nb[x] = 0;
            }
            int width = op.getMaxBandWidth(d.depth);
            if (width < 0) continue;
            for (int x = 0; x < 3; x++) {
                int sI = nd.l[x] - width, eI = nd.l[x] + width;
                int nboxes = f.mra.nboxes[x] * (1 << d.depth);
                int c_i = f.mra.corner[x] * (1 << d.depth);
                if (sI < c_i) sI = c_i;
                if (eI > c_i + nboxes - 1) eI = c_i + nboxes - 1;
                d.s[x] = sI;
                d.nb[x] = eI - sI + 1;
            }
            for (int z = 0; z < d.nb[2]; z++)
                for (int y = 0; y < d.nb[1]; y++)
                    for (int x = 0; x < d.nb[0]; x++)
                        nbr.push_back(f.getNodeTopo(nd.scale, {d.s[0] + x, d.s[1] + y, d.s[2] + z}, &newParents));
        }
        // ---- generated input nodes: parents in creation order; a parent created this iteration must be
        //      filled before its own children -> waves
        if (!newParents.empty()) {
            int nGenTotal = f.size() - fRealN;
            inp.dev.genCoefs.reserve((size_t)nGenTotal * Kd, true, st);
            inp.dev.genNorms.reserve((size_t)nGenTotal, true, st);
            std::vector<int> items;
            size_t pos = 0;
            while (pos < newParents.size()) {
                // wave = maximal run whose parents do not depend on children created inside the run
                size_t end = pos;
                int firstChildOfWave = f.nodes[newParents[pos]].child0;
                while (end < newParents.size() && newParents[end] < firstChildOfWave) end++;
                items.clear();
                for (size_t q = pos; q < end; q++) {
                    items.push_back(newParents[q]);
                    items.push_back(f.nodes[newParents[q]].child0);
                }
                scr.genItems.reserve(items.size(), false, st);
                MRX_CUDA(cudaMemcpyAsync(scr.genItems.p, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice, st));
                launch_gen_children(inp.dev.coefs.p, inp.dev.genCoefs.p, inp.dev.genNorms.p, fRealN, scr.genItems.p,
                                    (int)(end - pos), K, filt, st);
                MRX_CUDA(cudaStreamSynchronize(st)); // items buffer is reused
                pos = end;
            }
            inp.dev.nGen = nGenTotal;
            S.gen_nodes += 8 * (long long)newParents.size();
        }
        // ---- device storage for the output nodes of this iteration
        out.dev.coefs.reserve((size_t)g.nReal * ncoef, true, st);
        out.dev.norms.reserve((size_t)g.nReal * 8, true, st);
        out.dev.nNodes = g.nReal;

        scr.gdesc.reserve(nG, false, st);
        scr.nbr.reserve(std::max<size_t>(nbr.size(), 1), false, st);
        scr.gslots.reserve(nG, false, st);
        MRX_CUDA(cudaMemcpyAsync(scr.gdesc.p, gdesc.data(), sizeof(GDesc) * nG, cudaMemcpyHostToDevice, st));
        if (!nbr.empty())
            MRX_CUDA(cudaMemcpyAsync(scr.nbr.p, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(scr.gslots.p, workVec.data(), sizeof(int) * nG, cudaMemcpyHostToDevice, st));

        // gThrs (ConvolutionCalculator.cpp:241-248)
        double gThrs = g.squareNorm;
        if (gThrs > 0.0) {
            auto nTerms = static_cast<double>(M);
            double precFac = 1.0;
            gThrs = prec * precFac * std::sqrt(gThrs / nTerms);
        }

        ApplyParams P{};
        P.fReal = inp.dev.coefs.p;
        P.fGen = inp.dev.genCoefs.p;
        P.fNorms = inp.dev.norms.p;
        P.fGenNorms = inp.dev.genNorms.p;
        P.nRealF = fRealN;
        P.gCoefs = out.dev.coefs.p;
        P.gdesc = scr.gdesc.p;
        P.nbr = scr.nbr.p;
        P.mats = oper.dev.mats.p;
        P.onorms = oper.dev.norms.p;
        P.nodeOff = oper.dev.nodeOff.p;
        P.maxTransl = oper.dev.maxTransl.p;
        P.bw = oper.dev.bw.p;
        P.bsf = oper.dev.bsf.p;
        P.M = M;
        P.DM = DM;
        P.K = K;
        P.gThrs = gThrs;
        P.counters = scr.counters.p;
        P.derivDir = -1;

        MRX_CUDA(cudaEventRecord(ev0, st));
        launch_apply(P, nG, st);
        MRX_CUDA(cudaEventRecord(ev1, st));
        // calcNorms of the output nodes (ConvolutionCalculator.cpp:270-272)
        launch_norms(out.dev.coefs.p, out.dev.norms.p, scr.gslots.p, nG, Kd, st);
        normsHost.resize((size_t)nG * 8);
        // gather norms of the work vector: they are scattered by slot -> copy the covering range
        int lo = *std::min_element(workVec.begin(), workVec.end());
        int hi = *std::max_element(workVec.begin(), workVec.end());
        std::vector<double> range((size_t)(hi - lo + 1) * 8);
        MRX_CUDA(cudaMemcpyAsync(range.data(), out.dev.norms.p + (size_t)lo * 8, sizeof(double) * range.size(),
                                 cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        MRX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        kernel_ms += ms;
        for (int i = 0; i < nG; i++) {
            int n = workVec[i];
            double sq = 0.0;
            for (int c = 0; c < 8; c++) {
                double v = range[(size_t)(n - lo) * 8 + c];
                g.cnorm[(size_t)n * 8 + c] = v;
                sq += v * v;
            }
            g.sqn[n] = sq;
            g.nodes[n].flags |= FlagHasCoefs;
        }
        S.g_nodes += nG;

        // ---- TreeBuilder::build norm bookkeeping (TreeBuilder.cpp:56-66), work-vector order
        if (iter == 0) {
            sNorm = 0.0;
            for (int n : workVec) sNorm += g.scalingNorm(n);
        }
        for (int n : workVec) wNorm += g.waveletNorm(n);
        if (sNorm < 0.0 or wNorm < 0.0) g.squareNorm = -1.0;
        else g.squareNorm = sNorm + wNorm;

        // ---- splitNodeVector (TreeAdaptor.h:41-54) with WaveletAdaptor::splitNode
        std::vector<int> newVec;
        if (iter >= maxIter and maxIter >= 0) workVec.clear();
        for (int n : workVec) {
            if (g.isBranch(n)) continue;
            if (g.nodes[n].scale + 2 > maxScale) continue;
            if (split_check(g, n, prec, 1.0, absPrec)) {
                int c0 = g.createChildren(n, false);
                for (int c = 0; c < 8; c++) newVec.push_back(c0 + c);
            }
        }
        workVec.swap(newVec);
        iter++;
    }
    S.iterations = iter;
    S.ms_build = now_ms() - tb;
    S.ms_kernel = kernel_ms;

    // ---- post: TopDown(+=), BottomUp, square norm, cleanup (apply.cpp:81-87)
    double tp = now_ms();
    op.clearBandWidths();
    unsigned long long counters[4];
    MRX_CUDA(cudaMemcpyAsync(counters, scr.counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    S.f_applied = (long long)counters[0];
    out.dev.nNodes = g.nReal;
    device_mw_transform(out, MRX_TOP_DOWN, false);
    device_mw_transform(out, MRX_BOTTOM_UP, true);
    g.calcSquareNorm();
    f.deleteGenerated();
    inp.dev.nGen = 0;
    S.ms_post = now_ms() - tp;
    S.n_nodes_out = g.nReal;
    S.kernel_launches = launch_counter() - launches0;
    MRX_CUDA(cudaEventDestroy(ev0));
    MRX_CUDA(cudaEventDestroy(ev1));
    if (stats) *stats = S;
}

void device_apply_derivative(mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int dir, mrx_apply_stats *stats) {
    (void)out;
    (void)oper;
    (void)inp;
    (void)dir;
    (void)stats;
    MRX_ABORT("device_apply_derivative: not built yet");
}

} // namespace mrx
