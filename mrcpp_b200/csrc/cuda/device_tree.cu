// HBM-resident node store: upload/download, whole-tree two-scale transforms (level-synchronous),
// component norms, dot product, rescale, operator-table packing.
//
// Reference functions replaced here (file:line relative to the MRCPP tree):
//   MWTree::mwTransformDown / mwTransformUp          src/trees/MWTree.cpp:166-216
//   MWNode::giveChildrenCoefs / reCompress<3>        src/trees/MWNode.cpp:312-335, FunctionNode.cpp:391-410
//   tree_utils::mw_transform / mw_transform_back     src/utils/tree_utils.cpp:113-301
//   math_utils::apply_filter                         src/utils/math_utils.cpp:175-194
//   MWNode::calcNorms / calcComponentNorm            src/trees/MWNode.cpp:609-616, :643-655
#include <algorithm>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

#include "../engine.hpp"
#include "common.cuh"
#include "kernels.cuh"

namespace mrx {

// ---- filter tables per order, device resident: [op][idx][K*K] row-major F(i,j)
const double *device_filters(int k) {
    static std::map<int, double *> cache;
    auto it = cache.find(k);
    if (it != cache.end()) return it->second;
    const FilterSet &fs = filter_set(k);
    int K = k + 1;
    std::vector<double> h((size_t)8 * K * K);
    for (int op = 0; op < 2; op++)
        for (int i = 0; i < 4; i++) std::memcpy(h.data() + (size_t)(op * 4 + i) * K * K, fs.sub[op][i].data(), sizeof(double) * K * K);
    double *d = nullptr;
    MRX_CUDA(cudaMalloc(&d, h.size() * sizeof(double)));
    MRX_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    cache[k] = d;
    return d;
}

double *pinned_stage(size_t doubles) {
    static double *buf = nullptr;
    static size_t cap = 0;
    if (doubles > cap) {
        if (buf) cudaFreeHost(buf);
        cap = std::max(doubles, 2 * cap);
        MRX_CUDA(cudaMallocHost(&buf, cap * sizeof(double)));
    }
    return buf;
}

void host_parallel(size_t n, const std::function<void(size_t, size_t)> &fn) {
    static const unsigned maxT = [] {
        const char *e = getenv("MRX_HOST_THREADS");
        const int t = e ? atoi(e) : 4;
        return (unsigned)std::max(1, std::min(t, (int)std::max(1u, std::thread::hardware_concurrency())));
    }();
    const unsigned T = (n < ((size_t)1 << 15)) ? 1u : maxT;
    if (T == 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (n + T - 1) / T;
    for (unsigned t = 1; t < T; t++) {
        const size_t a = std::min(n, t * per), b = std::min(n, a + per);
        if (a < b) th.emplace_back([&fn, a, b] { fn(a, b); });
    }
    fn(0, std::min(n, per));
    for (auto &x : th) x.join();
}

void tree_upload(mrx_tree &t) {
    require_device("tree_upload");
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    int n = h.nReal;
    t.dev.coefs.reserve((size_t)n * h.ncoef, false, st);
    t.dev.norms.reserve((size_t)n * 8, false, st);
    // host chunks are 64 nodes each
    for (int s = 0; s < n; s += 64) {
        int cnt = std::min(64, n - s);
        MRX_CUDA(cudaMemcpyAsync(t.dev.coefs.p + (size_t)s * h.ncoef, h.coef(s), sizeof(double) * (size_t)cnt * h.ncoef,
                                 cudaMemcpyHostToDevice, st));
    }
    MRX_CUDA(cudaMemcpyAsync(t.dev.norms.p, h.cnorm.data(), sizeof(double) * (size_t)n * 8, cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    t.dev.nNodes = n;
    t.dev.nGen = 0;
    t.dev.topoNodes = -1;
    t.dev.partial = false;
    t.devValid = true;
}

void tree_lazy_begin(mrx_tree &t) {
    require_device("tree_lazy_begin");
    if (t.dev.partial) return; // keep what earlier applies already fetched
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    const int n = h.nReal;
    t.dev.coefs.reserve((size_t)n * h.ncoef, false, st);
    t.dev.norms.reserve((size_t)n * 8, false, st);
    t.dev.resident.reserve((size_t)8 * std::max(n, 1), false, st); // one flag per coefficient block
    const auto &chunks = h.coefChunks();
    t.dev.chunkTab.reserve(std::max<size_t>(chunks.size(), 1), false, st);
    // component norms: staged into pinned memory on a few threads, then one DMA transfer
    double *stage = pinned_stage((size_t)n * 8);
    const double *cn = h.cnorm.data();
    host_parallel((size_t)n * 8, [&](size_t a, size_t b) { std::memcpy(stage + a, cn + a, (b - a) * sizeof(double)); });
    MRX_CUDA(cudaMemcpyAsync(t.dev.norms.p, stage, sizeof(double) * (size_t)n * 8, cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaMemsetAsync(t.dev.resident.p, 0, sizeof(int) * (size_t)8 * std::max(n, 1), st));
    MRX_CUDA(cudaMemcpyAsync(t.dev.chunkTab.p, chunks.data(), sizeof(double *) * chunks.size(), cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    t.dev.nNodes = n;
    t.dev.nGen = 0;
    t.dev.partial = true;
    t.devValid = false;
}

void tree_download(mrx_tree &t) {
    require_device("tree_download");
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    int n = h.nReal;
    if (!t.devValid || t.dev.nNodes < n) MRX_ABORT("tree_download: device copy is not current");
    h.ensureCoefStorage();
    for (int s = 0; s < n; s += 64) {
        int cnt = std::min(64, n - s);
        MRX_CUDA(cudaMemcpyAsync(h.coef(s), t.dev.coefs.p + (size_t)s * h.ncoef, sizeof(double) * (size_t)cnt * h.ncoef,
                                 cudaMemcpyDeviceToHost, st));
    }
    MRX_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) h.nodes[i].flags |= FlagHasCoefs;
    t.hostCoefsValid = true;
}

void tree_drop_device(mrx_tree &t) {
    t.dev.coefs.release();
    t.dev.norms.release();
    t.dev.genCoefs.release();
    t.dev.genNorms.release();
    t.dev.topoChild0.release();
    t.dev.topoDepth.release();
    t.dev.topoBound.release();
    t.dev.topoNodes = -1;
    t.dev.resident.release();
    t.dev.chunkTab.release();
    t.dev.partial = false;
    t.dev.nNodes = 0;
    t.dev.nGen = 0;
    t.devValid = false;
}

// norms of every real node on the device -> host cnorm / sqn (MWNode::calcNorms)
void device_calc_norms_all(mrx_tree &t) {
    require_device("device_calc_norms_all");
    if (!t.devValid) tree_upload(t);
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    int n = h.nReal;
    t.dev.topoNodes = -1; // node norms change: the cached band-walk topology of this tree is stale
    launch_norms(t.dev.coefs.p, t.dev.norms.p, nullptr, n, h.Kd, st);
    MRX_CUDA(cudaMemcpyAsync(h.cnorm.data(), t.dev.norms.p, sizeof(double) * (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < n; i++) {
        double sq = 0.0;
        for (int c = 0; c < 8; c++) sq += h.cnorm[(size_t)i * 8 + c] * h.cnorm[(size_t)i * 8 + c];
        h.sqn[i] = sq;
    }
}

// level lists of branch nodes (counting sort by depth over the slot order; parents of one level are independent, so
// their order inside a level is irrelevant): flat = (parent, child0) pairs, level d = [levelOff[d], levelOff[d+1])
static int build_level_pairs(const Tree<3> &h, std::vector<int> &levelOff, std::vector<int> &flat) {
    const int n = h.nReal;
    levelOff.assign(2, 0);
    for (int i = 0; i < n; i++) {
        const auto &nd = h.nodes[i];
        if (nd.child0 < 0 || nd.child0 >= n) continue; // leaf, or children are generated nodes
        const int d = nd.scale - h.mra.rootScale;
        if (d + 2 > (int)levelOff.size()) levelOff.resize(d + 2, 0);
        levelOff[d + 1]++;
    }
    const int nLevels = (int)levelOff.size() - 1;
    for (int d = 0; d < nLevels; d++) levelOff[d + 1] += levelOff[d];
    const int nPairs = levelOff[nLevels];
    flat.resize((size_t)2 * nPairs);
    std::vector<int> fill(levelOff.begin(), levelOff.end() - 1);
    for (int i = 0; i < n; i++) {
        const auto &nd = h.nodes[i];
        if (nd.child0 < 0 || nd.child0 >= n) continue;
        const int pos = fill[nd.scale - h.mra.rootScale]++;
        flat[2 * (size_t)pos] = i;
        flat[2 * (size_t)pos + 1] = nd.child0;
    }
    return nPairs;
}

/// closing passes of mrcpp::apply (apply.cpp:82-84): mwTransform(TopDown, overwrite = false), mwTransform(BottomUp),
/// calcSquareNorm. One level list for both passes; where the transform kernels compute the norms of what they write
/// (scaling norm of every child on the way down, all eight norms of every parent on the way up) no separate pass over the
/// tree is needed: wavelet norms of the leaves are those of the apply itself. The tree norm is the sum of the end-node
/// square norms in slot order (the reference sums the same terms in end-node-table order).
// pinned bounce buffer for small device->host results (a pageable destination is staged by the driver at a fraction of
// the PCIe rate)
static double *pinned_scratch(size_t doubles) {
    static double *buf = nullptr;
    static size_t cap = 0;
    if (doubles > cap) {
        if (buf) cudaFreeHost(buf);
        cap = std::max(doubles, 2 * cap);
        MRX_CUDA(cudaMallocHost(&buf, cap * sizeof(double)));
    }
    return buf;
}

void device_apply_post(mrx_tree &t, const std::vector<std::vector<int>> *pairsByDepth, bool topDownDone) {
    require_device("device_apply_post");
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    const int n = h.nReal;
    std::vector<int> levelOff, flat;
    int nPairs = 0;
    if (pairsByDepth) { // collected by the apply while it replayed the split decisions (already grouped by depth)
        levelOff.assign(1, 0);
        for (const auto &lv : *pairsByDepth) {
            flat.insert(flat.end(), lv.begin(), lv.end());
            levelOff.push_back((int)flat.size() / 2);
        }
        if (levelOff.size() == 1) levelOff.push_back(0);
        nPairs = (int)flat.size() / 2;
    } else {
        nPairs = build_level_pairs(h, levelOff, flat);
    }
    const int nLevels = (int)levelOff.size() - 1;
    const bool fused = transform_fuses_norms(h.K);
    t.dev.topoNodes = -1;
    DevBuf<int> pairs;
    if (nPairs > 0) {
        pairs.reserve(flat.size(), false, st);
        MRX_CUDA(cudaMemcpyAsync(pairs.p, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice, st));
        const double *filt = device_filters(h.k);
        double *nrm = fused ? t.dev.norms.p : nullptr;
        if (!topDownDone)
            for (int d = 0; d < nLevels; d++) {
                const int cnt = levelOff[d + 1] - levelOff[d];
                if (cnt > 0) launch_transform(true, false, t.dev.coefs.p, pairs.p + 2 * (size_t)levelOff[d], cnt, h.K, filt, st, nrm);
            }
        for (int d = nLevels - 1; d >= 0; d--) {
            const int cnt = levelOff[d + 1] - levelOff[d];
            if (cnt > 0) launch_transform(false, true, t.dev.coefs.p, pairs.p + 2 * (size_t)levelOff[d], cnt, h.K, filt, st, nrm);
        }
        if (!fused) launch_norms(t.dev.coefs.p, t.dev.norms.p, nullptr, n, h.Kd, st);
    }
    double *bounce = pinned_scratch((size_t)n * 8);
    MRX_CUDA(cudaMemcpyAsync(bounce, t.dev.norms.p, sizeof(double) * (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st)); // also keeps `pairs` alive until the launches have consumed it
    t.devValid = true;
    t.hostCoefsValid = false;
    double tot = 0.0;
    double *cn = h.cnorm.data();
    for (int i = 0; i < n; i++) {
        double sq = 0.0;
        for (int c = 0; c < 8; c++) {
            const double v = bounce[(size_t)i * 8 + c];
            cn[(size_t)i * 8 + c] = v;
            sq += v * v;
        }
        h.sqn[i] = sq;
        if (h.nodes[i].flags & FlagEnd) tot += sq;
    }
    h.squareNorm = tot;
}

void device_mw_transform(mrx_tree &t, int type, bool overwrite, bool norms, int timedReps, double *timedMs, int *branchNodes) {
    require_device("device_mw_transform");
    if (!t.devValid) tree_upload(t);
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    std::vector<int> levelOff, flat;
    const int nPairs = build_level_pairs(h, levelOff, flat);
    const int nLevels = (int)levelOff.size() - 1;
    if (nPairs == 0) {
        if (norms) device_calc_norms_all(t);
        return;
    }
    DevBuf<int> pairs;
    pairs.reserve(flat.size(), false, st);
    MRX_CUDA(cudaMemcpyAsync(pairs.p, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice, st));
    const double *filt = device_filters(h.k);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timedReps > 0) {
        MRX_CUDA(cudaEventCreate(&e0));
        MRX_CUDA(cudaEventCreate(&e1));
        MRX_CUDA(cudaEventRecord(e0, st));
    }
    for (int rep = 0; rep < std::max(timedReps, 1); rep++) {
        if (type == MRX_TOP_DOWN) {
            for (int d = 0; d < nLevels; d++) {
                int cnt = levelOff[d + 1] - levelOff[d];
                if (cnt > 0)
                    launch_transform(true, overwrite, t.dev.coefs.p, pairs.p + 2 * (size_t)levelOff[d], cnt, h.K, filt, st);
            }
        } else {
            for (int d = nLevels - 1; d >= 0; d--) {
                int cnt = levelOff[d + 1] - levelOff[d];
                if (cnt > 0)
                    launch_transform(false, true, t.dev.coefs.p, pairs.p + 2 * (size_t)levelOff[d], cnt, h.K, filt, st);
            }
        }
    }
    if (timedReps > 0) {
        MRX_CUDA(cudaEventRecord(e1, st));
        MRX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        MRX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        MRX_CUDA(cudaEventDestroy(e0));
        MRX_CUDA(cudaEventDestroy(e1));
        if (timedMs) *timedMs = ms / timedReps;
        if (branchNodes) *branchNodes = nPairs;
    }
    t.devValid = true;
    t.hostCoefsValid = false;
    if (norms) device_calc_norms_all(t); // synchronises
}

double device_dot(mrx_tree &bra, mrx_tree &ket) {
    if (!bra.devValid) tree_upload(bra);
    if (!ket.devValid) tree_upload(ket);
    const Tree<3> &a = bra.host, &b = ket.host;
    cudaStream_t st = stream();
    // pairs of nodes present in both trees (multiply.cpp:286-318 walks the same intersection)
    std::vector<int> pairs;
    std::vector<std::pair<int, int>> stack;
    for (int r = a.nRoots - 1; r >= 0; r--) stack.push_back({r, r});
    while (!stack.empty()) {
        auto pr = stack.back();
        stack.pop_back();
        pairs.push_back(pr.first);
        pairs.push_back(pr.second);
        bool aB = a.isBranch(pr.first) && !a.isGen(a.nodes[pr.first].child0);
        bool bB = b.isBranch(pr.second) && !b.isGen(b.nodes[pr.second].child0);
        if (aB && bB)
            for (int c = 7; c >= 0; c--) stack.push_back({a.nodes[pr.first].child0 + c, b.nodes[pr.second].child0 + c});
    }
    int np = (int)pairs.size() / 2;
    DevBuf<int> dpairs;
    DevBuf<double> dres;
    dpairs.reserve(pairs.size(), false, st);
    dres.reserve(np, false, st);
    MRX_CUDA(cudaMemcpyAsync(dpairs.p, pairs.data(), sizeof(int) * pairs.size(), cudaMemcpyHostToDevice, st));
    launch_dot(bra.dev.coefs.p, ket.dev.coefs.p, dpairs.p, dres.p, np, a.nRoots, a.Kd, st);
    std::vector<double> res(np);
    MRX_CUDA(cudaMemcpyAsync(res.data(), dres.p, sizeof(double) * np, cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    double s = 0.0;
    for (double v : res) s += v;
    return s;
}

// add(prec < 0 | maxIter = 0, out, {(c_i, inp_i)}) on the grid `out` enters with (src/treebuilders/add.cpp:41-70,
// AdditionCalculator.h:42-66), restated for the compressed representation: the sum is linear, so its wavelet blocks are the
// sums of the inputs' wavelet blocks on the nodes they share with `out` (zero where an input is coarser: generated nodes
// have no wavelet part), its root scaling blocks the sums of the root scaling blocks, and every other scaling block follows
// from those by reconstruction -- one TopDown(+=) pass over the zero-initialised scaling blocks. That is what the reference's
// per-end-node sum followed by BottomUp produces (its branch nodes are compress(children) = the same (s, w) sums).
void device_add(mrx_tree &out, int n, const double *c, mrx_tree *const *inp, double prec, int maxIter, bool absPrec) {
    require_device("device_add");
    Tree<3> &h = out.host;
    cudaStream_t st = stream();
    for (int i = 0; i < n; i++)
        if (!inp[i]->devValid) tree_upload(*inp[i]);
    out.dev.coefs.reserve((size_t)h.nReal * h.ncoef, false, st);
    out.dev.norms.reserve((size_t)h.nReal * 8, false, st);
    MRX_CUDA(cudaMemsetAsync(out.dev.coefs.p, 0, sizeof(double) * (size_t)h.nReal * h.ncoef, st));
    out.dev.nNodes = h.nReal;
    out.dev.nGen = 0;
    out.dev.topoNodes = -1;
    out.dev.partial = false;
    std::vector<std::vector<int>> pairs(n);
    std::vector<DevBuf<int>> dpairs(n);
    for (int i = 0; i < n; i++) {
        const Tree<3> &b = inp[i]->host;
        // nodes present in both trees, walked together from the roots
        std::vector<std::pair<int, int>> stack;
        for (int r = h.nRoots - 1; r >= 0; r--) stack.push_back({r, r});
        while (!stack.empty()) {
            auto pr = stack.back();
            stack.pop_back();
            pairs[i].push_back(pr.first);
            pairs[i].push_back(pr.second);
            const bool oB = h.isBranch(pr.first) && !h.isGen(h.nodes[pr.first].child0);
            const bool iB = b.isBranch(pr.second) && !b.isGen(b.nodes[pr.second].child0);
            if (oB && iB)
                for (int k = 7; k >= 0; k--) stack.push_back({h.nodes[pr.first].child0 + k, b.nodes[pr.second].child0 + k});
        }
        const int np = (int)pairs[i].size() / 2;
        dpairs[i].reserve(pairs[i].size(), false, st);
        MRX_CUDA(cudaMemcpyAsync(dpairs[i].p, pairs[i].data(), sizeof(int) * pairs[i].size(), cudaMemcpyHostToDevice, st));
        launch_axpy_nodes(out.dev.coefs.p, inp[i]->dev.coefs.p, dpairs[i].p, np, h.nRoots, h.Kd, c[i], st);
    }
    MRX_CUDA(cudaStreamSynchronize(st)); // the pair lists are host vectors read by the asynchronous copies above
    for (int s = 0; s < h.nReal; s++) h.nodes[s].flags |= FlagHasCoefs;
    out.devValid = true;
    out.hostCoefsValid = false;
    device_mw_transform(out, MRX_TOP_DOWN, /*overwrite=*/false); // + norms of every node
    // Refinement (prec > 0): the TreeBuilder loop (TreeBuilder.cpp:38-86) with the WaveletAdaptor (WaveletAdaptor.h:51-54), from
    // the end nodes of the grid just computed. The (s, w) blocks of a new child are fixed by what is already there: its scaling
    // block is the reconstruction of its parent (one transform launch over the split parents, into zeroed storage), its wavelet blocks the sum of the wavelet blocks of the inputs that hold the node (axpy launches);
    // norms of the new nodes come back to the host, which takes the split decisions like the projection does.
    if (prec > 0.0 && maxIter != 0) {
        const double *filt = device_filters(h.k);
        const int maxScale = h.mra.maxScale();
        h.allocCoefs = false; // new nodes are born in HBM
        std::vector<int> work, next, parentPairs, slotPairs;
        h.endNodeTable(work);
        DevBuf<int> dParents, dSlots, dAxpy;
        DevBuf<double> dNormsW;
        std::vector<double> nrm;
        double sNorm = 0.0, wNorm = 0.0;
        int iter = 0;
        while (!work.empty()) {
            if (iter == 0) {
                sNorm = 0.0;
                for (int s : work) sNorm += h.scalingNorm(s);
            }
            for (int s : work) wNorm += h.waveletNorm(s);
            if (sNorm < 0.0 or wNorm < 0.0) h.squareNorm = -1.0;
            else h.squareNorm = sNorm + wNorm;
            next.clear();
            parentPairs.clear();
            if (iter >= maxIter and maxIter >= 0) work.clear();
            for (int s : work) {
                if (h.isBranch(s)) continue;
                if (h.nodes[s].scale + 2 > maxScale) continue;
                if (split_check(h, s, prec, 1.0, absPrec)) {
                    const int c0 = h.createChildren(s, false);
                    parentPairs.push_back(s);
                    parentPairs.push_back(c0);
                    for (int k = 0; k < 8; k++) next.push_back(c0 + k);
                }
            }
            if (next.empty()) break;
            const int nW = (int)next.size(), nP = (int)parentPairs.size() / 2;
            out.dev.coefs.reserve((size_t)h.nReal * h.ncoef, true, st);
            out.dev.norms.reserve((size_t)h.nReal * 8, true, st);
            // the new children are the last nW slots: zero them, then children.scaling += reconstruct(parent)
            MRX_CUDA(cudaMemsetAsync(out.dev.coefs.p + (size_t)(h.nReal - nW) * h.ncoef, 0, sizeof(double) * (size_t)nW * h.ncoef, st));
            dParents.reserve(parentPairs.size(), false, st);
            MRX_CUDA(cudaMemcpyAsync(dParents.p, parentPairs.data(), sizeof(int) * parentPairs.size(), cudaMemcpyHostToDevice, st));
            launch_transform(true, false, out.dev.coefs.p, dParents.p, nP, h.K, filt, st);
            for (int i = 0; i < n; i++) {
                const Tree<3> &b = inp[i]->host;
                slotPairs.clear();
                for (int s : next) {
                    const int m = b.findNode(h.nodes[s].scale, h.nodes[s].l);
                    if (m >= 0 && m < b.nReal) {
                        slotPairs.push_back(s);
                        slotPairs.push_back(m);
                    }
                }
                if (slotPairs.empty()) continue;
                MRX_CUDA(cudaStreamSynchronize(st)); // dAxpy and slotPairs are reused per input
                dAxpy.reserve(slotPairs.size(), false, st);
                MRX_CUDA(cudaMemcpyAsync(dAxpy.p, slotPairs.data(), sizeof(int) * slotPairs.size(), cudaMemcpyHostToDevice, st));
                launch_axpy_nodes(out.dev.coefs.p, inp[i]->dev.coefs.p, dAxpy.p, (int)slotPairs.size() / 2, /*nRoots=*/0, h.Kd, c[i], st);
            }
            dSlots.reserve(nW, false, st);
            dNormsW.reserve((size_t)nW * 8, false, st);
            MRX_CUDA(cudaMemcpyAsync(dSlots.p, next.data(), sizeof(int) * nW, cudaMemcpyHostToDevice, st));
            launch_norms(out.dev.coefs.p, out.dev.norms.p, dSlots.p, nW, h.Kd, st, dNormsW.p);
            nrm.resize((size_t)nW * 8);
            MRX_CUDA(cudaMemcpyAsync(nrm.data(), dNormsW.p, sizeof(double) * nrm.size(), cudaMemcpyDeviceToHost, st));
            MRX_CUDA(cudaStreamSynchronize(st));
            for (int i = 0; i < nW; i++) {
                const int s = next[i];
                double sq = 0.0;
                for (int k = 0; k < 8; k++) {
                    const double v = nrm[(size_t)i * 8 + k];
                    h.cnorm[(size_t)s * 8 + k] = v;
                    sq += v * v;
                }
                h.sqn[s] = sq;
                h.nodes[s].flags |= FlagHasCoefs;
            }
            work.swap(next);
            iter++;
        }
        out.dev.nNodes = h.nReal;
        out.dev.topoNodes = -1;
    }
    h.calcSquareNorm();
}

// multiply(prec, out, {(c_i, inp_i)}, maxIter, absPrec) (src/treebuilders/multiply.cpp:104-136, MultiplicationCalculator.h:43-72,
// WaveletAdaptor): per END node of the output grid the inputs' nodes at the same index are reconstructed, taken to function
// values at the children's quadrature points and multiplied; the product returns through cvTransform(Backward) and
// mwTransform(Compression). Built from the existing transform kernels and one element-wise kernel:
//  * side stores X_i: every input represented on the OUTPUT grid ((s, w) per output node; wavelet blocks where the input holds
//    the node, scaling blocks by TopDown(+=) -- the formulation of device_add, so nothing is generated), extended as the grid
//    refines;
//  * per chunk of work nodes: gather X_i(n) into a scratch "parent" slot, children.scaling += reconstruct(parent) into eight
//    zeroed scratch child slots (the in-node Reconstruction), values_kernel accumulates c_i * value_i into the product scratch
//    and finally applies the Backward map, BottomUp of the scratch children gives the compressed product node, which is
//    copied into the node store.
void device_multiply(mrx_tree &out, int n, const double *c, mrx_tree *const *inp, double prec, int maxIter, bool absPrec, bool useMaxNorms,
                     const double *power) {
    require_device("device_multiply");
    if (useMaxNorms && n != 2) MRX_ABORT("Invalid tree vec size"); // MultiplicationAdaptor.h:47
    if (power && n != 1) MRX_ABORT("power: one input");             // power(prec, out, inp, p): multiply.cpp:211-234, PowerCalculator.h:43-58
    Tree<3> &h = out.host;
    cudaStream_t st = stream();
    const int K = h.K, Kd = h.Kd, ncoef = h.ncoef;
    const double *filt = device_filters(h.k);
    const int maxScale = h.mra.maxScale();
    for (int i = 0; i < n; i++)
        if (!inp[i]->devValid) tree_upload(*inp[i]);
    // MWTree::makeMaxSquareNorms (MWTree.cpp:536-543) for the MultiplicationAdaptor: per real input node the largest scaled square
    // norm 2^(3 n) |node|^2 (and scaled wavelet norm) among the node and its descendants, from the norms the host holds
    std::vector<std::vector<double>> maxS(n), maxW(n);
    if (useMaxNorms)
        for (int i = 0; i < n; i++) {
            const Tree<3> &b = inp[i]->host;
            maxS[i].assign(b.nReal, 0.0);
            maxW[i].assign(b.nReal, 0.0);
            for (int m = b.nReal - 1; m >= 0; m--) { // children have larger slots than their parent
                const double f = std::pow(2.0, 3 * b.nodes[m].scale);
                maxS[i][m] = f * b.sqn[m];
                maxW[i][m] = f * b.waveletNorm(m);
                if (b.isBranch(m) && b.nodes[m].child0 < b.nReal)
                    for (int k = 0; k < 8; k++) {
                        maxS[i][m] = std::max(maxS[i][m], maxS[i][b.nodes[m].child0 + k]);
                        maxW[i][m] = std::max(maxW[i][m], maxW[i][b.nodes[m].child0 + k]);
                    }
            }
        }
    std::vector<double> hcv(K), hsw(K);
    {
        const Quadrature &q = quadrature(K);
        for (int j = 0; j < K; j++) {
            hcv[j] = std::sqrt(1.0 / q.weights[j]); // InterpolatingBasis::calcCVMaps (InterpolatingBasis.cpp:115-124)
            hsw[j] = std::sqrt(q.weights[j]);
        }
    }
    DevBuf<double> dcv, dsw;
    dcv.reserve(K, false, st);
    dsw.reserve(K, false, st);
    MRX_CUDA(cudaMemcpyAsync(dcv.p, hcv.data(), sizeof(double) * K, cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(dsw.p, hsw.data(), sizeof(double) * K, cudaMemcpyHostToDevice, st));

    // output node store and the side stores, all zero
    out.dev.coefs.reserve((size_t)h.nReal * ncoef, false, st);
    out.dev.norms.reserve((size_t)h.nReal * 8, false, st);
    MRX_CUDA(cudaMemsetAsync(out.dev.coefs.p, 0, sizeof(double) * (size_t)h.nReal * ncoef, st));
    out.dev.nNodes = h.nReal;
    out.dev.nGen = 0;
    out.dev.topoNodes = -1;
    out.dev.partial = false;
    std::vector<DevBuf<double>> X(n);
    DevBuf<int> dA, dB;
    std::vector<int> hostPairs;
    {
        std::vector<int> levelOff, flat;
        const int nPairs = build_level_pairs(h, levelOff, flat);
        const int nLevels = (int)levelOff.size() - 1;
        if (nPairs > 0) {
            dB.reserve(flat.size(), false, st);
            MRX_CUDA(cudaMemcpyAsync(dB.p, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice, st));
        }
        for (int i = 0; i < n; i++) {
            const Tree<3> &b = inp[i]->host;
            X[i].reserve((size_t)h.nReal * ncoef, false, st);
            MRX_CUDA(cudaMemsetAsync(X[i].p, 0, sizeof(double) * (size_t)h.nReal * ncoef, st));
            hostPairs.clear();
            std::vector<std::pair<int, int>> stack;
            for (int r = h.nRoots - 1; r >= 0; r--) stack.push_back({r, r});
            while (!stack.empty()) {
                auto pr = stack.back();
                stack.pop_back();
                hostPairs.push_back(pr.first);
                hostPairs.push_back(pr.second);
                const bool oB = h.isBranch(pr.first) && !h.isGen(h.nodes[pr.first].child0);
                const bool iB = b.isBranch(pr.second) && !b.isGen(b.nodes[pr.second].child0);
                if (oB && iB)
                    for (int k = 7; k >= 0; k--) stack.push_back({h.nodes[pr.first].child0 + k, b.nodes[pr.second].child0 + k});
            }
            MRX_CUDA(cudaStreamSynchronize(st)); // dA / hostPairs are reused per input
            dA.reserve(hostPairs.size(), false, st);
            MRX_CUDA(cudaMemcpyAsync(dA.p, hostPairs.data(), sizeof(int) * hostPairs.size(), cudaMemcpyHostToDevice, st));
            launch_axpy_nodes(X[i].p, inp[i]->dev.coefs.p, dA.p, (int)hostPairs.size() / 2, h.nRoots, Kd, 1.0, st);
            for (int d = 0; d < nLevels && nPairs > 0; d++) {
                const int cnt = levelOff[d + 1] - levelOff[d];
                if (cnt > 0) launch_transform(true, false, X[i].p, dB.p + 2 * (size_t)levelOff[d], cnt, K, filt, st);
            }
        }
        MRX_CUDA(cudaStreamSynchronize(st));
    }

    // scratch: per chunk node one "parent" slot and eight "child" slots, for the current input (S) and for the product (P)
    const size_t slotBytes = sizeof(double) * (size_t)ncoef;
    const int chunkCap = (int)std::max<size_t>(64, std::min<size_t>(16384, ((size_t)768 << 20) / (9 * slotBytes)));
    DevBuf<double> S, P, dNormsW, xNorms;
    DevBuf<int> dGather, dKids, dScale, dSlots, dParents;
    std::vector<int> work, next, parentPairs, gather, kids, scales;
    std::vector<double> nrm;
    h.endNodeTable(work);
    h.allocCoefs = false; // new nodes are born in HBM
    double sNorm = 0.0, wNorm = 0.0;
    int iter = 0;
    while (!work.empty()) {
        const int nW = (int)work.size();
        for (int w0 = 0; w0 < nW; w0 += chunkCap) {
            const int nC = std::min(chunkCap, nW - w0);
            S.reserve((size_t)9 * nC * ncoef, false, st);
            P.reserve((size_t)9 * nC * ncoef, false, st);
            gather.resize((size_t)2 * nC);
            kids.resize((size_t)2 * nC);
            scales.resize(nC);
            for (int j = 0; j < nC; j++) {
                gather[2 * (size_t)j] = j;                // scratch parent slot j <- node store slot
                gather[2 * (size_t)j + 1] = work[w0 + j];
                kids[2 * (size_t)j] = j;                  // scratch children of parent j: slots nC + 8 j .. + 7
                kids[2 * (size_t)j + 1] = nC + 8 * j;
                scales[j] = h.nodes[work[w0 + j]].scale;
            }
            dGather.reserve(gather.size(), false, st);
            dKids.reserve(kids.size(), false, st);
            dScale.reserve(nC, false, st);
            MRX_CUDA(cudaMemcpyAsync(dGather.p, gather.data(), sizeof(int) * gather.size(), cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemcpyAsync(dKids.p, kids.data(), sizeof(int) * kids.size(), cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemcpyAsync(dScale.p, scales.data(), sizeof(int) * nC, cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemsetAsync(P.p, 0, slotBytes * 9 * nC, st));
            for (int i = 0; i < n; i++) {
                MRX_CUDA(cudaMemsetAsync(S.p, 0, slotBytes * 9 * nC, st));
                launch_axpy_nodes(S.p, X[i].p, dGather.p, nC, /*nRoots: scaling block of every node*/ 0x7fffffff, Kd, 1.0, st);
                launch_transform(true, false, S.p, dKids.p, nC, K, filt, st); // in-node Reconstruction into the scratch children
                if (power) launch_product_values(P.p, S.p, dScale.p, nC, K, dcv.p, *power, 3, st); // values ^ p
                else launch_product_values(P.p, S.p, dScale.p, nC, K, dcv.p, c[i], i == 0 ? 0 : 1, st);
            }
            launch_product_values(P.p, nullptr, dScale.p, nC, K, dsw.p, 1.0, 2, st); // cvTransform(Backward)
            launch_transform(false, true, P.p, dKids.p, nC, K, filt, st);            // in-node Compression: parent slot j
            // compressed product nodes into the (zeroed) node store slots: gather list reversed
            for (int j = 0; j < nC; j++) std::swap(gather[2 * (size_t)j], gather[2 * (size_t)j + 1]);
            MRX_CUDA(cudaStreamSynchronize(st)); // dGather is still read by the launches above
            MRX_CUDA(cudaMemcpyAsync(dGather.p, gather.data(), sizeof(int) * gather.size(), cudaMemcpyHostToDevice, st));
            launch_axpy_nodes(out.dev.coefs.p, P.p, dGather.p, nC, 0x7fffffff, Kd, 1.0, st);
            MRX_CUDA(cudaStreamSynchronize(st));
        }
        // norms of the work nodes -> host bookkeeping (TreeBuilder.cpp:56-66)
        dSlots.reserve(nW, false, st);
        dNormsW.reserve((size_t)nW * 8, false, st);
        MRX_CUDA(cudaMemcpyAsync(dSlots.p, work.data(), sizeof(int) * nW, cudaMemcpyHostToDevice, st));
        launch_norms(out.dev.coefs.p, out.dev.norms.p, dSlots.p, nW, Kd, st, dNormsW.p);
        nrm.resize((size_t)nW * 8);
        MRX_CUDA(cudaMemcpyAsync(nrm.data(), dNormsW.p, sizeof(double) * nrm.size(), cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < nW; i++) {
            const int s = work[i];
            double sq = 0.0;
            for (int k = 0; k < 8; k++) {
                const double v = nrm[(size_t)i * 8 + k];
                h.cnorm[(size_t)s * 8 + k] = v;
                sq += v * v;
            }
            h.sqn[s] = sq;
            h.nodes[s].flags |= FlagHasCoefs;
        }
        if (iter == 0) {
            sNorm = 0.0;
            for (int s : work) sNorm += h.scalingNorm(s);
        }
        for (int s : work) wNorm += h.waveletNorm(s);
        if (sNorm < 0.0 or wNorm < 0.0) h.squareNorm = -1.0;
        else h.squareNorm = sNorm + wNorm;
        // MultiplicationAdaptor (MultiplicationAdaptor.h:46-66): where an input is coarser than the work node, its generated node
        // is the scaling block of the side store: scaled square norm from the device, no wavelet part, a leaf
        std::vector<std::vector<double>> xScal(n);
        if (useMaxNorms && !(iter >= maxIter and maxIter >= 0)) {
            xNorms.reserve((size_t)h.nReal * 8, false, st);
            for (int i = 0; i < n; i++) {
                launch_norms(X[i].p, xNorms.p, dSlots.p, nW, Kd, st, dNormsW.p);
                xScal[i].resize((size_t)nW * 8);
                MRX_CUDA(cudaMemcpyAsync(xScal[i].data(), dNormsW.p, sizeof(double) * xScal[i].size(), cudaMemcpyDeviceToHost, st));
                MRX_CUDA(cudaStreamSynchronize(st));
            }
        }
        auto split_max_norms = [&](int w, int s) {
            double S[2], W[2];
            bool leaf[2];
            for (int i = 0; i < 2; i++) {
                const Tree<3> &b = inp[i]->host;
                const int m = b.findNode(h.nodes[s].scale, h.nodes[s].l);
                const double f = std::pow(2.0, 3 * h.nodes[s].scale);
                if (m >= 0 && m < b.nReal) {
                    const double own = f * b.sqn[m], ownW = f * b.waveletNorm(m);
                    S[i] = std::sqrt(maxS[i][m] > 0.0 ? maxS[i][m] : own); // MWNode::getMaxSquareNorm (MWNode.h:84)
                    W[i] = std::sqrt(maxW[i][m] > 0.0 ? maxW[i][m] : ownW);
                    leaf[i] = !(b.isBranch(m) && b.nodes[m].child0 < b.nReal);
                } else {
                    const double v = xScal[i][(size_t)w * 8];
                    S[i] = std::sqrt(f * (v * v));
                    W[i] = 0.0;
                    leaf[i] = true;
                }
            }
            const double multNorm = W[0] * S[1] + W[1] * S[0] + W[0] * W[1];
            return multNorm > prec and not(leaf[0] and leaf[1]);
        };
        next.clear();
        parentPairs.clear();
        if (iter >= maxIter and maxIter >= 0) work.clear();
        for (int w = 0; w < (int)work.size(); w++) {
            const int s = work[w];
            if (h.isBranch(s)) continue;
            if (h.nodes[s].scale + 2 > maxScale) continue;
            if (useMaxNorms ? split_max_norms(w, s) : split_check(h, s, prec, 1.0, absPrec)) {
                const int c0 = h.createChildren(s, false);
                parentPairs.push_back(s);
                parentPairs.push_back(c0);
                for (int k = 0; k < 8; k++) next.push_back(c0 + k);
            }
        }
        if (!next.empty()) {
            // grow the node store and the side stores; new children are the last slots: zero, then X_i(child).s += reconstruct(X_i(parent)),
            // and the wavelet blocks of the children the input holds
            const int nNew = (int)next.size(), nP = (int)parentPairs.size() / 2;
            out.dev.coefs.reserve((size_t)h.nReal * ncoef, true, st);
            out.dev.norms.reserve((size_t)h.nReal * 8, true, st);
            MRX_CUDA(cudaMemsetAsync(out.dev.coefs.p + (size_t)(h.nReal - nNew) * ncoef, 0, slotBytes * nNew, st));
            dParents.reserve(parentPairs.size(), false, st);
            MRX_CUDA(cudaMemcpyAsync(dParents.p, parentPairs.data(), sizeof(int) * parentPairs.size(), cudaMemcpyHostToDevice, st));
            for (int i = 0; i < n; i++) {
                const Tree<3> &b = inp[i]->host;
                X[i].reserve((size_t)h.nReal * ncoef, true, st);
                MRX_CUDA(cudaMemsetAsync(X[i].p + (size_t)(h.nReal - nNew) * ncoef, 0, slotBytes * nNew, st));
                launch_transform(true, false, X[i].p, dParents.p, nP, K, filt, st);
                hostPairs.clear();
                for (int s : next) {
                    const int m = b.findNode(h.nodes[s].scale, h.nodes[s].l);
                    if (m >= 0 && m < b.nReal) {
                        hostPairs.push_back(s);
                        hostPairs.push_back(m);
                    }
                }
                if (hostPairs.empty()) continue;
                MRX_CUDA(cudaStreamSynchronize(st)); // dA / hostPairs are reused per input
                dA.reserve(hostPairs.size(), false, st);
                MRX_CUDA(cudaMemcpyAsync(dA.p, hostPairs.data(), sizeof(int) * hostPairs.size(), cudaMemcpyHostToDevice, st));
                launch_axpy_nodes(X[i].p, inp[i]->dev.coefs.p, dA.p, (int)hostPairs.size() / 2, /*nRoots=*/0, Kd, 1.0, st);
            }
            MRX_CUDA(cudaStreamSynchronize(st));
            out.dev.nNodes = h.nReal;
        }
        work.swap(next);
        iter++;
    }
    for (int s = 0; s < h.nReal; s++) h.nodes[s].flags |= FlagHasCoefs;
    out.dev.nNodes = h.nReal;
    out.devValid = true;
    out.hostCoefsValid = false;
    device_mw_transform(out, MRX_BOTTOM_UP, true); // branch nodes of the grid + norms of every node (multiply.cpp:122-123)
    h.calcSquareNorm();
}

// refine_grid(out, prec, absPrec) / refine_grid(out, scales) (src/treebuilders/grid.cpp:271-302, TreeBuilder::split
// TreeBuilder.cpp:106-131): the host takes the split decisions from the norms it already holds; the new children get their
// coefficients by giveChildrenCoefs(overwrite) = reconstruction of the parent into zeroed storage. Returns the new nodes.
int device_refine_grid(mrx_tree &t, double prec, bool absPrec, int scales) {
    Tree<3> &h = t.host;
    // a tree carries coefficients once something has computed its norm (projection, apply, add, ...); grids from build_grid /
    // copy_grid / clear_grid have squareNorm = -1 and nothing on the device
    const bool hasCoefs = t.devValid || h.squareNorm >= 0.0;
    cudaStream_t st = nullptr;
    const double *filt = nullptr;
    if (hasCoefs) {
        require_device("device_refine_grid");
        st = stream();
        filt = device_filters(h.k);
        if (!t.devValid) tree_upload(t);
        h.allocCoefs = false; // the new nodes are born in HBM
    }
    const int maxScale = h.mra.maxScale();
    int nNew = 0;
    std::vector<int> work, parentPairs, slots;
    DevBuf<int> dParents, dSlots;
    DevBuf<double> dNormsW;
    std::vector<double> nrm;
    for (int pass = 0; pass < std::max(scales, 1); pass++) {
        h.endNodeTable(work);
        parentPairs.clear();
        slots.clear();
        for (int n : work) {
            if (h.isBranch(n)) continue;
            if (h.nodes[n].scale + 2 > maxScale) continue;
            if (scales > 0 || split_check(h, n, prec, 1.0, absPrec)) {
                const int c0 = h.createChildren(n, false);
                parentPairs.push_back(n);
                parentPairs.push_back(c0);
                for (int k = 0; k < 8; k++) slots.push_back(c0 + k);
            }
        }
        const int nW = (int)slots.size();
        nNew += nW;
        if (nW == 0 || !hasCoefs) continue;
        t.dev.coefs.reserve((size_t)h.nReal * h.ncoef, true, st);
        t.dev.norms.reserve((size_t)h.nReal * 8, true, st);
        MRX_CUDA(cudaMemsetAsync(t.dev.coefs.p + (size_t)(h.nReal - nW) * h.ncoef, 0, sizeof(double) * (size_t)nW * h.ncoef, st));
        dParents.reserve(parentPairs.size(), false, st);
        dSlots.reserve(nW, false, st);
        dNormsW.reserve((size_t)nW * 8, false, st);
        MRX_CUDA(cudaMemcpyAsync(dParents.p, parentPairs.data(), sizeof(int) * parentPairs.size(), cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(dSlots.p, slots.data(), sizeof(int) * nW, cudaMemcpyHostToDevice, st));
        launch_transform(true, false, t.dev.coefs.p, dParents.p, (int)parentPairs.size() / 2, h.K, filt, st);
        launch_norms(t.dev.coefs.p, t.dev.norms.p, dSlots.p, nW, h.Kd, st, dNormsW.p);
        nrm.resize((size_t)nW * 8);
        MRX_CUDA(cudaMemcpyAsync(nrm.data(), dNormsW.p, sizeof(double) * nrm.size(), cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < nW; i++) {
            const int s = slots[i];
            double sq = 0.0;
            for (int k = 0; k < 8; k++) {
                const double v = nrm[(size_t)i * 8 + k];
                h.cnorm[(size_t)s * 8 + k] = v;
                sq += v * v;
            }
            h.sqn[s] = sq;
            h.nodes[s].flags |= FlagHasCoefs;
        }
        t.dev.nNodes = h.nReal;
        t.hostCoefsValid = false;
    }
    t.dev.topoNodes = -1;
    if (!hasCoefs && nNew > 0) { // a grid without coefficients: nothing lives on the device
        t.devValid = false;
        t.dev.nNodes = 0;
        t.dev.partial = false;
    }
    return nNew;
}

// FunctionTree::add(c, inp) in place (src/trees/FunctionTree.cpp:687-706), in the formulation of device_add: wavelet blocks of the
// shared nodes and root scaling blocks += c * input; every other scaling block is regenerated by TopDown(+=) from zero (the
// tree is consistent, so that reproduces its own part and adds the input's, including where the input is coarser).
void device_add_inplace(mrx_tree &out, double c, mrx_tree &inp) {
    require_device("device_add_inplace");
    if (!out.devValid) tree_upload(out);
    if (!inp.devValid) tree_upload(inp);
    Tree<3> &h = out.host;
    const Tree<3> &b = inp.host;
    cudaStream_t st = stream();
    std::vector<int> pairs;
    std::vector<std::pair<int, int>> stack;
    for (int r = h.nRoots - 1; r >= 0; r--) stack.push_back({r, r});
    while (!stack.empty()) {
        auto pr = stack.back();
        stack.pop_back();
        pairs.push_back(pr.first);
        pairs.push_back(pr.second);
        const bool oB = h.isBranch(pr.first) && !h.isGen(h.nodes[pr.first].child0);
        const bool iB = b.isBranch(pr.second) && !b.isGen(b.nodes[pr.second].child0);
        if (oB && iB)
            for (int k = 7; k >= 0; k--) stack.push_back({h.nodes[pr.first].child0 + k, b.nodes[pr.second].child0 + k});
    }
    DevBuf<int> dpairs;
    dpairs.reserve(pairs.size(), false, st);
    MRX_CUDA(cudaMemcpyAsync(dpairs.p, pairs.data(), sizeof(int) * pairs.size(), cudaMemcpyHostToDevice, st));
    launch_axpy_nodes(out.dev.coefs.p, inp.dev.coefs.p, dpairs.p, (int)pairs.size() / 2, h.nRoots, h.Kd, c, st);
    if (h.nReal > h.nRoots)
        MRX_CUDA(cudaMemset2DAsync(out.dev.coefs.p + (size_t)h.nRoots * h.ncoef, sizeof(double) * h.ncoef, 0, sizeof(double) * h.Kd,
                                   (size_t)(h.nReal - h.nRoots), st));
    MRX_CUDA(cudaStreamSynchronize(st));
    out.hostCoefsValid = false;
    device_mw_transform(out, MRX_TOP_DOWN, /*overwrite=*/false); // + norms of every node
    h.calcSquareNorm();
}

// MWNode::mwTransform(Compression | Reconstruction) (MWNode.cpp:557-594) and MWNode::cvTransform(Forward | Backward)
// (MWNode.cpp:448-490) of a list of nodes (n < 0: every node), in place on the resident node store.
// what: 0 = mwTransform (kind 0 Compression, 1 Reconstruction), 1 = cvTransform (kind 0 Forward, 1 Backward).
// timedReps > 0 (cvTransform only): Forward then Backward, timedReps times between two CUDA events; *timedMs = ms per pass.
void device_node_transform(mrx_tree &t, int what, int kind, int n, const int *slots, int timedReps, double *timedMs) {
    require_device("device_node_transform");
    if (!t.devValid) tree_upload(t);
    Tree<3> &h = t.host;
    cudaStream_t st = stream();
    const int cnt = n < 0 ? h.nReal : n;
    if (cnt == 0) return;
    std::vector<int> items((size_t)2 * cnt);
    for (int i = 0; i < cnt; i++) {
        const int slot = n < 0 ? i : slots[i];
        if (slot < 0 || slot >= h.nReal) MRX_ABORT("node transform: slot out of range");
        items[2 * i] = slot;
        items[2 * i + 1] = h.nodes[slot].scale;
    }
    DevBuf<int> dItems;
    dItems.reserve(items.size(), false, st);
    MRX_CUDA(cudaMemcpyAsync(dItems.p, items.data(), sizeof(int) * items.size(), cudaMemcpyHostToDevice, st));
    if (what == 0) {
        const double *filt = device_filters(h.k);
        if (kind == 0) launch_compress_nodes(t.dev.coefs.p, dItems.p, cnt, h.K, filt, st, nullptr);
        else launch_reconstruct_nodes(t.dev.coefs.p, dItems.p, cnt, h.K, filt, st);
    } else {
        const Quadrature &q = quadrature(h.K);
        std::vector<double> m(2 * h.K);
        for (int j = 0; j < h.K; j++) {
            m[j] = std::sqrt(1.0 / q.weights[j]); // InterpolatingBasis::calcCVMaps (InterpolatingBasis.cpp:115-124)
            m[h.K + j] = std::sqrt(q.weights[j]);
        }
        DevBuf<double> dMap;
        dMap.reserve(m.size(), false, st);
        MRX_CUDA(cudaMemcpyAsync(dMap.p, m.data(), sizeof(double) * m.size(), cudaMemcpyHostToDevice, st));
        if (timedReps > 0) {
            cudaEvent_t e0, e1;
            MRX_CUDA(cudaEventCreate(&e0));
            MRX_CUDA(cudaEventCreate(&e1));
            MRX_CUDA(cudaEventRecord(e0, st));
            for (int r = 0; r < timedReps; r++) {
                launch_cv_transform(t.dev.coefs.p, dItems.p, cnt, h.K, dMap.p, false, st);
                launch_cv_transform(t.dev.coefs.p, dItems.p, cnt, h.K, dMap.p + h.K, true, st);
            }
            MRX_CUDA(cudaEventRecord(e1, st));
            MRX_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            MRX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (timedMs) *timedMs = ms / (2.0 * timedReps);
            MRX_CUDA(cudaEventDestroy(e0));
            MRX_CUDA(cudaEventDestroy(e1));
        } else {
            launch_cv_transform(t.dev.coefs.p, dItems.p, cnt, h.K, kind == 0 ? dMap.p : dMap.p + h.K, kind != 0, st);
        }
        MRX_CUDA(cudaStreamSynchronize(st)); // dMap / m
    }
    MRX_CUDA(cudaStreamSynchronize(st)); // items
    t.hostCoefsValid = false;
    t.dev.topoNodes = -1;
    device_calc_norms_all(t);
}

void device_rescale(mrx_tree &t, double c) {
    if (!t.devValid) tree_upload(t);
    cudaStream_t st = stream();
    launch_scale(t.dev.coefs.p, (size_t)t.host.nReal * t.host.ncoef, c, st);
    MRX_CUDA(cudaStreamSynchronize(st));
    t.hostCoefsValid = false;
    device_calc_norms_all(t);
    t.host.calcSquareNorm();
}

// pack the [term][depth][transl] operator tables once into HBM
void oper_upload(mrx_oper &o) {
    require_device("oper_upload");
    if (o.dev.tablesValid) return;
    cudaStream_t st = stream();
    Operator &op = o.op;
    int M = op.size(), DM = 0;
    for (auto &t : op.terms) DM = std::max(DM, t.nDepth);
    size_t totalNodes = 0;
    o.dev.termNodeBase.assign(M, 0);
    for (int i = 0; i < M; i++) {
        o.dev.termNodeBase[i] = totalNodes;
        totalNodes += op.terms[i].norms.size() / 4;
    }
    const size_t stride = (size_t)4 * op.K * op.K;
    // one extra node at the end: block 0 = K x K identity (used by the derivative apply for the passive dimensions)
    std::vector<double> mats((totalNodes + 1) * stride, 0.0), norms((totalNodes + 1) * 4, 0.0);
    for (int i = 0; i < op.K; i++) mats[totalNodes * stride + (size_t)i * op.K + i] = 1.0;
    o.dev.identIdx = (int)(totalNodes * 4);
    std::vector<int> nodeOff((size_t)M * DM, -1), maxT((size_t)M * DM, 0), nodeBase((size_t)M * DM, -1);
    for (int i = 0; i < M; i++) {
        const OperTerm &t = op.terms[i];
        std::memcpy(mats.data() + o.dev.termNodeBase[i] * stride, t.mats.data(), sizeof(double) * t.mats.size());
        std::memcpy(norms.data() + o.dev.termNodeBase[i] * 4, t.norms.data(), sizeof(double) * t.norms.size());
        for (int d = 0; d < t.nDepth; d++) {
            nodeOff[(size_t)i * DM + d] = (int)(o.dev.termNodeBase[i] + t.offset[d]);
            maxT[(size_t)i * DM + d] = t.maxTransl[d];
            nodeBase[(size_t)i * DM + d] = nodeOff[(size_t)i * DM + d] + t.maxTransl[d];
        }
    }
    o.dev.mats.reserve(mats.size(), false, st);
    o.dev.norms.reserve(norms.size(), false, st);
    o.dev.nodeOff.reserve(nodeOff.size(), false, st);
    o.dev.maxTransl.reserve(maxT.size(), false, st);
    o.dev.nodeBase.reserve(nodeBase.size(), false, st);
    MRX_CUDA(cudaMemcpyAsync(o.dev.nodeBase.p, nodeBase.data(), sizeof(int) * nodeBase.size(), cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(o.dev.mats.p, mats.data(), sizeof(double) * mats.size(), cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(o.dev.norms.p, norms.data(), sizeof(double) * norms.size(), cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(o.dev.nodeOff.p, nodeOff.data(), sizeof(int) * nodeOff.size(), cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(o.dev.maxTransl.p, maxT.data(), sizeof(int) * maxT.size(), cudaMemcpyHostToDevice, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    o.dev.M = M;
    o.dev.DM = DM;
    o.dev.tablesValid = true;
}

} // namespace mrx
