// Flat adaptive multiwavelet tree (D = 1,2,3) with the node-local transforms the HOST needs for
// input generation and operator construction. Whole-tree transforms of 3-D function trees are GPU work.
#include "mrx_host.hpp"

#include <algorithm>
#include <cstring>

namespace mrx {

// math_utils::apply_filter (src/utils/math_utils.cpp:175-194):
// out (kp1_dm1 x kp1, col-major) (=|+=) in(kp1 x kp1_dm1, col-major)^T * F
static inline void apply_filter(double *out, const double *in, const double *F, int K, int Kdm1, bool accumulate) {
    for (int j = 0; j < K; j++) {
        for (int m = 0; m < Kdm1; m++) {
            double s = 0.0;
            const double *col = in + (size_t)K * m;
            for (int i = 0; i < K; i++) s += col[i] * F[i * K + j];
            if (accumulate) out[m + (size_t)Kdm1 * j] += s;
            else out[m + (size_t)Kdm1 * j] = s;
        }
    }
}

static void *default_chunk_alloc(size_t bytes) { return ::operator new[](bytes); }
static void default_chunk_free(void *p) { ::operator delete[](p); }
void *(*chunk_alloc)(size_t) = default_chunk_alloc;
void (*chunk_free)(void *) = default_chunk_free;

template <int D> Tree<D>::~Tree() {
    for (double *c : chunks_) free_(c);
}

template <int D>
Tree<D>::Tree(const MRA<D> &m)
        : mra(m) {
    k = m.order;
    K = k + 1;
    Kd = ipow(K, D);
    ncoef = tdim * Kd;
    nRoots = m.nRoots();
    fs_ = &filter_set(k);
    alloc_ = chunk_alloc;
    free_ = chunk_free;
    if (m.maxDepth > MaxDepth) MRX_ABORT("Beyond MaxDepth");
    if (m.maxScale() > MaxScale) MRX_ABORT("Beyond MaxScale");
    // root nodes in box order, x fastest (BoundingBox.cpp:322-335)
    int s = allocNodes(nRoots);
    for (int r = 0; r < nRoots; r++) {
        NodeRec<D> &nd = nodes[s + r];
        nd.scale = m.rootScale;
        int rem = r;
        for (int d = 0; d < D; d++) {
            nd.l[d] = m.corner[d] + rem % m.nboxes[d];
            rem /= m.nboxes[d];
        }
        nd.parent = -1;
        nd.child0 = -1;
        nd.flags = FlagEnd;
        nd.hpath = 0;
    }
    nReal = nRoots;
}

template <int D> bool Tree<D>::coefsPinned() const { return alloc_ != default_chunk_alloc; }

template <int D> int Tree<D>::allocNodes(int count) {
    int s = (int)nodes.size();
    nodes.resize(s + count);
    cnorm.resize((size_t)(s + count) * tdim, -1.0);
    sqn.resize(s + count, -1.0);
    if (allocCoefs) ensureCoefStorage();
    return s;
}

template <int D> void Tree<D>::ensureCoefStorage() {
    size_t needChunks = (nodes.size() + chunkMask_) >> chunkShift_;
    while (chunks_.size() < needChunks) {
        size_t n = (size_t)(chunkMask_ + 1) * ncoef;
        chunks_.push_back(static_cast<double *>(alloc_(n * sizeof(double))));
        if (allocCoefs) std::memset(chunks_.back(), 0, n * sizeof(double));
    }
}

template <int D> void Tree<D>::ensureCoefStorageFor(size_t nSlots) {
    size_t needChunks = (nSlots + chunkMask_) >> chunkShift_;
    while (chunks_.size() < needChunks) chunks_.push_back(static_cast<double *>(alloc_((size_t)(chunkMask_ + 1) * ncoef * sizeof(double))));
}

template <int D> void Tree<D>::rebaseChunks(void *(*alloc)(size_t), void (*free)(void *)) {
    const size_t bytes = (size_t)(chunkMask_ + 1) * ncoef * sizeof(double);
    for (double *&c : chunks_) {
        double *n = static_cast<double *>(alloc(bytes));
        std::memcpy(n, c, bytes);
        free_(c);
        c = n;
    }
    alloc_ = alloc;
    free_ = free;
}

template <int D>
int Tree<D>::getNodeTopo(int scale, const std::array<int, D> &l, std::vector<int> *newParents, bool withCoefStorage) {
    int n = rootIndex(scale, l);
    if (n < 0) return -1;
    while (nodes[n].scale < scale) {
        if (nodes[n].child0 < 0) {
            bool saved = allocCoefs;
            allocCoefs = withCoefStorage;
            createChildren(n, true);
            allocCoefs = saved;
            if (newParents) newParents->push_back(n);
        }
        int shift = scale - nodes[n].scale - 1;
        int c = 0;
        for (int d = 0; d < D; d++) c |= ((l[d] >> shift) & 1) << d;
        n = nodes[n].child0 + c;
    }
    return n;
}

template <int D> void Tree<D>::clearToRoots() {
    nodes.resize(nRoots);
    cnorm.assign((size_t)nRoots * tdim, -1.0);
    sqn.assign(nRoots, -1.0);
    for (int r = 0; r < nRoots; r++) {
        nodes[r].child0 = -1;
        nodes[r].flags = FlagEnd;
        if (allocCoefs && !chunks_.empty()) std::memset(coef(r), 0, sizeof(double) * ncoef);
    }
    nReal = nRoots;
    squareNorm = -1.0;
}

template <int D> void Tree<D>::copyGridFrom(const Tree<D> &other) {
    // same slot order as `other` (real nodes only): replay the splits in creation order
    clearToRoots();
    std::vector<std::pair<int, int>> splits;
    for (int n = 0; n < other.nReal; n++)
        if (other.isBranch(n) && !other.isGen(other.nodes[n].child0)) splits.push_back({other.nodes[n].child0, n});
    std::sort(splits.begin(), splits.end());
    for (auto &s : splits) {
        int c0 = createChildren(s.second, false);
        if (c0 != s.first) MRX_ABORT("copyGridFrom: slot order mismatch");
    }
}

// build_grid(out, inp) (src/treebuilders/grid.cpp:144-153): TreeBuilder loop with a CopyAdaptor -- the end nodes of this tree
// split wherever `other` has real children at the same index, until every node of `other` is covered; nodes this tree
// already has stay
template <int D> void Tree<D>::extendGridFrom(const Tree<D> &other) {
    std::vector<int> work, next;
    endNodeTable(work);
    while (!work.empty()) {
        next.clear();
        for (int n : work) {
            if (isBranch(n)) continue;
            const int m = other.findNode(nodes[n].scale, nodes[n].l);
            if (m < 0 || !other.isBranch(m) || other.isGen(other.nodes[m].child0)) continue;
            if (nodes[n].scale + 1 > mra.maxScale()) continue;
            const int c0 = createChildren(n, false);
            for (int c = 0; c < tdim; c++) next.push_back(c0 + c);
        }
        work.swap(next);
    }
}

template <int D> int Tree<D>::nDepths() const {
    int md = 0;
    for (const auto &nd : nodes) md = std::max(md, nd.scale - mra.rootScale);
    return md + 1;
}

template <int D> int Tree<D>::rootIndex(int scale, const std::array<int, D> &l) const {
    int rel = scale - mra.rootScale;
    if (rel < 0) return -1;
    int bIdx = 0, ncells = 1;
    for (int d = 0; d < D; d++) {
        int req = (l[d] >> rel) - mra.corner[d];
        if (req < 0 or req >= mra.nboxes[d]) return -1;
        bIdx += ncells * req;
        ncells *= mra.nboxes[d];
    }
    return bIdx;
}

template <int D> int Tree<D>::createChildren(int n, bool gen) {
    if (isBranch(n)) MRX_ABORT("Node already has children");
    int s = allocNodes(tdim);
    NodeRec<D> parent = nodes[n];
    for (int c = 0; c < tdim; c++) {
        NodeRec<D> &nd = nodes[s + c];
        nd.scale = parent.scale + 1;
        for (int d = 0; d < D; d++) nd.l[d] = 2 * parent.l[d] + ((c >> d) & 1);
        nd.parent = n;
        nd.child0 = -1;
        nd.flags = gen ? FlagGen : FlagEnd;
        int h = hilbert_h_index(D, parent.hpath, c);
        nd.hpath = (uint8_t)hilbert_child_path(D, parent.hpath, h);
        for (int t = 0; t < tdim; t++) cnorm[(size_t)(s + c) * tdim + t] = -1.0;
        sqn[s + c] = -1.0;
        if (allocCoefs) std::memset(coef(s + c), 0, sizeof(double) * ncoef);
    }
    nodes[n].child0 = s;
    nodes[n].flags |= FlagBranch;
    if (!gen) {
        nodes[n].flags &= ~FlagEnd;
        if (s != nReal) MRX_ABORT("real nodes must be created before generated ones");
        nReal = s + tdim;
    }
    return s;
}

template <int D> int Tree<D>::findNode(int scale, const std::array<int, D> &l) const {
    int r = rootIndex(scale, l);
    if (r < 0) return -1;
    int n = r;
    while (nodes[n].scale < scale) {
        if (nodes[n].child0 < 0) return -1;
        int shift = scale - nodes[n].scale - 1;
        int c = 0;
        for (int d = 0; d < D; d++) c |= ((l[d] >> shift) & 1) << d;
        n = nodes[n].child0 + c;
    }
    return n;
}

template <int D> int Tree<D>::getNode(int scale, const std::array<int, D> &l) {
    int r = rootIndex(scale, l);
    if (r < 0) MRX_ABORT("getNode: index outside world");
    int n = r;
    while (nodes[n].scale < scale) {
        if (nodes[n].child0 < 0) {
            // FunctionNode::genChildren + giveChildrenCoefs (MWNode.cpp:410-418); operator-tree
            // generated nodes are regular nodes (OperatorNode.cpp:147-150)
            createChildren(n, !operNorms);
            giveChildrenCoefs(n, true);
        }
        int shift = scale - nodes[n].scale - 1;
        int c = 0;
        for (int d = 0; d < D; d++) c |= ((l[d] >> shift) & 1) << d;
        n = nodes[n].child0 + c;
    }
    return n;
}

template <int D> void Tree<D>::deleteGenerated() {
    if ((int)nodes.size() == nReal) return;
    nodes.resize(nReal);
    cnorm.resize((size_t)nReal * tdim);
    sqn.resize(nReal);
    for (auto &nd : nodes)
        if (nd.child0 >= nReal) {
            nd.child0 = -1;
            nd.flags &= ~FlagBranch;
        }
}

template <int D> void Tree<D>::zeroCoefs(int n) {
    std::memset(coef(n), 0, sizeof(double) * ncoef);
    nodes[n].flags |= FlagHasCoefs;
    for (int t = 0; t < tdim; t++) cnorm[(size_t)n * tdim + t] = 0.0;
    sqn[n] = 0.0;
}

template <int D> void Tree<D>::calcNorms(int n) {
    const double *c = coef(n);
    double sq = 0.0;
    for (int i = 0; i < tdim; i++) {
        double norm_i = 0.0;
        if (isGen(n) and i != 0) {
            norm_i = 0.0;
        } else if (operNorms) {
            // OperatorNode::calcComponentNorm (OperatorNode.cpp:56-80); D == 2
            int dep = depth(n);
            double thrs = std::max(MachinePrec, normPrec / (8.0 * (1 << dep)));
            const double *v = c + (size_t)i * Kd;
            double vs = 0.0;
            for (int j = 0; j < Kd; j++) vs += v[j] * v[j];
            double vecNorm = std::sqrt(vs);
            if (vecNorm > thrs) {
                double infNorm = 0.0, oneNorm = 0.0;
                for (int r = 0; r < K; r++) {
                    double s = 0.0;
                    for (int cc = 0; cc < K; cc++) s += std::abs(v[r + K * cc]);
                    infNorm = std::max(infNorm, s);
                }
                for (int cc = 0; cc < K; cc++) {
                    double s = 0.0;
                    for (int r = 0; r < K; r++) s += std::abs(v[r + K * cc]);
                    oneNorm = std::max(oneNorm, s);
                }
                if (std::sqrt(infNorm * oneNorm) > thrs) {
                    double twoNorm = vecNorm; // matrix_norm_2 == lpNorm<2> == Frobenius (math_utils.cpp:60-62)
                    if (twoNorm > thrs) norm_i = twoNorm;
                }
            }
        } else {
            const double *v = c + (size_t)i * Kd;
            double s = 0.0;
            for (int j = 0; j < Kd; j++) s += v[j] * v[j];
            norm_i = std::sqrt(s);
        }
        cnorm[(size_t)n * tdim + i] = norm_i;
        sq += norm_i * norm_i;
    }
    sqn[n] = sq;
}

template <int D> double Tree<D>::scalingNorm(int n) const {
    double s = cnorm[(size_t)n * tdim];
    return (s >= 0.0) ? s * s : -1.0;
}

template <int D> double Tree<D>::waveletNorm(int n) const {
    double w = 0.0;
    for (int i = 1; i < tdim; i++) {
        double norm_i = cnorm[(size_t)n * tdim + i];
        if (norm_i >= 0.0) w = std::fma(norm_i, norm_i, w); // explicit: the device restates this sum (apply_split.cu)
        else w = -1.0;
    }
    return w;
}

template <int D> void Tree<D>::mwTransformNode(int n, int op) {
    const FilterSet &fs = *fs_;
    int Kdm1 = Kd / K;
    std::vector<double> tmp(ncoef);
    double *in_vec = coef(n), *out_vec = tmp.data();
    for (int i = 0; i < D; i++) {
        int mask = 1 << i;
        for (int gt = 0; gt < tdim; gt++) {
            double *out = out_vec + (size_t)gt * Kd;
            bool acc = false;
            for (int ft = 0; ft < tdim; ft++) {
                if ((gt | mask) == (ft | mask)) {
                    const double *in = in_vec + (size_t)ft * Kd;
                    int fIdx = 2 * ((gt >> i) & 1) + ((ft >> i) & 1);
                    apply_filter(out, in, fs.sub[op][fIdx].data(), K, Kdm1, acc);
                    acc = true;
                }
            }
        }
        std::swap(in_vec, out_vec);
    }
    if (D % 2 == 1) std::memcpy(coef(n), in_vec, sizeof(double) * ncoef);
}

template <int D> void Tree<D>::cvTransformBackward(int n) {
    // interpolating basis: vcMap = diag(sqrt(w)) applied along each dimension, then the
    // 2^{-D(n+1)/2} factor (MWNode.cpp:448-490, InterpolatingBasis.cpp:115-124); scaling factor 1.
    const Quadrature &q = quadrature(K);
    std::vector<double> sw(K);
    for (int j = 0; j < K; j++) sw[j] = std::sqrt(q.weights[j]);
    int np1 = nodes[n].scale + 1;
    double two_fac = std::sqrt(1.0 / std::pow(2.0, D * np1));
    double *c = coef(n);
    for (int t = 0; t < tdim; t++) {
        double *b = c + (size_t)t * Kd;
        for (int idx = 0; idx < Kd; idx++) {
            double v = b[idx];
            int rem = idx;
            for (int d = 0; d < D; d++) {
                v = v * sw[rem % K];
                rem /= K;
            }
            b[idx] = two_fac * v;
        }
    }
}

// tree_utils::mw_transform (tree_utils.cpp:113-216): parent blocks -> children scaling blocks
template <int D>
void Tree<D>::mwTransformCoefs(const double *in0, double *out_children, bool readOnlyScaling, int stride,
                               bool overwrite) const {
    const FilterSet &fs = *fs_;
    int Kdm1 = Kd / K;
    std::vector<double> bufA(ncoef), bufB(ncoef);
    const double *in_vec = in0;
    double *out_vec = bufA.data();
    for (int i = 0; i < D; i++) {
        int mask = 1 << i;
        int ftlim = readOnlyScaling ? (1 << i) : tdim;
        for (int gt = 0; gt < tdim; gt++) {
            double *out = out_vec + (size_t)gt * Kd;
            bool acc = false;
            for (int ft = 0; ft < ftlim; ft++) {
                if ((gt | mask) == (ft | mask)) {
                    const double *in = in_vec + (size_t)ft * Kd;
                    int fIdx = 2 * ((gt >> i) & 1) + ((ft >> i) & 1);
                    apply_filter(out, in, fs.sub[Reconstruction][fIdx].data(), K, Kdm1, acc);
                    acc = true;
                }
            }
            if (!acc) std::memset(out, 0, sizeof(double) * Kd);
        }
        in_vec = out_vec;
        out_vec = (out_vec == bufA.data()) ? bufB.data() : bufA.data();
    }
    for (int j = 0; j < tdim; j++)
        for (int i = 0; i < Kd; i++) {
            if (overwrite) out_children[i + (size_t)j * stride] = in_vec[i + (size_t)j * Kd];
            else out_children[i + (size_t)j * stride] += in_vec[i + (size_t)j * Kd];
        }
}

template <int D> void Tree<D>::giveChildrenCoefs(int n, bool overwrite) {
    if (!isBranch(n)) MRX_ABORT("giveChildrenCoefs on leaf");
    int c0 = nodes[n].child0;
    std::vector<double> out((size_t)tdim * Kd, 0.0);
    if (!overwrite)
        for (int c = 0; c < tdim; c++) std::memcpy(out.data() + (size_t)c * Kd, coef(c0 + c), sizeof(double) * Kd);
    mwTransformCoefs(coef(n), out.data(), isGen(n), Kd, overwrite);
    for (int c = 0; c < tdim; c++) {
        if (overwrite) std::memset(coef(c0 + c), 0, sizeof(double) * ncoef);
        std::memcpy(coef(c0 + c), out.data() + (size_t)c * Kd, sizeof(double) * Kd);
        nodes[c0 + c].flags |= FlagHasCoefs;
        calcNorms(c0 + c);
    }
}

template <int D> void Tree<D>::reCompress(int n) {
    if (isGen(n)) MRX_ABORT("reCompress on generated node");
    if (!isBranch(n)) return;
    int c0 = nodes[n].child0;
    for (int c = 0; c < tdim; c++) std::memcpy(coef(n) + (size_t)c * Kd, coef(c0 + c), sizeof(double) * Kd);
    mwTransformNode(n, Compression);
    nodes[n].flags |= FlagHasCoefs;
    calcNorms(n);
}

template <int D> void Tree<D>::mwTransformUpSerial() {
    std::vector<std::vector<int>> table;
    nodeTableByDepth(table);
    for (int n = (int)table.size() - 2; n >= 0; n--)
        for (int node : table[n])
            if (isBranch(node)) reCompress(node);
}

template <int D> void Tree<D>::calcSquareNorm() {
    std::vector<int> ends;
    endNodeTable(ends);
    double s = 0.0;
    for (int n : ends) s += sqn[n];
    squareNorm = s;
}

template <int D> void Tree<D>::nodeTable(std::vector<int> &table) const {
    table.clear();
    std::vector<std::pair<int, int>> stack; // (node, next hilbert child)
    for (int r = 0; r < nRoots; r++) {
        stack.push_back({r, -1});
        while (!stack.empty()) {
            auto &top = stack.back();
            int n = top.first;
            if (top.second < 0) {
                table.push_back(n);
                top.second = 0;
            }
            bool descend = isBranch(n) && !isEnd(n) && nodes[n].child0 >= 0 && !isGen(nodes[n].child0);
            if (descend && top.second < tdim) {
                int h = top.second++;
                int c = hilbert_z_index(D, nodes[n].hpath, h);
                stack.push_back({nodes[n].child0 + c, -1});
            } else {
                stack.pop_back();
            }
        }
    }
}

template <int D> void Tree<D>::endNodeTable(std::vector<int> &table) const {
    std::vector<int> all;
    nodeTable(all);
    table.clear();
    for (int n : all)
        if (isEnd(n)) table.push_back(n);
}

template <int D> void Tree<D>::nodeTableByDepth(std::vector<std::vector<int>> &table) const {
    std::vector<int> all;
    nodeTable(all);
    table.clear();
    for (int n : all) {
        int d = depth(n);
        if (d + 1 > (int)table.size()) table.resize(d + 1);
        table[d].push_back(n);
    }
}

template <int D> void Tree<D>::lowerBounds(int n, double *lb) const {
    double len = std::pow(2.0, -nodes[n].scale);
    for (int d = 0; d < D; d++) lb[d] = len * nodes[n].l[d];
}
template <int D> void Tree<D>::upperBounds(int n, double *ub) const {
    double len = std::pow(2.0, -nodes[n].scale);
    for (int d = 0; d < D; d++) ub[d] = len * (nodes[n].l[d] + 1);
}

template <int D> bool split_check(const Tree<D> &tree, int n, double prec, double splitFac, bool absPrec) {
    bool split = false;
    if (prec > 0.0) {
        double t_norm = 1.0;
        double sq_norm = tree.squareNorm;
        if (sq_norm > 0.0 and not absPrec) t_norm = std::sqrt(sq_norm);
        double scale_fac = 1.0;
        if (splitFac > MachineZero) {
            double expo = 0.5 * splitFac * (tree.nodes[n].scale + 1);
            scale_fac = std::pow(2.0, -expo);
        }
        double w_thrs = std::max(2.0 * MachinePrec, prec * t_norm * scale_fac);
        double w_norm = std::sqrt(tree.waveletNorm(n));
        if (w_norm > w_thrs) split = true;
    }
    return split;
}

template <int D>
void build_tree(Tree<D> &tree, const std::function<void(Tree<D> &, int)> &calcNode,
                const std::function<bool(const Tree<D> &, int)> &splitNode, int maxIter, bool allNodes, bool parallel) {
    std::vector<int> workVec;
    if (allNodes) tree.nodeTable(workVec);
    else tree.endNodeTable(workVec);
    double sNorm = 0.0, wNorm = 0.0;
    int iter = 0;
    int maxScale = tree.mra.maxScale();
    while (!workVec.empty()) {
        int nNodes = (int)workVec.size();
        if (parallel) {
#pragma omp parallel for schedule(guided)
            for (int i = 0; i < nNodes; i++) calcNode(tree, workVec[i]);
        } else {
            for (int i = 0; i < nNodes; i++) calcNode(tree, workVec[i]);
        }
        if (iter == 0) {
            sNorm = 0.0;
            for (int n : workVec) sNorm += tree.scalingNorm(n);
        }
        for (int n : workVec) wNorm += tree.waveletNorm(n);
        if (sNorm < 0.0 or wNorm < 0.0) tree.squareNorm = -1.0;
        else tree.squareNorm = sNorm + wNorm;

        std::vector<int> newVec;
        if (iter >= maxIter and maxIter >= 0) workVec.clear();
        for (int n : workVec) {
            if (tree.isBranch(n)) continue;
            if (tree.nodes[n].scale + 2 > maxScale) continue;
            if (splitNode(tree, n)) {
                int c0 = tree.createChildren(n, false);
                for (int c = 0; c < Tree<D>::tdim; c++) newVec.push_back(c0 + c);
            }
        }
        workVec.swap(newVec);
        iter++;
    }
}

// ------------------------------------------------------------------ projection
template <int D> static bool analytic_split(const Tree<D> &t, int n, const GaussFunc<D> &f) {
    // AnalyticAdaptor::splitNode (AnalyticAdaptor.h:42-50)
    if (f.isVisibleAtScale(t.nodes[n].scale, t.K)) return false;
    double lb[D], ub[D];
    t.lowerBounds(n, lb);
    t.upperBounds(n, ub);
    if (f.isZeroOnInterval(lb, ub)) return false;
    return true;
}

template <int D> static void default_calc(Tree<D> &t, int n) {
    // DefaultCalculator::calcNode: clearHasCoefs + clearNorms
    t.nodes[n].flags &= ~FlagHasCoefs;
    for (int i = 0; i < Tree<D>::tdim; i++) t.cnorm[(size_t)n * Tree<D>::tdim + i] = -1.0;
    t.sqn[n] = -1.0;
}

/// build_grid for one Gaussian on a tree without coefficients, without touching the nodes the Gaussian cannot split:
/// the first work vector of TreeBuilder::build is the end-node table (DFS, Hilbert order); AnalyticAdaptor::splitNode is
/// false for a node outside the 5-sigma box or at/after the visible scale, and then false for all its descendants as well,
/// so a DFS pruned by those two tests visits the splitting end nodes in the same relative order. Same node set, same
/// creation order as the generic loop, O(nodes near the Gaussian) instead of O(all end nodes) per term.
template <int D> static void build_grid_pruned(Tree<D> &t, const GaussFunc<D> &f) {
    const int maxScale = t.mra.maxScale();
    std::vector<int> workVec;
    std::vector<std::pair<int, int>> stack;
    auto reachable = [&](int n) {
        if (f.isVisibleAtScale(t.nodes[n].scale, t.K)) return false;
        double lb[D], ub[D];
        t.lowerBounds(n, lb);
        t.upperBounds(n, ub);
        return !f.isZeroOnInterval(lb, ub);
    };
    for (int r = 0; r < t.nRoots; r++) {
        if (!reachable(r)) continue;
        stack.push_back({r, 0});
        while (!stack.empty()) {
            auto &top = stack.back();
            const int n = top.first;
            if (!t.isBranch(n)) {
                workVec.push_back(n); // end node that AnalyticAdaptor would split
                stack.pop_back();
                continue;
            }
            if (top.second < Tree<D>::tdim) {
                const int h = top.second++;
                const int c = t.nodes[n].child0 + hilbert_z_index(D, t.nodes[n].hpath, h);
                if (reachable(c)) stack.push_back({c, 0});
            } else {
                stack.pop_back();
            }
        }
    }
    while (!workVec.empty()) {
        std::vector<int> newVec;
        for (int n : workVec) {
            if (t.isBranch(n)) continue;
            if (t.nodes[n].scale + 2 > maxScale) continue;
            if (analytic_split(t, n, f)) {
                const int c0 = t.createChildren(n, false);
                for (int c = 0; c < Tree<D>::tdim; c++) newVec.push_back(c0 + c);
            }
        }
        workVec.swap(newVec);
    }
    t.squareNorm = -1.0;
}

template <int D> void build_grid(Tree<D> &out, const GaussFunc<D> &f, int maxIter) {
    bool fresh = maxIter < 0 && !getenv("MRX_GRID_GENERIC"); // debug switch: force the generic TreeBuilder loop
    for (int r = 0; r < out.nRoots && fresh; r++)
        if (out.nodes[r].flags & FlagHasCoefs) fresh = false;
    if (fresh) {
        build_grid_pruned(out, f);
        return;
    }
    build_tree<D>(
        out, [](Tree<D> &t, int n) { default_calc(t, n); },
        [&f](const Tree<D> &t, int n) { return analytic_split(t, n, f); }, maxIter, false, false);
}

template <int D> void build_grid(Tree<D> &out, const GaussExp<D> &fs, int maxIter) {
    // grid.cpp:106-123 (non-periodic branch): one pass per Gaussian
    for (const auto &f : fs) build_grid(out, f, maxIter);
}

template <int D>
void project(double prec, Tree<D> &out, const std::function<double(const double *)> &f, int maxIter, bool absPrec,
             bool finalize) {
    const Quadrature &q = quadrature(out.K);
    auto calc = [&](Tree<D> &t, int n) {
        // ProjectionCalculator::calcNode (ProjectionCalculator.cpp:34-51) with
        // MWNode::getExpandedChildPts (MWNode.cpp:903-925)
        const int K = t.K, Kd = t.Kd;
        double sFac = std::pow(2.0, -(t.nodes[n].scale + 1));
        double *c = t.coef(n);
        double r[D];
        for (int tt = 0; tt < Tree<D>::tdim; tt++) {
            for (int idx = 0; idx < Kd; idx++) {
                int rem = idx;
                for (int d = 0; d < D; d++) {
                    int j = rem % K;
                    rem /= K;
                    int b = (tt >> d) & 1;
                    r[d] = sFac * (q.roots[j] + 2.0 * static_cast<double>(t.nodes[n].l[d]) + (b ? 1.0 : 0.0));
                }
                c[(size_t)tt * Kd + idx] = f(r);
            }
        }
        t.cvTransformBackward(n);
        t.mwTransformNode(n, Compression);
        t.nodes[n].flags |= FlagHasCoefs;
        t.calcNorms(n);
    };
    auto split = [&](const Tree<D> &t, int n) { return split_check(t, n, prec, 1.0, absPrec); };
    build_tree<D>(out, calc, split, maxIter, false, true);
    if (finalize) {
        out.mwTransformUpSerial();
        out.calcSquareNorm();
    }
}

/// project() of a Gaussian expansion with exact node-level screening: a term whose exponent exceeds 746 everywhere
/// in the node's box evaluates to exactly 0 at every quadrature point (GaussFunc::evalf), so leaving it out of the sum
/// changes nothing -- except the time: far, narrow Gaussians dominate the unscreened cost.
template <int D>
void project_gaussians(double prec, Tree<D> &out, const GaussExp<D> &gexp, int maxIter, bool absPrec, bool finalize) {
    const Quadrature &q = quadrature(out.K);
    auto calc = [&](Tree<D> &t, int n) {
        const int K = t.K, Kd = t.Kd;
        const double len = std::pow(2.0, -t.nodes[n].scale);
        std::vector<int> active;
        active.reserve(gexp.size());
        for (int g = 0; g < (int)gexp.size(); g++) {
            double minq2 = 0.0;
            for (int d = 0; d < D; d++) {
                const double lb = len * t.nodes[n].l[d], ub = len * (t.nodes[n].l[d] + 1);
                const double p = gexp[g].pos[d];
                const double dist = (p < lb) ? lb - p : (p > ub ? p - ub : 0.0);
                minq2 += gexp[g].alpha * dist * dist;
            }
            if (!(minq2 > 747.0)) active.push_back(g); // 747 > 746: conservative against rounding of the box test
        }
        double sFac = std::pow(2.0, -(t.nodes[n].scale + 1));
        double *c = t.coef(n);
        double r[D];
        for (int tt = 0; tt < Tree<D>::tdim; tt++) {
            for (int idx = 0; idx < Kd; idx++) {
                int rem = idx;
                for (int d = 0; d < D; d++) {
                    int j = rem % K;
                    rem /= K;
                    int b = (tt >> d) & 1;
                    r[d] = sFac * (q.roots[j] + 2.0 * static_cast<double>(t.nodes[n].l[d]) + (b ? 1.0 : 0.0));
                }
                double s = 0.0;
                for (int g : active) s += gexp[g].evalf(r);
                c[(size_t)tt * Kd + idx] = s;
            }
        }
        t.cvTransformBackward(n);
        t.mwTransformNode(n, Compression);
        t.nodes[n].flags |= FlagHasCoefs;
        t.calcNorms(n);
    };
    auto split = [&](const Tree<D> &t, int n) { return split_check(t, n, prec, 1.0, absPrec); };
    build_tree<D>(out, calc, split, maxIter, false, true);
    if (finalize) {
        out.mwTransformUpSerial();
        out.calcSquareNorm();
    }
}
template void project_gaussians<3>(double, Tree<3> &, const GaussExp<3> &, int, bool, bool);

template class Tree<1>;
template class Tree<2>;
template class Tree<3>;
template bool split_check<1>(const Tree<1> &, int, double, double, bool);
template bool split_check<2>(const Tree<2> &, int, double, double, bool);
template bool split_check<3>(const Tree<3> &, int, double, double, bool);
template void build_tree<1>(Tree<1> &, const std::function<void(Tree<1> &, int)> &,
                            const std::function<bool(const Tree<1> &, int)> &, int, bool, bool);
template void build_tree<2>(Tree<2> &, const std::function<void(Tree<2> &, int)> &,
                            const std::function<bool(const Tree<2> &, int)> &, int, bool, bool);
template void build_tree<3>(Tree<3> &, const std::function<void(Tree<3> &, int)> &,
                            const std::function<bool(const Tree<3> &, int)> &, int, bool, bool);
template void build_grid<1>(Tree<1> &, const GaussFunc<1> &, int);
template void build_grid<3>(Tree<3> &, const GaussFunc<3> &, int);
template void build_grid<1>(Tree<1> &, const GaussExp<1> &, int);
template void build_grid<3>(Tree<3> &, const GaussExp<3> &, int);
template void project<1>(double, Tree<1> &, const std::function<double(const double *)> &, int, bool, bool);
template void project<3>(double, Tree<3> &, const std::function<double(const double *)> &, int, bool, bool);

} // namespace mrx
