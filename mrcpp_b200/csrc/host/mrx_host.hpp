// Host-side data model for the B200 multiwavelet operator-application path.
//
// This is NOT the hot path: it is the host code that sits either side of it:
// numerical tables, the MRA, flat adaptive trees in D = 1,2,3, projection of Gaussians (input
// generator) and construction of separated convolution / ABGV derivative operators (one-off cost,
// packed once for HBM). The hot path itself (apply, tree transforms, derivative apply) runs on the
// GPU through the C-ABI in include/mrcpp_b200.h; nothing here applies an operator to a function tree.
//
// Conventions restated from the reference (file:line relative to /root/reference):
//  * node = 2^D blocks of (k+1)^D coefficients, x index fastest      src/utils/math_utils.cpp:223-235
//  * child index bit d = upper half along dimension d                 src/trees/NodeIndex.h:49-53
//  * filters H0,G0 from tables; H1,G1 by interpolating symmetry       src/core/MWFilter.cpp:229-235
#pragma once
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <string>
#include <vector>

namespace mrx {

// api/constants.h:30-36, :55-56
constexpr double MachinePrec = 1.0e-15;
constexpr double MachineZero = 1.0e-14;
constexpr int MaxOrder = 41;
constexpr int MaxDepth = 30;
constexpr int MaxScale = 31;
constexpr int MaxSepRank = 1000;
constexpr double pi = 3.1415926535897932384626433832795;
constexpr double root_pi = 1.7724538509055160273;

enum { Compression = 0, Reconstruction = 1 };
enum { Forward = 0, Backward = 1 };

[[noreturn]] void abort_msg(const char *file, int line, const std::string &msg);
#define MRX_ABORT(msg) ::mrx::abort_msg(__FILE__, __LINE__, (msg))

inline int ipow(int b, int e) {
    int r = 1;
    for (int i = 0; i < e; i++) r *= b;
    return r;
}

// ---------------------------------------------------------------- tables
void set_table_path(const std::string &path); // mrcpp_b200/data/mwtables.bin
const std::string &table_path();

/// Two-scale filter blocks of order k. sub[op][idx] is a KxK row-major matrix F(i,j) used as
/// out[j] (+)= sum_i in[i] F(i,j); idx = 2*gt_bit + ft_bit (MWFilter::getSubFilter, MWFilter.cpp:100-133).
struct FilterSet {
    int k = 0, K = 0;
    std::vector<double> H0, G0, H1, G1; // row-major KxK
    std::vector<double> sub[2][4];
};
const FilterSet &filter_set(int k);

/// Cross-correlation tables (K*K) x (2K), row-major, |c|<MachinePrec zeroed (CrossCorrelation.cpp:99-124).
struct CrossCorr {
    int k = 0, K = 0;
    std::vector<double> L, R;
};
const CrossCorr &cross_corr(int k);
/// S_{+1}, S_0, S_{-1} (three K x K row-major matrices) of the tabulated derivative operators: kind 4/5 = Holoborodko order
/// 1/2 (PHCalculator.cpp:47-79), 6/7/8 = B-spline order 1/2/3 (BSCalculator.cpp:47-79)
const std::vector<double> &derivative_table(int kind, int k);

/// Gauss-Legendre points/weights on [0,1] (GaussQuadrature.cpp:157-193, QuadratureCache.cpp:41-45).
struct Quadrature {
    int n = 0;
    std::vector<double> roots, weights;
};
const Quadrature &quadrature(int n);

/// Interpolating scaling functions phi_j on [0,1] (InterpolatingBasis.cpp:62-84): value and derivative.
double interp_scaling_eval(int k, int j, double x);
double interp_scaling_deriv(int k, int j, double x);

// ---------------------------------------------------------------- MRA
template <int D> struct MRA {
    int order = 0;                 // polynomial order k
    int rootScale = 0;             // world box scale n (box length 2^-n)
    std::array<int, D> corner{};   // translation of the lower corner root box
    std::array<int, D> nboxes{};   // number of root boxes per dimension
    int maxDepth = MaxDepth;
    // periodic world (BoundingBox(..., pbc = true), BoundingBox.cpp:95-117): the unit cell is [-1, 1]^D in box units, i.e. root
    // scale 0, corner -1, two root boxes per dimension (periodic_utils.cpp:35-85 assumes exactly that)
    bool periodic = false;
    int maxScale() const { return rootScale + maxDepth; }
    int nRoots() const {
        int n = 1;
        for (int d = 0; d < D; d++) n *= nboxes[d];
        return n;
    }
    double lower(int d) const { return std::pow(2.0, -rootScale) * corner[d]; }
    double upper(int d) const { return std::pow(2.0, -rootScale) * (corner[d] + nboxes[d]); }
    bool operator==(const MRA &o) const {
        return order == o.order && rootScale == o.rootScale && corner == o.corner && nboxes == o.nboxes &&
               maxDepth == o.maxDepth && periodic == o.periodic;
    }
};

// ---------------------------------------------------------------- Hilbert path (HilbertPath.cpp:70-118)
int hilbert_child_path(int D, int path, int hIdx);
int hilbert_z_index(int D, int path, int hIdx);
int hilbert_h_index(int D, int path, int zIdx);

// ---------------------------------------------------------------- analytic Gaussians
/// coef * prod_d (x_d-pos_d)^pow_d * exp(-alpha sum_d (x_d-pos_d)^2)   (GaussFunc.cpp:47-66)
template <int D> struct GaussFunc {
    double coef = 1.0;
    double alpha = 1.0;
    std::array<double, D> pos{};
    std::array<int, D> power{};
    double evalf(const double *r) const {
        double q2 = 0.0, p2 = 1.0;
        for (int d = 0; d < D; d++) {
            double q = r[d] - pos[d];
            q2 += alpha * q * q;
            if (power[d] == 0) continue;
            if (power[d] == 1) p2 *= q;
            else p2 *= std::pow(q, power[d]);
        }
        if (q2 > 746.0) return 0.0 * coef * p2; // exp(-q2) underflows to exactly 0 in IEEE double
        return coef * p2 * std::exp(-q2);
    }
    // Gaussian.cpp:116-135
    bool isVisibleAtScale(int scale, int nQuadPts) const {
        double stdDeviation = std::pow(2.0 * alpha, -0.5);
        auto visibleScale = static_cast<int>(-std::floor(std::log2(nQuadPts * 0.5 * stdDeviation)));
        return !(scale < visibleScale);
    }
    bool isZeroOnInterval(const double *a, const double *b) const {
        for (int i = 0; i < D; i++) {
            double stdDeviation = std::pow(2.0 * alpha, -0.5);
            double gaussBoxMin = pos[i] - 5.0 * stdDeviation;
            double gaussBoxMax = pos[i] + 5.0 * stdDeviation;
            if (a[i] > gaussBoxMax or b[i] < gaussBoxMin) return true;
        }
        return false;
    }
};
template <int D> using GaussExp = std::vector<GaussFunc<D>>;

// coefficient-chunk allocator hooks (default: new[]/delete[]; the C-ABI installs cudaMallocHost/cudaFreeHost
// when a device is selected so that host<->device copies run at full PCIe/NVLink-C2C speed)
extern void *(*chunk_alloc)(size_t bytes);
extern void (*chunk_free)(void *p);

// ---------------------------------------------------------------- flat adaptive tree
enum NodeFlags : uint8_t { FlagBranch = 1, FlagGen = 2, FlagHasCoefs = 4, FlagEnd = 8 };

template <int D> struct NodeRec {
    int scale;
    std::array<int, D> l;
    int parent; // slot or -1
    int child0; // slot of child 0 (children contiguous) or -1
    uint8_t flags;
    uint8_t hpath;
};

/// Flat MW tree: nodes in creation order (roots in box order, then 2^D contiguous children per split).
template <int D> class Tree {
public:
    static constexpr int tdim = 1 << D;
    MRA<D> mra;
    int k, K, Kd, ncoef;
    int nRoots;
    double squareNorm = -1.0;
    bool operNorms = false; // OperatorTree: component norms are thresholded matrix norms
    double normPrec = 0.0;
    std::vector<NodeRec<D>> nodes;
    std::vector<double> cnorm; // tdim per node
    std::vector<double> sqn;   // per node
    int nReal = 0;             // nodes that are not generated (prefix of `nodes` once generation starts)

    explicit Tree(const MRA<D> &m);
    int size() const { return (int)nodes.size(); }
    double *coef(int n) { return chunks_[n >> chunkShift_] + (size_t)(n & chunkMask_) * ncoef; }
    const double *coef(int n) const { return chunks_[n >> chunkShift_] + (size_t)(n & chunkMask_) * ncoef; }
    ~Tree();
    Tree(const Tree &) = delete;
    Tree &operator=(const Tree &) = delete;
    int depth(int n) const { return nodes[n].scale - mra.rootScale; }
    bool isBranch(int n) const { return nodes[n].flags & FlagBranch; }
    bool isGen(int n) const { return nodes[n].flags & FlagGen; }
    bool isEnd(int n) const { return nodes[n].flags & FlagEnd; }
    int nDepths() const; // nodesAtDepth.size(): 1 + deepest depth holding a node

    int rootIndex(int scale, const std::array<int, D> &l) const; // BoundingBox.cpp:346-372 (-1 outside)
    int createChildren(int n, bool gen = false);                 // FunctionNode.cpp:255-291 / :293-328
    int findNode(int scale, const std::array<int, D> &l) const;  // never generates; -1 if absent
    int getNode(int scale, const std::array<int, D> &l);         // generates (MWTree.cpp:340-352)
    void deleteGenerated();                                      // MWNode.cpp:736-744
    /// topology-only variant of getNode for device-resident trees: missing nodes are created as
    /// generated nodes WITHOUT host coefficients; parents that got children are appended to newParents
    int getNodeTopo(int scale, const std::array<int, D> &l, std::vector<int> *newParents, bool withCoefStorage = false);
    void clearToRoots();          // FunctionTree::clear
    void copyGridFrom(const Tree<D> &other); // copy_grid (grid.cpp:150-166)
    void extendGridFrom(const Tree<D> &other); // build_grid(out, inp) (grid.cpp:144-153): union with the grid of `other`
    bool allocCoefs = true;       // false: new nodes get no host coefficient storage (device-resident)
    void ensureCoefStorage();     // allocate host storage for all nodes (before a download)
    void ensureCoefStorageFor(size_t nSlots); // ... for the first nSlots slots, whether or not the topology holds them yet
    /// host coefficient chunks (64 nodes each) for device-side gathers; pinned (device-readable) unless the tree was
    /// created before a CUDA device was selected
    const std::vector<double *> &coefChunks() const { return chunks_; }
    bool coefsPinned() const;
    static constexpr int chunkNodes = 64;
    /// move the coefficient chunks to another allocator (e.g. a host arena shared by the ranks of a sharded apply); contents kept
    void rebaseChunks(void *(*alloc)(size_t), void (*free)(void *));

    void zeroCoefs(int n);
    void calcNorms(int n);                  // MWNode.cpp:609-616 (+ OperatorNode.cpp:56-80 when operNorms)
    double scalingNorm(int n) const;        // squared, MWNode.cpp:619-627
    double waveletNorm(int n) const;        // squared, MWNode.cpp:630-640
    void mwTransformNode(int n, int op);    // MWNode.cpp:557-594
    void cvTransformBackward(int n);        // MWNode.cpp:448-490 (interpolating vcMap)
    void giveChildrenCoefs(int n, bool overwrite = true); // MWNode.cpp:312-335 (host: D<3 trees and gen nodes)
    void reCompress(int n);                 // MWNode.cpp:660-669
    void mwTransformUpSerial();             // OperatorTree.cpp:254-265
    void calcSquareNorm();                  // MWTree.cpp:109-118

    /// DFS pre-order, roots in box order, children in Hilbert order; gen nodes skipped (tree_utils.cpp:69-82).
    void nodeTable(std::vector<int> &table) const;
    void endNodeTable(std::vector<int> &table) const; // MWTree.cpp:454-462
    void nodeTableByDepth(std::vector<std::vector<int>> &table) const;

    void lowerBounds(int n, double *lb) const;
    void upperBounds(int n, double *ub) const;

private:
    static constexpr int chunkShift_ = 6, chunkMask_ = 63;
    std::vector<double *> chunks_; // 64 nodes each; pinned host memory when a CUDA device is in use
    void *(*alloc_)(size_t) = nullptr; // allocator pair captured at construction
    void (*free_)(void *) = nullptr;
    const FilterSet *fs_ = nullptr;
    int allocNodes(int count);
    void mwTransformCoefs(const double *in, double *out_children, bool readOnlyScaling, int stride, bool overwrite) const;
};

/// TreeBuilder::build (TreeBuilder.cpp:38-86) with calculator / adaptor given as callables.
/// initial work vector: end nodes (TreeCalculator.h:37) unless allNodes.
template <int D>
void build_tree(Tree<D> &tree, const std::function<void(Tree<D> &, int)> &calcNode,
                const std::function<bool(const Tree<D> &, int)> &splitNode, int maxIter, bool allNodes = false,
                bool parallel = true);

/// tree_utils::split_check (tree_utils.cpp:47-65)
template <int D> bool split_check(const Tree<D> &tree, int n, double prec, double splitFac, bool absPrec);

// ---------------------------------------------------------------- projection (input generator)
template <int D> void build_grid(Tree<D> &out, const GaussFunc<D> &f, int maxIter = -1);
template <int D> void build_grid(Tree<D> &out, const GaussExp<D> &f, int maxIter = -1);
/// ProjectionCalculator::calcNode for every work-vector node + adaptive refinement; `finalize`
/// runs the BottomUp transform on the host (only for D<3 operator-construction trees; the 3-D
/// function path calls the GPU transform through the C-ABI instead).
template <int D>
void project(double prec, Tree<D> &out, const std::function<double(const double *)> &f, int maxIter = -1,
             bool absPrec = false, bool finalize = true);

/// same as project() for f = sum of Gaussians, with exact node-level screening of terms that are identically zero
template <int D>
void project_gaussians(double prec, Tree<D> &out, const GaussExp<D> &gexp, int maxIter = -1, bool absPrec = false,
                       bool finalize = true);

// ---------------------------------------------------------------- operators
/// One separated term: the 2-D non-standard-form operator tree flattened to its [depth][transl] cache
/// (OperatorTree::setupOperNodeCache, OperatorTree.cpp:200-238).
struct OperTerm {
    int nDepth = 0;                // tree depth count (getDepth())
    std::vector<int> maxTransl;    // per depth
    std::vector<size_t> offset;    // per depth: index of node transl=-maxTransl in `mats` (units of nodes)
    std::vector<double> mats;      // nodes * 4*K*K  (block c at c*K*K; element p[i + K*m], i=input, m=output)
    std::vector<double> norms;     // nodes * 4      thresholded component norms (OperatorNode.cpp:56-80)
    const double *node(int depth, int transl) const {
        return mats.data() + (offset[depth] + (size_t)(transl + maxTransl[depth])) * matStride;
    }
    const double *nodeNorms(int depth, int transl) const {
        return norms.data() + (offset[depth] + (size_t)(transl + maxTransl[depth])) * 4;
    }
    size_t matStride = 0;
    // band widths [depth][T,C,B,A,max], -1 = empty (BandWidth.h:36-63); filled by calcBandWidths
    std::vector<std::array<int, 5>> bw;
    int bwDepth() const { return (int)bw.size() - 1; }
    int width(int depth, int c) const { return (depth > bwDepth() || depth < 0) ? -1 : bw[depth][c]; }
    int maxWidth(int depth) const { return width(depth, 4); }
};

struct Operator {
    int k = 0, K = 0;
    int operRoot = 0;
    int operReach = -10;   // MWOperator::oper_reach (MWOperator.h:75): >= 0 for operators built for a periodic world
    int order = 0;         // derivative order (0 for convolution operators)
    bool derivative = false;
    double buildPrec = 0.0;
    std::vector<OperTerm> terms;
    std::vector<int> bandMax; // per depth (MWOperator.cpp:78-108)
    int size() const { return (int)terms.size(); }
    void calcBandWidths(double prec); // MWOperator.cpp:78-108 + OperatorTree.cpp:109-134
    void clearBandWidths();
    int getMaxBandWidth(int depth = -1) const; // MWOperator.cpp:63-71
};

/// PoissonKernel (PoissonKernel.cpp:48-86) / HelmholtzKernel (HelmholtzKernel.cpp:47-80)
GaussExp<1> poisson_kernel(double epsilon, double r_min, double r_max);
GaussExp<1> helmholtz_kernel(double mu, double epsilon, double r_min, double r_max);

/// ConvolutionOperator::initialize (ConvolutionOperator.cpp:78-108) for a 3-D MRA.
/// root / reach: ConvolutionOperator(mra, kernel, prec, root, reach) (ConvolutionOperator.cpp:63-76); the defaults are those of
/// ConvolutionOperator(mra): root = the world's root scale, reach = -10 (operator boxes = world boxes)
Operator build_convolution_operator(const MRA<3> &mra, const GaussExp<1> &kernel, double k_prec, double o_prec, int oper_root,
                                    int oper_reach);
Operator build_convolution_operator(const MRA<3> &mra, const GaussExp<1> &kernel, double k_prec, double o_prec);
Operator build_poisson_operator(const MRA<3> &mra, double prec);                  // PoissonOperator.cpp:40-55
Operator build_helmholtz_operator(const MRA<3> &mra, double mu, double prec);     // HelmholtzOperator.cpp:44-59
/// PoissonOperator(mra, prec, root, reach) PoissonOperator.cpp:56-77 / HelmholtzOperator(mra, mu, prec, root, reach)
/// HelmholtzOperator.cpp:60-81: kernel precision prec / 100, r_max stretched over the reach (periodic worlds)
Operator build_poisson_operator(const MRA<3> &mra, double prec, int oper_root, int oper_reach);
Operator build_helmholtz_operator(const MRA<3> &mra, double mu, double prec, int oper_root, int oper_reach);
/// periodic::index_manipulation (periodic_utils.cpp:49-73) for scale >= 0: translation wrapped into the unit cell [-2^n, 2^n)
inline int periodic_wrap(int l, int scale) {
    const int two_n = 1 << (scale + 1);
    int t = l + two_n / 2;
    if (t >= two_n) t = t % two_n;
    if (t < 0) t = (t + 1) % two_n + two_n - 1;
    return t - two_n / 2;
}
/// periodic::in_unit_cell (periodic_utils.cpp:35-47)
inline bool periodic_in_unit_cell(const int l[3], int scale) {
    const int two_n = 1 << (scale + 1);
    for (int i = 0; i < 3; i++) {
        const int t = l[i] + two_n / 2;
        if (t >= two_n || t < 0) return false;
    }
    return true;
}
Operator build_abgv_operator(const MRA<3> &mra, double a, double b);
/// PHOperator<3>(mra, order) (PHOperator.cpp:40-69) / BSOperator<3>(mra, order) (BSOperator.cpp:40-66): bandwidth-1 derivative
/// operators from tabulated matrices
Operator build_ph_operator(const MRA<3> &mra, int order);
Operator build_bs_operator(const MRA<3> &mra, int order);              // ABGVOperator.cpp:52-74

double calc_min_distance(const MRA<3> &mra, double eps); // MultiResolutionAnalysis.h:65
double calc_max_distance(const MRA<3> &mra);             // MultiResolutionAnalysis.cpp:208

} // namespace mrx
