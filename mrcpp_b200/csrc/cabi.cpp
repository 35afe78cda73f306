// C-ABI entry points (include/mrcpp_b200.h). Host-side object management lives here; every hot-path
// call forwards to the CUDA engine and aborts when no device was selected (no CPU fallback).
#include <omp.h>

#include <algorithm>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "engine.hpp"

using namespace mrx;

namespace {
bool g_device_on = false;
int g_device = -1;
cudaStream_t g_stream = nullptr;
long long g_launches = 0;
std::mutex g_pin_mutex;
std::map<size_t, std::vector<void *>> g_pin_free;
std::map<void *, size_t> g_pin_size;
} // namespace

namespace {
// caching device allocator: size classes = powers of two up to 1 GiB, multiples of 1 GiB above
// (leaked on purpose: handles may be destroyed by the host language after static destructors ran)
std::mutex &g_dev_mutex = *new std::mutex;
std::map<size_t, std::vector<void *>> &g_dev_free = *new std::map<size_t, std::vector<void *>>;
std::map<void *, size_t> &g_dev_size = *new std::map<void *, size_t>;
size_t g_dev_cached = 0;
size_t dev_size_class(size_t bytes) {
    if (bytes < 512) return 512;
    const size_t GiB = (size_t)1 << 30;
    if (bytes > GiB) return (bytes + GiB - 1) / GiB * GiB;
    size_t c = 512;
    while (c < bytes) c <<= 1;
    return c;
}
} // namespace

namespace mrx {
void *dev_alloc(size_t bytes) {
    const size_t cls = dev_size_class(bytes);
    {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        auto &fl = g_dev_free[cls];
        if (!fl.empty()) {
            void *p = fl.back();
            fl.pop_back();
            g_dev_cached -= cls;
            return p;
        }
    }
    void *p = nullptr;
    if (cudaMalloc(&p, cls) != cudaSuccess) {
        // give the cached blocks back to the driver and retry once
        {
            std::lock_guard<std::mutex> lock(g_dev_mutex);
            cudaDeviceSynchronize();
            for (auto &kv : g_dev_free) {
                for (void *q : kv.second) {
                    cudaFree(q);
                    g_dev_size.erase(q);
                }
                kv.second.clear();
            }
            g_dev_cached = 0;
        }
        cudaGetLastError();
        if (cudaMalloc(&p, cls) != cudaSuccess) MRX_ABORT("cudaMalloc failed (out of device memory)");
    }
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    g_dev_size[p] = cls;
    return p;
}
void dev_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    auto it = g_dev_size.find(p);
    if (it == g_dev_size.end()) return;
    g_dev_free[it->second].push_back(p);
    g_dev_cached += it->second;
}
size_t dev_cached_bytes() { return g_dev_cached; }
bool device_enabled() { return g_device_on; }
void require_device(const char *what) {
    if (!g_device_on) MRX_ABORT(std::string(what) + ": no CUDA device selected (mrx_init device<0); there is no CPU fallback");
}
cudaStream_t stream() { return g_stream; }
long long &launch_counter() { return g_launches; }
} // namespace mrx

extern "C" {

const char *mrx_version(void) { return "mrcpp_b200 0.1 (sm_100a)"; }

int mrx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mrx_init(const char *table_path, int device) {
    if (table_path && table_path[0]) set_table_path(table_path);
    if (device >= 0) {
        int n = mrx_device_count();
        if (device >= n) {
            std::fprintf(stderr, "mrx_init: CUDA device %d requested but %d visible\n", device, n);
            return 1;
        }
        if (cudaSetDevice(device) != cudaSuccess) return 2;
        if (!g_stream) {
            if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) return 3;
        }
        g_device = device;
        g_device_on = true;
        // pinned coefficient chunks from now on
        // pinned coefficient chunks, recycled through a free list (cudaMallocHost is slow)
        chunk_alloc = [](size_t bytes) -> void * {
            {
                std::lock_guard<std::mutex> lock(g_pin_mutex);
                auto &fl = g_pin_free[bytes];
                if (!fl.empty()) {
                    void *p = fl.back();
                    fl.pop_back();
                    return p;
                }
            }
            void *p = nullptr;
            if (cudaMallocHost(&p, bytes) != cudaSuccess) MRX_ABORT("cudaMallocHost failed");
            std::lock_guard<std::mutex> lock(g_pin_mutex);
            g_pin_size[p] = bytes;
            return p;
        };
        chunk_free = [](void *p) {
            std::lock_guard<std::mutex> lock(g_pin_mutex);
            g_pin_free[g_pin_size[p]].push_back(p);
        };
        // device pool: keep freed memory cached
        {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
    }
    return 0;
}

mrx_mra *mrx_mra_create(int order, int root_scale, const int corner[3], const int nboxes[3], int max_depth) {
    auto *m = new mrx_mra;
    m->m.order = order;
    m->m.rootScale = root_scale;
    for (int d = 0; d < 3; d++) {
        m->m.corner[d] = corner[d];
        m->m.nboxes[d] = nboxes[d];
    }
    m->m.maxDepth = max_depth;
    if (order < 1 || order > MaxOrder) MRX_ABORT("Invalid scaling order");
    if (m->m.maxDepth > MaxDepth) MRX_ABORT("Beyond MaxDepth");
    if (m->m.maxScale() > MaxScale) MRX_ABORT("Beyond MaxScale");
    return m;
}
int mrx_mra_set_periodic(mrx_mra *mra, int periodic) {
    // periodic_utils.cpp:35-85 works on the unit cell [-1, 1]^3 in box units: root scale 0, corner -1, two boxes per dimension
    if (periodic) {
        if (mra->m.rootScale != 0) MRX_ABORT("periodic world: root scale must be 0");
        for (int d = 0; d < 3; d++)
            if (mra->m.corner[d] != -1 || mra->m.nboxes[d] != 2) MRX_ABORT("periodic world: corner -1 and two boxes per dimension");
    }
    mra->m.periodic = periodic != 0;
    return 0;
}
void mrx_mra_destroy(mrx_mra *mra) { delete mra; }

mrx_tree *mrx_tree_create(const mrx_mra *mra) { return new mrx_tree(mra->m); }
void mrx_tree_destroy(mrx_tree *tree) { delete tree; }
int mrx_tree_n_nodes(const mrx_tree *tree) { return tree->host.nReal; }
int mrx_tree_n_end_nodes(const mrx_tree *tree) {
    std::vector<int> e;
    tree->host.endNodeTable(e);
    return (int)e.size();
}
double mrx_tree_square_norm(const mrx_tree *tree) { return tree->host.squareNorm; }
void mrx_tree_clear(mrx_tree *tree) {
    tree->host.deleteGenerated();
    tree->host.clearToRoots();
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.nNodes = 0;
    tree->dev.nGen = 0;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
}
long long mrx_tree_bytes(const mrx_tree *tree) { return (long long)tree->host.nReal * tree->host.ncoef * 8; }

mrx_tree *mrx_tree_from_arrays(const mrx_mra *mra, int n_nodes, const int *scale, const int *transl, const int *parent,
                               const int *child0, const double *coefs) {
    auto *t = new mrx_tree(mra->m);
    Tree<3> &h = t->host;
    if (n_nodes < h.nRoots) MRX_ABORT("tree_from_arrays: fewer nodes than root boxes");
    // replay the splits in slot order of the children
    std::vector<std::pair<int, int>> splits;
    for (int n = 0; n < n_nodes; n++)
        if (child0[n] >= 0) splits.push_back({child0[n], n});
    std::sort(splits.begin(), splits.end());
    // coefficients given: storage for all nodes in one go and without zeroing it first (every block is overwritten below)
    const bool allocWas = h.allocCoefs;
    if (coefs) h.allocCoefs = false;
    for (auto &s : splits) {
        if (s.second >= h.size()) MRX_ABORT("tree_from_arrays: parent slot after its children");
        int c0 = h.createChildren(s.second, false);
        if (c0 != s.first) MRX_ABORT("tree_from_arrays: children must be contiguous in creation order");
    }
    h.allocCoefs = allocWas;
    if (h.size() != n_nodes) MRX_ABORT("tree_from_arrays: node count mismatch");
    if (coefs) h.ensureCoefStorageFor((size_t)n_nodes);
    for (int n = 0; n < n_nodes; n++) {
        if (h.nodes[n].scale != scale[n]) MRX_ABORT("tree_from_arrays: scale mismatch");
        for (int d = 0; d < 3; d++)
            if (h.nodes[n].l[d] != transl[3 * n + d]) MRX_ABORT("tree_from_arrays: translation mismatch");
        if (parent && h.nodes[n].parent != parent[n]) MRX_ABORT("tree_from_arrays: parent mismatch");
    }
    if (coefs) {
        // the copy a binding pays per tree it hands over (32 KB per node at k = 7): all host cores, node by node (norms while the
        // node is in cache)
#pragma omp parallel for schedule(static, 16)
        for (int n = 0; n < n_nodes; n++) {
            std::memcpy(h.coef(n), coefs + (size_t)n * h.ncoef, sizeof(double) * h.ncoef);
            h.nodes[n].flags |= FlagHasCoefs;
            h.calcNorms(n);
        }
        h.calcSquareNorm();
    }
    return t;
}

int mrx_tree_to_arrays(mrx_tree *tree, int *scale, int *transl, int *parent, int *child0, double *coefs, double *norms) {
    Tree<3> &h = tree->host;
    if (coefs && !tree->hostCoefsValid) mrx_tree_sync_host(tree);
#pragma omp parallel for schedule(static, 16)
    for (int n = 0; n < h.nReal; n++) {
        if (scale) scale[n] = h.nodes[n].scale;
        if (transl)
            for (int d = 0; d < 3; d++) transl[3 * n + d] = h.nodes[n].l[d];
        if (parent) parent[n] = h.nodes[n].parent;
        if (child0) child0[n] = (h.nodes[n].child0 >= 0 && h.nodes[n].child0 < h.nReal) ? h.nodes[n].child0 : -1;
        if (coefs) std::memcpy(coefs + (size_t)n * h.ncoef, h.coef(n), sizeof(double) * h.ncoef);
        if (norms)
            for (int t = 0; t < 8; t++) norms[(size_t)n * 8 + t] = h.cnorm[(size_t)n * 8 + t];
    }
    return 0;
}

int mrx_tree_copy_grid(mrx_tree *out, const mrx_tree *inp) {
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    out->host.copyGridFrom(inp->host);
    out->hostCoefsValid = true;
    out->devValid = false;
    out->dev.nNodes = 0;
    out->dev.topoNodes = -1;
    out->dev.partial = false;
    return 0;
}

// FunctionTree::integrate (src/trees/FunctionTree.cpp:438-454) with FunctionNode::integrateInterpolating
// (src/trees/FunctionNode.cpp:128-157): sum over the root nodes of 2^(-3 n / 2) * sum_ijk s_ijk sqrt(w_i w_j w_k). Only the
// scaling blocks of the root nodes are needed: they are read back from HBM when the host copy is not current.
double mrx_tree_integrate(mrx_tree *tree) {
    Tree<3> &h = tree->host;
    const int K = h.K, Kd = h.Kd;
    std::vector<double> roots((size_t)h.nRoots * Kd, 0.0);
    if (tree->hostCoefsValid) {
        for (int r = 0; r < h.nRoots; r++)
            if (h.nodes[r].flags & FlagHasCoefs) std::memcpy(roots.data() + (size_t)r * Kd, h.coef(r), sizeof(double) * Kd);
    } else {
        require_device("mrx_tree_integrate (root blocks live in HBM)");
        if (!tree->devValid || tree->dev.nNodes < h.nRoots) MRX_ABORT("mrx_tree_integrate: no current copy of the coefficients");
        if (cudaMemcpy2DAsync(roots.data(), sizeof(double) * Kd, tree->dev.coefs.p, sizeof(double) * h.ncoef, sizeof(double) * Kd,
                              h.nRoots, cudaMemcpyDeviceToHost, g_stream) != cudaSuccess ||
            cudaStreamSynchronize(g_stream) != cudaSuccess)
            MRX_ABORT("mrx_tree_integrate: device->host copy of the root blocks failed");
    }
    const Quadrature &q = quadrature(K);
    std::vector<double> sw(K);
    for (int i = 0; i < K; i++) sw[i] = std::sqrt(q.weights[i]);
    double result = 0.0;
    for (int r = 0; r < h.nRoots; r++) {
        const double *c = roots.data() + (size_t)r * Kd;
        double sum = 0.0;
        for (int z = 0; z < K; z++)
            for (int y = 0; y < K; y++)
                for (int x = 0; x < K; x++) sum += ((c[x + K * (y + K * z)] * sw[x]) * sw[y]) * sw[z];
        result += std::pow(2.0, -(3 * h.nodes[r].scale) / 2.0) * sum;
    }
    return result;
}

// FunctionTree::evalf / evalf_precise (src/trees/FunctionTree.cpp:374-436) with FunctionNode::evalScaling / evalf
// (src/trees/FunctionNode.cpp:47-93): the end node that holds the point evaluates its scaling block (evalf), or -- precise -- the
// scaling block of the child it generates from its full coefficients. Host arithmetic on the downloaded tree: reading values
// back is a boundary operation, not part of the hot path.
namespace {
double eval_scaling(const Tree<3> &h, int n, const double r[3]) {
    const int K = h.K;
    const double nf = std::pow(2.0, h.nodes[n].scale);
    double val[3][MaxOrder + 1];
    for (int d = 0; d < 3; d++) {
        const double x = r[d] * nf - static_cast<double>(h.nodes[n].l[d]);
        for (int j = 0; j < K; j++) val[d][j] = (x < 0.0 || x > 1.0) ? 0.0 : interp_scaling_eval(h.k, j, x);
    }
    const double *c = h.coef(n);
    double result = 0.0;
    for (int z = 0; z < K; z++)
        for (int y = 0; y < K; y++)
            for (int x = 0; x < K; x++) result += ((c[x + K * (y + K * z)] * val[0][x]) * val[1][y]) * val[2][z];
    return std::pow(2.0, (3 * h.nodes[n].scale) / 2.0) * result;
}
} // namespace

int mrx_tree_evalf(mrx_tree *tree, int n_points, const double *r, double *values, int precise) {
    if (!tree->hostCoefsValid) mrx_tree_sync_host(tree);
    Tree<3> &h = tree->host;
    const double unit = std::pow(2.0, -h.mra.rootScale);
    for (int p = 0; p < n_points; p++) {
        const double *x = r + 3 * (size_t)p;
        double xw[3];
        if (h.mra.periodic) { // periodic::coord_manipulation (periodic_utils.cpp:75-85): the point mapped into the unit cell
            for (int d = 0; d < 3; d++) {
                const double lo = h.mra.lower(d), len = h.mra.upper(d) - h.mra.lower(d);
                double t = std::fmod(x[d] - lo, len);
                if (t < 0.0) t += len;
                xw[d] = lo + t;
            }
            x = xw;
        }
        // outside the world the function is zero (FunctionTree.cpp:386); root box as BoundingBox::getBoxIndex(Coord)
        int n = 0, cells = 1;
        bool outside = false;
        for (int d = 0; d < 3; d++) {
            if (x[d] < h.mra.lower(d) || x[d] >= h.mra.upper(d)) outside = true;
            if (outside) break;
            double iint;
            std::modf((x[d] - h.mra.lower(d)) / unit, &iint);
            n += cells * (int)iint;
            cells *= h.mra.nboxes[d];
        }
        if (outside) {
            values[p] = 0.0;
            continue;
        }
        // MWNode::retrieveNodeOrEndNode(r): descend by MWNode::getChildIndex(r) (MWNode.cpp:804-815)
        auto child_of = [&h](int node, const double *pt) {
            const double sFac = std::pow(2.0, -h.nodes[node].scale);
            int c = 0;
            for (int d = 0; d < 3; d++)
                if (pt[d] > sFac * (h.nodes[node].l[d] + 0.5)) c += 1 << d;
            return c;
        };
        while (h.isBranch(n) && !h.isGen(h.nodes[n].child0)) n = h.nodes[n].child0 + child_of(n, x);
        if (!(h.nodes[n].flags & FlagHasCoefs)) MRX_ABORT("Evaluating node without coefs");
        if (!precise) {
            values[p] = eval_scaling(h, n, x);
        } else {
            const int c = child_of(n, x);
            std::array<int, 3> l;
            for (int d = 0; d < 3; d++) l[d] = 2 * h.nodes[n].l[d] + ((c >> d) & 1);
            const bool saved = h.allocCoefs; // trees born on the device create nodes without host storage
            h.allocCoefs = true;
            const int m = h.getNode(h.nodes[n].scale + 1, l); // generates the children from the node's full coefficients
            values[p] = eval_scaling(h, m, x);
            h.deleteGenerated();
            h.allocCoefs = saved;
        }
    }
    return 0;
}

// ---- text interchange: FunctionTree::saveTreeTXT / loadTreeTXT (src/trees/FunctionTree.cpp:240-372). Host arithmetic on the
// downloaded tree. The file holds, per end node, the function values at the quadrature points of its eight children
// (mwTransform(Reconstruction) + cvTransform(Forward)), child by child, with MADNESS conventions: level counted from the box
// [-L, L]^3, translations from 0, index order z fastest and descending.
namespace {
void cv_map_node(Tree<3> &h, int n, bool forward) { // MWNode::cvTransform for the interpolating basis (MWNode.cpp:448-490)
    const int K = h.K;
    const Quadrature &q = quadrature(K);
    std::vector<double> m(K);
    for (int j = 0; j < K; j++) m[j] = forward ? std::sqrt(1.0 / q.weights[j]) : std::sqrt(q.weights[j]);
    const double two = std::pow(2.0, 3 * (h.nodes[n].scale + 1));
    const double two_fac = forward ? std::sqrt(two) : std::sqrt(1.0 / two);
    double *c = h.coef(n);
    for (int b = 0; b < 8; b++)
        for (int idx = 0; idx < h.Kd; idx++) {
            double v = c[(size_t)b * h.Kd + idx];
            v = ((v * m[idx % K]) * m[(idx / K) % K]) * m[idx / (K * K)];
            c[(size_t)b * h.Kd + idx] = two_fac * v;
        }
}
} // namespace

int mrx_tree_save_txt(mrx_tree *tree, const char *path) {
    if (!tree->hostCoefsValid) mrx_tree_sync_host(tree);
    Tree<3> &h = tree->host;
    FILE *f = std::fopen(path, "w");
    if (!f) MRX_ABORT(std::string("cannot open ") + path);
    const int K = h.K, Kd = h.Kd, rscale = h.mra.rootScale;
    // the format assumes the world [-L, L]^3 with two root boxes per direction (FunctionTree.cpp:313): any other world would be
    // written with a wrong header and shifted translations -- refuse instead
    for (int d = 0; d < 3; d++)
        if (h.mra.corner[d] != -1 || h.mra.nboxes[d] != 2) MRX_ABORT("saveTreeTXT: the text format needs the world [-L, L]^3 (corner -1, 2 boxes per dimension)");
    if (rscale > 0) MRX_ABORT("saveTreeTXT: the text format needs a root scale <= 0");
    double Lw = 1.0;
    for (int i = 0; i > rscale; i--) Lw *= 2;
    std::fprintf(f, "3\n");
    for (int d = 0; d < 3; d++) std::fprintf(f, "%.14g %.14g\n", -Lw, Lw);
    std::vector<int> ends;
    h.endNodeTable(ends);
    std::fprintf(f, "%d\n%d\n", K, 8 * (int)ends.size());
    std::vector<int> map; // MADNESS index order (FunctionTree.cpp:335-342)
    for (int x = K - 1; x >= 0; x--)
        for (int y = K - 1; y >= 0; y--)
            for (int z = K - 1; z >= 0; z--) map.push_back(z * K * K + y * K + x);
    const int L = (int)std::pow(2.0, -rscale);
    std::vector<double> keep(h.ncoef);
    for (int n : ends) {
        std::memcpy(keep.data(), h.coef(n), sizeof(double) * h.ncoef);
        h.mwTransformNode(n, Reconstruction);
        cv_map_node(h, n, true);
        const int s = h.nodes[n].scale;
        for (int c = 0; c < 8; c++) {
            std::fprintf(f, "%d ", s - rscale + 2);
            for (int d = 0; d < 3; d++) std::fprintf(f, "%d ", (int)(2 * (h.nodes[n].l[d] + std::pow(2.0, s) * L)) + ((c >> d) & 1));
            std::fprintf(f, "\n");
            const double *v = h.coef(n) + (size_t)c * Kd;
            for (int i = 0; i < Kd; i++) std::fprintf(f, "%.14g ", v[map[i]]);
            std::fprintf(f, "\n");
        }
        std::memcpy(h.coef(n), keep.data(), sizeof(double) * h.ncoef); // the node keeps its coefficients
    }
    std::fclose(f);
    return 0;
}

int mrx_tree_load_txt(mrx_tree *tree, const char *path) {
    Tree<3> &h = tree->host;
    FILE *f = std::fopen(path, "r");
    if (!f) MRX_ABORT(std::string("cannot open ") + path);
    int D = 0, K = 0, nblk = 0;
    if (std::fscanf(f, "%d", &D) != 1 || D != 3) MRX_ABORT("load_txt: not a 3-D tree file");
    const int rscale = h.mra.rootScale, Kd = h.Kd;
    // loadTreeTXT (FunctionTree.cpp:240-262) insists on the world [-L, L]^3 of THIS tree in all three dimensions, L = 2^(-root scale)
    for (int d = 0; d < 3; d++)
        if (h.mra.corner[d] != -1 || h.mra.nboxes[d] != 2) MRX_ABORT("loadTreeTXT: the text format needs the world [-L, L]^3 (corner -1, 2 boxes per dimension)");
    if (rscale > 0) MRX_ABORT("loadTreeTXT: the text format needs a root scale <= 0");
    double Lw = 1.0;
    for (int i = 0; i > rscale; i--) Lw *= 2;
    for (int d = 0; d < 3; d++) {
        double lo = 0.0, hi = 0.0;
        if (std::fscanf(f, "%lf %lf", &lo, &hi) != 2) MRX_ABORT("load_txt: bad header");
        if (std::abs(lo + Lw) > 1e-12 * Lw || std::abs(hi - Lw) > 1e-12 * Lw) MRX_ABORT("load_txt: world of the file differs from the tree's");
    }
    if (std::fscanf(f, "%d %d", &K, &nblk) != 2 || K != h.K) MRX_ABORT("load_txt: polynomial order of the file differs from the tree's");
    const int L = (int)std::pow(2.0, -rscale);
    std::vector<int> map;
    for (int x = K - 1; x >= 0; x--)
        for (int y = K - 1; y >= 0; y--)
            for (int z = K - 1; z >= 0; z--) map.push_back(z * K * K + y * K + x);
    h.deleteGenerated();
    h.clearToRoots();
    h.allocCoefs = true;
    h.ensureCoefStorage();
    std::vector<double> vals(Kd);
    std::vector<unsigned char> fileMask; // per node: which of its eight child blocks the file has delivered
    for (int b = 0; b < nblk; b++) {
        int lev = 0, lm[3];
        if (std::fscanf(f, "%d %d %d %d", &lev, &lm[0], &lm[1], &lm[2]) != 4) MRX_ABORT("load_txt: truncated file");
        for (int i = 0; i < Kd; i++)
            if (std::fscanf(f, "%lf", &vals[i]) != 1) MRX_ABORT("load_txt: truncated file");
        const int cscale = lev + rscale - 1; // scale of the child the block belongs to
        std::array<int, 3> lc, lp;
        int c = 0;
        for (int d = 0; d < 3; d++) {
            lc[d] = lm[d] - (int)(std::pow(2.0, cscale) * L);
            lp[d] = lc[d] >> 1;
            c |= (lc[d] & 1) << d;
        }
        // the node the block belongs to (parent of the block's box): created on the way down if the grid does not have it yet
        int n = h.rootIndex(cscale - 1, lp);
        if (n < 0) MRX_ABORT("load_txt: block outside the world");
        while (h.nodes[n].scale < cscale - 1) {
            if (h.nodes[n].child0 < 0) {
                const int c0 = h.createChildren(n, false);
                for (int q = 0; q < 8; q++) std::memset(h.coef(c0 + q), 0, sizeof(double) * h.ncoef);
            }
            const int shift = cscale - 1 - h.nodes[n].scale - 1;
            int k = 0;
            for (int d = 0; d < 3; d++) k |= ((lp[d] >> shift) & 1) << d;
            n = h.nodes[n].child0 + k;
        }
        if ((int)fileMask.size() < h.size()) fileMask.resize(h.size(), 0);
        if (fileMask[n] & (1u << c)) MRX_ABORT("load_txt: the same block twice");
        double *dst = h.coef(n) + (size_t)c * Kd;
        for (int i = 0; i < Kd; i++) dst[map[i]] = vals[i];
        fileMask[n] |= (unsigned char)(1u << c);
    }
    std::fclose(f);
    fileMask.resize(h.size(), 0);
    // quadrature-point values -> scaling coefficients of the children (cvTransform(Backward), FunctionTree.cpp:264-271); blocks the
    // file did not deliver are zero and stay zero under the (diagonal) map
    for (int n = 0; n < h.size(); n++)
        if (fileMask[n]) cv_map_node(h, n, false);
    // bottom-up (FunctionTree.cpp:273-303): a complete end node is compressed; a node with finer data below some of its children
    // takes their scaling blocks, hands the blocks the file gave it directly to the children the file left undefined (scaling =
    // the block, wavelets zero: they become end nodes), and is compressed itself -- the MADNESS convention allows sibling groups
    // whose members end at different depths
    std::vector<int> order(h.size());
    for (int n = 0; n < h.size(); n++) order[n] = n;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h.nodes[a].scale > h.nodes[b].scale; });
    std::vector<char> defined(h.size(), 0);
    for (int n : order) {
        if (!h.isBranch(n)) {
            if (fileMask[n] == 0) continue; // an undefined child: its parent deals with it (or nothing refers to it: checked below)
            if (fileMask[n] != 0xFF) MRX_ABORT("load_txt: incomplete sibling group at the finest level of a branch");
            h.mwTransformNode(n, Compression);
            h.nodes[n].flags |= FlagHasCoefs;
            h.calcNorms(n);
            defined[n] = 1;
            continue;
        }
        const int c0 = h.nodes[n].child0;
        for (int c = 0; c < 8; c++) {
            double *blk = h.coef(n) + (size_t)c * Kd;
            if (fileMask[n] & (1u << c)) {
                // the file gave this child's values at this level: the child must not carry finer data
                if (defined[c0 + c]) MRX_ABORT("load_txt: values for a box that also has finer boxes (not a tree file)");
                double *cc = h.coef(c0 + c);
                std::memcpy(cc, blk, sizeof(double) * Kd);
                std::memset(cc + Kd, 0, sizeof(double) * (size_t)7 * Kd);
                h.nodes[c0 + c].flags |= FlagHasCoefs;
                h.calcNorms(c0 + c);
                defined[c0 + c] = 1;
            } else {
                if (!defined[c0 + c]) MRX_ABORT("load_txt: incomplete sibling group (a box without values)");
                std::memcpy(blk, h.coef(c0 + c), sizeof(double) * Kd); // the child's scaling block
            }
        }
        h.mwTransformNode(n, Compression);
        h.nodes[n].flags |= FlagHasCoefs;
        h.calcNorms(n);
        defined[n] = 1;
    }
    for (int n = 0; n < h.size(); n++)
        if (!defined[n]) MRX_ABORT("load_txt: incomplete tree (a node without values)");
    h.calcSquareNorm();
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.nNodes = 0;
    tree->dev.nGen = 0;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
    return 0;
}

// build_grid(out, GaussExp) alone (src/treebuilders/grid.cpp:78-123): refine the grid where the Gaussians are visible, no
// coefficients. Host only.
int mrx_build_grid_gaussians(mrx_tree *tree, int n_gauss, const double *coef, const double *alpha, const double *pos,
                             const int *power, int max_iter) {
    GaussExp<3> gexp(n_gauss);
    for (int i = 0; i < n_gauss; i++) {
        gexp[i].coef = coef[i];
        gexp[i].alpha = alpha[i];
        for (int d = 0; d < 3; d++) {
            gexp[i].pos[d] = pos[3 * i + d];
            gexp[i].power[d] = power ? power[3 * i + d] : 0;
        }
    }
    build_grid<3>(tree->host, gexp, max_iter);
    tree->host.squareNorm = -1.0; // a bare grid from here on
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.nNodes = 0;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
    return 0;
}

int mrx_project_gaussians(mrx_tree *tree, double prec, int n_gauss, const double *coef, const double *alpha,
                          const double *pos, const int *power, int do_build_grid, int finalize) {
    if (finalize) require_device("mrx_project_gaussians (final BottomUp transform)");
    GaussExp<3> gexp(n_gauss);
    for (int i = 0; i < n_gauss; i++) {
        gexp[i].coef = coef[i];
        gexp[i].alpha = alpha[i];
        for (int d = 0; d < 3; d++) {
            gexp[i].pos[d] = pos[3 * i + d];
            gexp[i].power[d] = power ? power[3 * i + d] : 0;
        }
    }
    Tree<3> &h = tree->host;
    h.allocCoefs = true; // host quadrature writes host coefficient chunks
    h.ensureCoefStorage();
    if (do_build_grid) build_grid<3>(h, gexp, -1);
    project_gaussians<3>(prec, h, gexp, -1, false, /*finalize=*/false);
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
    // project.cpp:96-97: out.mwTransform(BottomUp); out.calcSquareNorm() -- on the device
    if (finalize) {
        mrx_mw_transform(tree, MRX_BOTTOM_UP, 1);
        mrx_calc_square_norm(tree);
    }
    return 0;
}

int mrx_project_function(mrx_tree *tree, double prec, mrx_func3 f, void *user, int threads_ok, int finalize) {
    if (finalize) require_device("mrx_project_function (final BottomUp transform)");
    if (!f) MRX_ABORT("mrx_project_function: null callback");
    Tree<3> &h = tree->host;
    h.allocCoefs = true; // host quadrature writes host coefficient chunks
    h.ensureCoefStorage();
    // ProjectionCalculator::calcNode on the host (quadrature at the expanded child points); callbacks that are not thread safe
    // (e.g. a host-language closure) are called from one thread only
    if (threads_ok) {
        project<3>(prec, h, [f, user](const double *r) { return f(r, user); }, -1, false, /*finalize=*/false);
    } else {
        const int saved = omp_get_max_threads();
        omp_set_num_threads(1);
        project<3>(prec, h, [f, user](const double *r) { return f(r, user); }, -1, false, /*finalize=*/false);
        omp_set_num_threads(saved);
    }
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
    if (finalize) {
        mrx_mw_transform(tree, MRX_BOTTOM_UP, 1);
        mrx_calc_square_norm(tree);
    }
    return 0;
}

namespace {
struct CosineSum {
    int n;
    const double *amp, *k; // f(r) = sum_i amp[i] prod_d cos(pi k[3 i + d] r_d)
};
double cosine_sum_eval(const double *r, void *user) {
    const CosineSum *c = static_cast<const CosineSum *>(user);
    double s = 0.0;
    for (int i = 0; i < c->n; i++) {
        double p = c->amp[i];
        for (int d = 0; d < 3; d++) p *= std::cos(mrx::pi * c->k[3 * i + d] * r[d]);
        s += p;
    }
    return s;
}
} // namespace
int mrx_project_cosines(mrx_tree *tree, double prec, int n_terms, const double *amp, const double *kvec, int finalize) {
    CosineSum c{n_terms, amp, kvec};
    return mrx_project_function(tree, prec, cosine_sum_eval, &c, 1, finalize);
}

int mrx_project_gaussians_device(mrx_tree *tree, double prec, int n_gauss, const double *coef, const double *alpha,
                                 const double *pos, const int *power, int do_build_grid) {
    require_device("mrx_project_gaussians_device");
    GaussExp<3> gexp(n_gauss);
    for (int i = 0; i < n_gauss; i++) {
        gexp[i].coef = coef[i];
        gexp[i].alpha = alpha[i];
        for (int d = 0; d < 3; d++) {
            gexp[i].pos[d] = pos[3 * i + d];
            gexp[i].power[d] = power ? power[3 * i + d] : 0;
        }
    }
    Tree<3> &h = tree->host;
    h.allocCoefs = false; // device-resident from the first node on
    if (do_build_grid) build_grid<3>(h, gexp, -1);
    device_project_gaussians(*tree, prec, gexp, -1, false);
    return 0;
}

// ---- operators
} // extern "C"

// ---- operator cache: an SCF loop constructs PoissonOperator / HelmholtzOperator(mu) objects again and again
// (examples/scf.cpp:102); construction (kernel projection + cross correlation of every term, ConvolutionOperator.cpp:78-108)
// costs 30-200 ms on the host, a copy of the finished tables 1-2 ms. The cache holds the HOST tables of the last operators built,
// keyed by every parameter they depend on; a handle created from it owns its own copy (band widths are per-apply state of a
// handle). MRX_OPER_CACHE=0 switches it off.
namespace {
struct OperCache {
    std::mutex mu;
    std::list<std::pair<std::string, std::shared_ptr<const Operator>>> lru;
    long long hits = 0, misses = 0;
};
OperCache &oper_cache() {
    static OperCache c;
    return c;
}
std::string mra_key(const MRA<3> &m) {
    char b[160];
    std::snprintf(b, sizeof(b), "k%d n%d c%d,%d,%d b%d,%d,%d d%d p%d", m.order, m.rootScale, m.corner[0], m.corner[1], m.corner[2], m.nboxes[0],
                  m.nboxes[1], m.nboxes[2], m.maxDepth, m.periodic ? 1 : 0);
    return b;
}
template <typename F> Operator cached_operator(const std::string &key, F build) {
    static const bool on = !(getenv("MRX_OPER_CACHE") && getenv("MRX_OPER_CACHE")[0] == '0');
    OperCache &c = oper_cache();
    if (on) {
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto it = c.lru.begin(); it != c.lru.end(); ++it)
            if (it->first == key) {
                c.lru.splice(c.lru.begin(), c.lru, it);
                c.hits++;
                return *c.lru.front().second; // copy
            }
    }
    Operator op = build();
    if (on) {
        std::lock_guard<std::mutex> lk(c.mu);
        c.misses++;
        c.lru.emplace_front(key, std::make_shared<const Operator>(op));
        while (c.lru.size() > 16) c.lru.pop_back();
    }
    return op;
}
std::string num_key(double v) {
    char b[40];
    std::snprintf(b, sizeof(b), " %.17g", v);
    return b;
}
} // namespace

extern "C" {
void mrx_oper_cache_stats(long long *hits, long long *misses) {
    OperCache &c = oper_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    if (hits) *hits = c.hits;
    if (misses) *misses = c.misses;
}
mrx_oper *mrx_poisson_create(const mrx_mra *mra, double prec) {
    auto *o = new mrx_oper;
    o->op = cached_operator("poisson " + mra_key(mra->m) + num_key(prec), [&] { return build_poisson_operator(mra->m, prec); });
    return o;
}
mrx_oper *mrx_helmholtz_create(const mrx_mra *mra, double mu, double prec) {
    auto *o = new mrx_oper;
    o->op = cached_operator("helmholtz " + mra_key(mra->m) + num_key(mu) + num_key(prec), [&] { return build_helmholtz_operator(mra->m, mu, prec); });
    return o;
}
mrx_oper *mrx_convolution_create(const mrx_mra *mra, int n_terms, const double *coef, const double *expo, double prec) {
    GaussExp<1> kernel(n_terms);
    for (int i = 0; i < n_terms; i++) {
        kernel[i].coef = coef[i];
        kernel[i].alpha = expo[i];
    }
    auto *o = new mrx_oper;
    o->op = build_convolution_operator(mra->m, kernel, prec / 10.0, prec);
    return o;
}
mrx_oper *mrx_poisson_create_reach(const mrx_mra *mra, double prec, int root, int reach) {
    auto *o = new mrx_oper;
    o->op = cached_operator("poisson_reach " + mra_key(mra->m) + num_key(prec) + num_key(root) + num_key(reach),
                            [&] { return build_poisson_operator(mra->m, prec, root, reach); });
    return o;
}
mrx_oper *mrx_helmholtz_create_reach(const mrx_mra *mra, double mu, double prec, int root, int reach) {
    auto *o = new mrx_oper;
    o->op = cached_operator("helmholtz_reach " + mra_key(mra->m) + num_key(mu) + num_key(prec) + num_key(root) + num_key(reach),
                            [&] { return build_helmholtz_operator(mra->m, mu, prec, root, reach); });
    return o;
}
mrx_oper *mrx_convolution_create_reach(const mrx_mra *mra, int n_terms, const double *coef, const double *expo, double prec, int root,
                                       int reach) {
    GaussExp<1> kernel(n_terms);
    for (int i = 0; i < n_terms; i++) {
        kernel[i].coef = coef[i];
        kernel[i].alpha = expo[i];
    }
    auto *o = new mrx_oper;
    o->op = build_convolution_operator(mra->m, kernel, prec / 100.0, prec, root, reach);
    return o;
}
mrx_oper *mrx_abgv_create(const mrx_mra *mra, double a, double b) {
    auto *o = new mrx_oper;
    o->op = build_abgv_operator(mra->m, a, b);
    return o;
}
mrx_oper *mrx_ph_create(const mrx_mra *mra, int order) {
    auto *o = new mrx_oper;
    o->op = build_ph_operator(mra->m, order);
    return o;
}
mrx_oper *mrx_bs_create(const mrx_mra *mra, int order) {
    auto *o = new mrx_oper;
    o->op = build_bs_operator(mra->m, order);
    return o;
}
mrx_oper *mrx_oper_from_arrays(const mrx_mra *mra, int n_terms, const int *n_depth, const int *max_transl, const double *mats,
                               const double *norms, int oper_root, int derivative_order, double build_prec) {
    auto *o = new mrx_oper;
    Operator &op = o->op;
    op.k = mra->m.order;
    op.K = op.k + 1;
    op.operRoot = oper_root;
    op.derivative = derivative_order > 0;
    op.order = derivative_order;
    op.buildPrec = build_prec;
    op.terms.resize(n_terms);
    size_t mt_pos = 0, node_pos = 0;
    const size_t stride = (size_t)4 * op.K * op.K;
    for (int t = 0; t < n_terms; t++) {
        OperTerm &term = op.terms[t];
        term.nDepth = n_depth[t];
        term.matStride = stride;
        term.maxTransl.assign(max_transl + mt_pos, max_transl + mt_pos + n_depth[t]);
        mt_pos += n_depth[t];
        term.offset.resize(n_depth[t]);
        size_t total = 0;
        for (int d = 0; d < n_depth[t]; d++) {
            term.offset[d] = total;
            total += 2 * (size_t)term.maxTransl[d] + 1;
        }
        term.mats.assign(mats + node_pos * stride, mats + (node_pos + total) * stride);
        term.norms.assign(norms + node_pos * 4, norms + (node_pos + total) * 4);
        node_pos += total;
    }
    return o;
}
void mrx_oper_destroy(mrx_oper *oper) { delete oper; }
int mrx_oper_n_terms(const mrx_oper *oper) { return oper->op.size(); }
int mrx_oper_band_widths(mrx_oper *oper, double prec, int *band_max, int band_max_len) {
    oper->op.calcBandWidths(prec);
    int n = (int)oper->op.bandMax.size();
    if (band_max)
        for (int i = 0; i < n && i < band_max_len; i++) band_max[i] = oper->op.bandMax[i];
    oper->op.clearBandWidths();
    return n;
}
int mrx_oper_depth(const mrx_oper *oper, int term) { return oper->op.terms[term].nDepth; }
int mrx_oper_max_transl(const mrx_oper *oper, int term, int depth) { return oper->op.terms[term].maxTransl[depth]; }
int mrx_oper_node(const mrx_oper *oper, int term, int depth, int transl, double *mats, double *norms) {
    const OperTerm &t = oper->op.terms[term];
    if (depth < 0 || depth >= t.nDepth || std::abs(transl) > t.maxTransl[depth]) return 1;
    if (mats) std::memcpy(mats, t.node(depth, transl), sizeof(double) * t.matStride);
    if (norms) std::memcpy(norms, t.nodeNorms(depth, transl), sizeof(double) * 4);
    return 0;
}
int mrx_quadrature(int n, double *roots, double *weights) {
    const Quadrature &q = quadrature(n);
    for (int i = 0; i < n; i++) {
        if (roots) roots[i] = q.roots[i];
        if (weights) weights[i] = q.weights[i];
    }
    return n;
}
double mrx_interp_scaling(int k, int j, double x, int derivative) {
    return derivative ? interp_scaling_deriv(k, j, x) : interp_scaling_eval(k, j, x);
}
int mrx_poisson_kernel(double epsilon, double r_min, double r_max, double *coef, double *expo, int cap) {
    GaussExp<1> k = poisson_kernel(epsilon, r_min, r_max);
    for (int i = 0; i < (int)k.size() && i < cap; i++) {
        if (coef) coef[i] = k[i].coef;
        if (expo) expo[i] = k[i].alpha;
    }
    return (int)k.size();
}
int mrx_helmholtz_kernel(double mu, double epsilon, double r_min, double r_max, double *coef, double *expo, int cap) {
    GaussExp<1> k = helmholtz_kernel(mu, epsilon, r_min, r_max);
    for (int i = 0; i < (int)k.size() && i < cap; i++) {
        if (coef) coef[i] = k[i].coef;
        if (expo) expo[i] = k[i].alpha;
    }
    return (int)k.size();
}

// ---- hot path
int mrx_apply(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec, mrx_apply_stats *stats) {
    require_device("mrx_apply");
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (oper->op.derivative) MRX_ABORT("mrx_apply: derivative operator passed to the convolution apply");
    device_apply(prec, *out, *oper, *inp, max_iter, abs_prec != 0, stats);
    return 0;
}
int mrx_apply_unit_cell(int inside, double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec,
                        mrx_apply_stats *stats) {
    require_device("mrx_apply_unit_cell");
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (oper->op.derivative) MRX_ABORT("mrx_apply_unit_cell: derivative operator passed to the convolution apply");
    device_apply(prec, *out, *oper, *inp, max_iter, abs_prec != 0, stats, nullptr, nullptr, inside ? 1 : 2);
    return 0;
}
int mrx_apply_prec_trees(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int n_prec, mrx_tree *const *prec_trees, int max_iter,
                         int abs_prec, const mrx_comm *comm, mrx_apply_stats *stats) {
    require_device("mrx_apply_prec_trees");
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (oper->op.derivative) MRX_ABORT("mrx_apply_prec_trees: derivative operator passed to the convolution apply");
    std::vector<mrx_tree *> pt(prec_trees, prec_trees + (n_prec > 0 ? n_prec : 0));
    device_apply(prec, *out, *oper, *inp, max_iter, abs_prec != 0, stats, comm, &pt);
    return 0;
}
double mrx_bench_mw_transform(mrx_tree *tree, int type, int reps, int *branch_nodes) {
    require_device("mrx_bench_mw_transform");
    double ms = 0.0;
    // type 2: TopDown(+=), the mode mrcpp::apply closes with (apply.cpp:82); it accumulates, the tree is scratch afterwards
    device_mw_transform(*tree, type == 2 ? MRX_TOP_DOWN : type, type != 2, true, reps > 0 ? reps : 1, &ms, branch_nodes);
    return ms;
}
int mrx_apply_sharded(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec, const mrx_comm *comm,
                      mrx_apply_stats *stats) {
    require_device("mrx_apply_sharded");
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (oper->op.derivative) MRX_ABORT("mrx_apply_sharded: derivative operator passed to the convolution apply");
    device_apply(prec, *out, *oper, *inp, max_iter, abs_prec != 0, stats, comm);
    return 0;
}
int mrx_apply_derivative(mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int dir, mrx_apply_stats *stats) {
    require_device("mrx_apply_derivative");
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (dir < 0 || dir >= 3) MRX_ABORT("Invalid apply dir");
    if (!oper->op.derivative) MRX_ABORT("mrx_apply_derivative: not a derivative operator");
    device_apply_derivative(*out, *oper, *inp, dir, stats);
    return 0;
}
int mrx_mw_transform(mrx_tree *tree, int type, int overwrite) {
    require_device("mrx_mw_transform");
    if (type == MRX_BOTTOM_UP && !overwrite) MRX_ABORT("BottomUp without overwrite is not implemented (MWTree.cpp:148)");
    if (type != MRX_BOTTOM_UP && type != MRX_TOP_DOWN) MRX_ABORT("Invalid wavelet transform");
    device_mw_transform(*tree, type, overwrite != 0);
    return 0;
}
int mrx_node_mw_transform(mrx_tree *tree, int kind, int n_nodes, const int *slots) {
    require_device("mrx_node_mw_transform");
    if (kind != MRX_COMPRESSION && kind != MRX_RECONSTRUCTION) MRX_ABORT("Invalid operation");
    device_node_transform(*tree, 0, kind == MRX_COMPRESSION ? 0 : 1, n_nodes, slots);
    return 0;
}
int mrx_node_cv_transform(mrx_tree *tree, int kind, int n_nodes, const int *slots) {
    require_device("mrx_node_cv_transform");
    if (kind != MRX_FORWARD && kind != MRX_BACKWARD) MRX_ABORT("Invalid operation");
    device_node_transform(*tree, 1, kind == MRX_FORWARD ? 0 : 1, n_nodes, slots);
    return 0;
}
double mrx_bench_cv_transform(mrx_tree *tree, int reps) {
    require_device("mrx_bench_cv_transform");
    double ms = 0.0;
    device_node_transform(*tree, 1, 0, -1, nullptr, reps > 0 ? reps : 1, &ms);
    return ms;
}
double mrx_calc_square_norm(mrx_tree *tree) {
    require_device("mrx_calc_square_norm");
    device_calc_norms_all(*tree);
    tree->host.calcSquareNorm();
    return tree->host.squareNorm;
}
double mrx_dot(mrx_tree *bra, mrx_tree *ket) {
    require_device("mrx_dot");
    if (!(bra->host.mra == ket->host.mra)) MRX_ABORT("Incompatible MRA");
    return device_dot(*bra, *ket);
}
int mrx_tree_rescale(mrx_tree *tree, double c) {
    require_device("mrx_tree_rescale");
    device_rescale(*tree, c);
    return 0;
}
namespace {
// the tree is a bare grid from here on: no coefficients, no norms, nothing on the device
void mark_grid_only(mrx_tree *tree) {
    Tree<3> &h = tree->host;
    for (int n = 0; n < h.nReal; n++) {
        h.nodes[n].flags &= ~FlagHasCoefs;
        for (int t = 0; t < 8; t++) h.cnorm[(size_t)n * 8 + t] = -1.0;
        h.sqn[n] = -1.0;
    }
    h.squareNorm = -1.0;
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.nNodes = 0;
    tree->dev.nGen = 0;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
}
} // namespace

// clear_grid(out) (src/treebuilders/grid.cpp:180-186): keep the grid, drop coefficients and norms
int mrx_tree_clear_grid(mrx_tree *tree) {
    tree->host.deleteGenerated();
    mark_grid_only(tree);
    return 0;
}
int mrx_tree_build_grid_from(mrx_tree *out, const mrx_tree *inp) {
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    out->host.extendGridFrom(inp->host);
    mark_grid_only(out);
    return 0;
}
int mrx_tree_add_adaptive(double prec, mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp, int max_iter, int abs_prec) {
    require_device("mrx_tree_add");
    if (n <= 0) MRX_ABORT("mrx_tree_add: empty input vector");
    for (int i = 0; i < n; i++) {
        if (!(out->host.mra == inp[i]->host.mra)) MRX_ABORT("Incompatible MRA");
        if (inp[i] == out) MRX_ABORT("mrx_tree_add: output tree among the inputs");
    }
    device_add(*out, n, coefs, inp, prec, max_iter, abs_prec != 0);
    return 0;
}
int mrx_tree_refine_grid(mrx_tree *tree, double prec, int abs_prec, int scales) {
    return device_refine_grid(*tree, prec, abs_prec != 0, scales);
}
int mrx_tree_add_inplace(mrx_tree *tree, double c, mrx_tree *inp) {
    require_device("mrx_tree_add_inplace");
    if (!(tree->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (tree == inp) MRX_ABORT("mrx_tree_add_inplace: a tree cannot be added to itself in place (use rescale)");
    device_add_inplace(*tree, c, *inp);
    return 0;
}
int mrx_tree_multiply(double prec, mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp, int max_iter, int abs_prec,
                      int use_max_norms) {
    require_device("mrx_tree_multiply");
    if (n <= 0) MRX_ABORT("mrx_tree_multiply: empty input vector");
    for (int i = 0; i < n; i++) {
        if (!(out->host.mra == inp[i]->host.mra)) MRX_ABORT("Incompatible MRA");
        if (inp[i] == out) MRX_ABORT("mrx_tree_multiply: output tree among the inputs");
    }
    device_multiply(*out, n, coefs, inp, prec, max_iter, abs_prec != 0, use_max_norms != 0);
    return 0;
}
int mrx_tree_power(double prec, mrx_tree *out, mrx_tree *inp, double p, int max_iter, int abs_prec) {
    require_device("mrx_tree_power");
    if (!(out->host.mra == inp->host.mra)) MRX_ABORT("Incompatible MRA");
    if (inp == out) MRX_ABORT("mrx_tree_power: output tree is the input");
    const double one = 1.0;
    mrx_tree *v[1] = {inp};
    device_multiply(*out, 1, &one, v, prec, max_iter, abs_prec != 0, false, &p);
    return 0;
}
int mrx_tree_add(mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp) {
    return mrx_tree_add_adaptive(-1.0, out, n, coefs, inp, 0, 0);
}
int mrx_tree_sync_device(mrx_tree *tree) {
    require_device("mrx_tree_sync_device");
    if (!tree->devValid) tree_upload(*tree);
    return 0;
}
int mrx_tree_sync_host(mrx_tree *tree) {
    if (tree->hostCoefsValid) return 0;
    require_device("mrx_tree_sync_host");
    tree_download(*tree);
    return 0;
}
int mrx_tree_set_host_mirror(mrx_tree *tree, int on) {
    tree->hostMirror = on != 0;
    return 0;
}
int mrx_tree_set_shared_host_mirror(mrx_tree *tree, mrx_comm *comm) {
    tree->hostMirror = true;
    if (!comm_has_host_arena(comm)) return 1;
    if (tree->mirrorComm == comm) return 0;
    if (!tree->hostCoefsValid) mrx_tree_sync_host(tree);
    tree->host.rebaseChunks(host_arena_alloc, host_arena_free);
    tree->mirrorComm = comm;
    return 0;
}
int mrx_tree_drop_device(mrx_tree *tree) {
    if (!tree->hostCoefsValid) mrx_tree_sync_host(tree);
    tree_drop_device(*tree);
    return 0;
}

static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
void mrx_timer_start(void) {
    require_device("mrx_timer_start");
    if (!g_ev0) {
        cudaEventCreate(&g_ev0);
        cudaEventCreate(&g_ev1);
    }
    cudaEventRecord(g_ev0, g_stream);
}
double mrx_timer_stop_ms(void) {
    require_device("mrx_timer_stop_ms");
    cudaEventRecord(g_ev1, g_stream);
    cudaEventSynchronize(g_ev1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_ev0, g_ev1);
    return ms;
}
void *mrx_tree_host_handle(mrx_tree *tree) { return &tree->host; }
void *mrx_oper_host_handle(mrx_oper *oper) { return &oper->op; }
void mrx_tree_host_modified(mrx_tree *tree) {
    tree->hostCoefsValid = true;
    tree->devValid = false;
    tree->dev.nNodes = 0;
    tree->dev.topoNodes = -1;
    tree->dev.partial = false;
}
}
