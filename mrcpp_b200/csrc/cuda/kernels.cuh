// Launchers of the tree-maintenance kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace mrx {

const double *device_filters(int k);

/// component norms: norms[node*8 + c] = ||block c|| for node = slots ? slots[i] : i, i < n; normsW (optional): the same
/// values in list order, normsW[i*8 + c]
void launch_norms(const double *coefs, double *norms, const int *slots, int n, int Kd, cudaStream_t st, double *normsW = nullptr);

/// one level of the two-scale transform. pairs = (parent slot, child0 slot) x cnt.
/// down: children.scaling (=|+=) reconstruct(parent 8 blocks); up: parent 8 blocks = compress(children.scaling)
/// norms (optional, honoured when transform_fuses_norms(K)): component norms of the blocks written, in the node-store
/// layout norms[node*8 + c]: scaling norm of every child (down) / all 8 norms of every parent (up)
void launch_transform(bool down, bool overwrite, double *coefs, const int *pairs, int cnt, int K, const double *filters,
                      cudaStream_t st, double *norms = nullptr);
bool transform_fuses_norms(int K);

/// in-node compression MWNode::mwTransform(Compression) of the nodes pairs[2 i] (pairs[2 i + 1] unused)
void launch_compress_nodes(double *coefs, const int *pairs, int cnt, int K, const double *filters, cudaStream_t st,
                           double *norms = nullptr);

/// in-node reconstruction MWNode::mwTransform(Reconstruction) of the nodes pairs[2 i] (pairs[2 i + 1] unused)
void launch_reconstruct_nodes(double *coefs, const int *pairs, int cnt, int K, const double *filters, cudaStream_t st);
/// MWNode::cvTransform(Forward | Backward) of the nodes items[2 i] (items[2 i + 1] = the node's scale); map = sqrt(1 / w) or
/// sqrt(w) per quadrature index
void launch_cv_transform(double *coefs, const int *items, int cnt, int K, const double *map, bool backward, cudaStream_t st);

/// ProjectionCalculator::calcNode for a Gaussian expansion (project.cu): function values at the expanded child quadrature
/// points of every work node, scaled to scaling coefficients (cvTransform Backward); nodeInfo = (scale, lx, ly, lz)
struct GaussTable {
    const double *coef, *alpha, *pos; // [n], [n], [n][3]
    const int *power;                 // [n][3]
    int n;
};
void launch_project_eval(double *coefs, const int *slots, const int4 *nodeInfo, int cnt, int K, const GaussTable &g, const double *roots,
                         const double *sqrtw, cudaStream_t st);

/// generated children of input-tree nodes (FunctionNode::genChildren + giveChildrenCoefs):
/// items = (parent slot, child0 slot) in the unified slot space (slot >= nReal -> generated pool)
void launch_gen_children(const double *realCoefs, double *genCoefs, double *genNorms, int nReal, const int *items, int cnt,
                         int K, const double *filters, cudaStream_t st);

/// fixed-order sum of the per-chunk partial results of nodes that were split across CTAs:
/// items = (slot, firstPartial, nPartials) x cnt
void launch_reduce_partials(double *coefs, const double *partials, const int *items, int cnt, int ncoef, cudaStream_t st);

void launch_dot(const double *a, const double *b, const int *pairs, double *res, int np, int nRoots, int Kd, cudaStream_t st);
void launch_scale(double *x, size_t n, double c, cudaStream_t st);
/// out node pairs[2 p] += c * in node pairs[2 p + 1] (wavelet blocks; scaling block only where the out node is a root)
void launch_axpy_nodes(double *out, const double *in, const int *pairs, int np, int nRoots, int Kd, double c, cudaStream_t st);

/// element-wise part of the multiplication on scratch nodes (chunk of nC nodes; children of node j in slots nC + 8 j + t):
/// mode 0: P = c * forward(S), mode 1: P *= c * forward(S), mode 2: P = backward(P), mode 3: P = forward(S) ^ c; map = sqrt(1 / w) or
/// sqrt(w) per index
void launch_product_values(double *P, const double *S, const int *scale, int nC, int K, const double *map, double c, int mode,
                           cudaStream_t st);

} // namespace mrx
