/* C-ABI of the B200-native MRCPP operator-application path (libmrcpp_b200.so).
 *
 * Plain pointers and sizes only. Every entry point names the reference interface it replaces
 * (file:line relative to the MRCPP source tree). Error convention follows the reference: hard
 * failures print and abort() (src/utils/Printer.h:165-169); calls that can fail softly return a
 * non-zero status. All coefficient arrays are FP64; all index arrays int32.
 *
 * Node layout (identical to the reference, src/utils/math_utils.cpp:223-235, MWNode.h): one node =
 * 8 blocks of (k+1)^3 doubles; block t bit d = wavelet (compressed form) along dimension d; inside a
 * block the x index is fastest.
 *
 * Node order in every array interface: slot order of the flat tree = roots in box order (x fastest),
 * then the 8 children of each split node contiguously, in creation order (the reference's
 * NodeAllocator serial index order, src/trees/NodeAllocator.cpp:115-216).
 */
#ifndef MRCPP_B200_H
#define MRCPP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrx_mra mrx_mra;   /* MultiResolutionAnalysis<3>  (src/trees/MultiResolutionAnalysis.h:49) */
typedef struct mrx_tree mrx_tree; /* FunctionTree<3,double>      (src/trees/FunctionTree.h)               */
typedef struct mrx_oper mrx_oper; /* ConvolutionOperator<3> / DerivativeOperator<3> packed for HBM       */
typedef struct mrx_comm mrx_comm; /* one rank of a multi-GPU job (NCCL communicator over NVLink)          */

enum { MRX_TOP_DOWN = 0, MRX_BOTTOM_UP = 1 }; /* api/constants.h TopDown/BottomUp */
enum { MRX_FORWARD = 0, MRX_BACKWARD = 1 };          /* api/constants.h:48 CV_Transform */
enum { MRX_COMPRESSION = 0, MRX_RECONSTRUCTION = 1 }; /* api/constants.h:49 MW_Transform */

/* counters of one apply; mirrors OperatorStatistics (src/operators/OperatorStatistics.cpp:83-106) */
typedef struct mrx_apply_stats {
    long long g_nodes;     /* totGCount: calcNode invocations over all refinement iterations          */
    long long f_applied;   /* totFCount: (g,f,ft,gt,term) tuples that passed screening; x 6(k+1)^4 flop */
    long long gen_nodes;   /* generated input nodes materialised on the device                          */
    int iterations;        /* refinement iterations of TreeBuilder::build                               */
    int n_nodes_out;       /* nodes of the output tree                                                  */
    double ms_upload;      /* host->device copies of the input tree + tables                            */
    double ms_build;       /* TreeBuilder::build loop (device kernels + host split logic)               */
    double ms_kernel;      /* device time of the apply kernels of all iterations (CUDA events)          */
    double ms_contract;    /* device time of the contraction kernel alone (the FP64 tensor-pipe kernel) */
    double ms_post;        /* TopDown(+=) and BottomUp transforms + norms                               */
    double ms_download;    /* device->host copy of the result                                           */
    long long kernel_launches; /* CUDA kernels launched by this call                                    */
    long long f_applied_rank;  /* sharded apply: tuples contracted by THIS rank (f_applied = sum over ranks) */
    long long h2d_bytes;       /* input-tree bytes this call moved host -> device (0 if the input was resident)     */
    long long d2h_bytes;       /* result bytes this call left in host memory (output tree with a host mirror), else 0 */
} mrx_apply_stats;

/* ---- library / device ------------------------------------------------------------------------ */
/* Loads the filter tables (share/mwfilters subset packed by tools/pack_tables.py; replaces
 * details::find_filters, src/utils/details.cpp:53-69) and selects the CUDA device. device < 0 keeps the
 * library host-only (construction helpers work; every hot-path call aborts: there is no CPU fallback). */
int mrx_init(const char *table_path, int device);
int mrx_device_count(void);
const char *mrx_version(void);

/* ---- MRA -------------------------------------------------------------------------------------- */
/* BoundingBox<3>(scale, corner, boxes) + InterpolatingBasis(order) + MultiResolutionAnalysis<3>(world,
 * basis, max_depth): src/trees/MultiResolutionAnalysis.cpp:69-77, examples/poisson.cpp:24-31 */
mrx_mra *mrx_mra_create(int order, int root_scale, const int corner[3], const int nboxes[3], int max_depth);
/* BoundingBox(n, l, nb, sf, pbc = true) (src/trees/BoundingBox.cpp:95-117): periodic world. The reference's periodic index
 * arithmetic (src/utils/periodic_utils.cpp:35-85) works on the unit cell [-1, 1]^3 in box units, so the world must have root
 * scale 0, corner (-1,-1,-1) and 2 x 2 x 2 root boxes (aborts otherwise); scaling factors are not supported (length unit = box). */
int mrx_mra_set_periodic(mrx_mra *mra, int periodic);
void mrx_mra_destroy(mrx_mra *mra);

/* ---- function trees --------------------------------------------------------------------------- */
mrx_tree *mrx_tree_create(const mrx_mra *mra);                  /* FunctionTree<3>(MRA): empty roots        */
void mrx_tree_destroy(mrx_tree *tree);
int mrx_tree_n_nodes(const mrx_tree *tree);                     /* MWTree::getNNodes                        */
int mrx_tree_n_end_nodes(const mrx_tree *tree);                 /* MWTree::getNEndNodes                     */
double mrx_tree_square_norm(const mrx_tree *tree);              /* MWTree::getSquareNorm                    */
void mrx_tree_clear(mrx_tree *tree);                            /* FunctionTree::clear: back to empty roots */

/* Import a tree built by the reference (or anybody): arrays in slot order. child0[i] = slot of child 0
 * or -1. Replaces nothing in the reference; it is what a binding on the MRCPP side calls with the
 * contents of its NodeAllocator chunks (src/trees/NodeAllocator.cpp:362-415 reassemble()). */
mrx_tree *mrx_tree_from_arrays(const mrx_mra *mra, int n_nodes, const int *scale, const int *transl /*[n][3]*/,
                               const int *parent, const int *child0, const double *coefs /*[n][8*(k+1)^3]*/);
/* Export: any pointer may be NULL. norms = component norms [n][8] (MWNode::getComponentNorm). Makes the
 * host copy current first (device->host copy if the device copy is newer). */
int mrx_tree_to_arrays(mrx_tree *tree, int *scale, int *transl, int *parent, int *child0, double *coefs, double *norms);
/* copy_grid (src/treebuilders/grid.cpp:150-166): give `out` the node structure of `inp`, no coefs */
int mrx_tree_copy_grid(mrx_tree *out, const mrx_tree *inp);

/* FunctionTree::integrate (src/trees/FunctionTree.cpp:438-454, FunctionNode.cpp:128-157): integral of the function over
 * the world, from the scaling blocks of the root nodes (read back from HBM if the host copy is not current) */
double mrx_tree_integrate(mrx_tree *tree);
/* FunctionTree::evalf (precise == 0) / evalf_precise (precise != 0) at n_points points r[n][3] (src/trees/FunctionTree.cpp:
 * 374-436): function values from the downloaded tree (host arithmetic; the tree is brought to the host first if its
 * current copy is in HBM). Zero outside the world. */
int mrx_tree_evalf(mrx_tree *tree, int n_points, const double *r /*[n][3]*/, double *values /*[n]*/, int precise);
/* FunctionTree::saveTreeTXT / loadTreeTXT (src/trees/FunctionTree.cpp:240-372): the reference's text interchange format (function
 * values at the quadrature points of the children of every end node, MADNESS conventions). Host arithmetic. load replaces the
 * content of `tree` (same MRA) by the tree of the file; files with complete sibling groups, as saveTreeTXT writes them. */
int mrx_tree_save_txt(mrx_tree *tree, const char *path);
int mrx_tree_load_txt(mrx_tree *tree, const char *path);
/* build_grid(out, GaussExp) alone (src/treebuilders/grid.cpp:78-123): host only, leaves a grid without coefficients;
 * max_iter < 0: no bound */
int mrx_build_grid_gaussians(mrx_tree *tree, int n_gauss, const double *coef, const double *alpha,
                             const double *pos /*[n][3]*/, const int *power /*[n][3] or NULL*/, int max_iter);

/* Host-side input generator: build_grid + project of a Gaussian expansion
 * (src/treebuilders/grid.cpp:78-123, project.cpp:85-104, ProjectionCalculator.cpp:34-51). The per-node
 * quadrature runs on the host; with finalize != 0 the closing mwTransform(BottomUp) + calcSquareNorm
 * (project.cpp:96-97) run on the device. finalize == 0 stops before them (no device needed): the
 * caller owes the tree an mrx_mw_transform(BOTTOM_UP) + mrx_calc_square_norm. */
int mrx_project_gaussians(mrx_tree *tree, double prec, int n_gauss, const double *coef, const double *alpha,
                          const double *pos /*[n][3]*/, const int *power /*[n][3] or NULL*/, int build_grid, int finalize);

/* project(prec, out, f) of an arbitrary function given as a callback (src/treebuilders/project.cpp:85-104,
 * ProjectionCalculator.cpp:34-51): host quadrature, refinement by the wavelet norm from the tree's current grid.
 * threads_ok != 0: the callback may be called from several OpenMP threads at once. finalize as above. */
typedef double (*mrx_func3)(const double r[3], void *user);
int mrx_project_function(mrx_tree *tree, double prec, mrx_func3 f, void *user, int threads_ok, int finalize);

/* The same build_grid + project with the per-node quadrature on the device:
 * ProjectionCalculator::calcNode (src/treebuilders/ProjectionCalculator.cpp:34-51: function values at the expanded child
 * quadrature points, MWNode::cvTransform(Backward), MWNode::mwTransform(Compression), MWNode::calcNorms) runs as CUDA kernels
 * for every work vector of TreeBuilder::build (TreeBuilder.cpp:38-86); the host keeps the topology and takes the split
 * decisions (WaveletAdaptor.h:51-54). The tree is born resident in HBM and has no host coefficient storage until it is
 * downloaded (mrx_tree_to_arrays / mrx_tree_sync_host). */
int mrx_project_gaussians_device(mrx_tree *tree, double prec, int n_gauss, const double *coef, const double *alpha,
                                 const double *pos /*[n][3]*/, const int *power /*[n][3] or NULL*/, int build_grid);

/* ---- operators -------------------------------------------------------------------------------- */
/* PoissonOperator(MRA, prec): src/operators/PoissonOperator.cpp:40-55 */
mrx_oper *mrx_poisson_create(const mrx_mra *mra, double prec);
/* HelmholtzOperator(MRA, mu, prec): src/operators/HelmholtzOperator.cpp:44-59 */
mrx_oper *mrx_helmholtz_create(const mrx_mra *mra, double mu, double prec);
/* The Poisson / Helmholtz constructors keep the host tables of the last 16 operators built, keyed by every parameter they depend
 * on (an SCF loop rebuilds the same operators every iteration, examples/scf.cpp:102); a repeated construction
 * copies the tables (1-2 ms) instead of rebuilding them (30-200 ms). Counters of this process: */
void mrx_oper_cache_stats(long long *hits, long long *misses);
/* ConvolutionOperator<3>(MRA, GaussExp<1> kernel, prec): src/operators/ConvolutionOperator.cpp:50-62 */
mrx_oper *mrx_convolution_create(const mrx_mra *mra, int n_terms, const double *coef, const double *expo, double prec);
/* PoissonOperator(MRA, prec, root, reach) src/operators/PoissonOperator.cpp:56-77, HelmholtzOperator(MRA, mu, prec, root, reach)
 * src/operators/HelmholtzOperator.cpp:60-81, ConvolutionOperator<3>(MRA, kernel, prec, root, reach)
 * src/operators/ConvolutionOperator.cpp:63-76: operators for periodic worlds -- operator trees with reach + 1 root boxes
 * (MWOperator::getOperatorMRA, MWOperator.cpp:110-129), kernel precision prec / 100, r_max stretched over the reach.
 * The apply supports root == the world's root scale (0). */
mrx_oper *mrx_poisson_create_reach(const mrx_mra *mra, double prec, int root, int reach);
mrx_oper *mrx_helmholtz_create_reach(const mrx_mra *mra, double mu, double prec, int root, int reach);
mrx_oper *mrx_convolution_create_reach(const mrx_mra *mra, int n_terms, const double *coef, const double *expo, double prec, int root,
                                       int reach);
/* ABGVOperator<3>(MRA, a, b): src/operators/ABGVOperator.cpp:46-74 */
mrx_oper *mrx_abgv_create(const mrx_mra *mra, double a, double b);
/* PHOperator<3>(MRA, order), order 1 or 2: src/operators/PHOperator.cpp:40-69 (Holoborodko smoothing derivative) */
mrx_oper *mrx_ph_create(const mrx_mra *mra, int order);
/* BSOperator<3>(MRA, order), order 1, 2 or 3: src/operators/BSOperator.cpp:40-66 (B-spline derivative) */
mrx_oper *mrx_bs_create(const mrx_mra *mra, int order);
/* Import operator trees built by the reference: per term the [depth][transl] node cache of
 * OperatorTree::setupOperNodeCache (src/trees/OperatorTree.cpp:200-238). max_transl[t][d] for d <
 * n_depth[t]; mats/norms hold, term after term, depth after depth, transl = -max..max, the node's four
 * (k+1)^2 blocks (column-major, element [i + (k+1) m]) and four component norms. */
mrx_oper *mrx_oper_from_arrays(const mrx_mra *mra, int n_terms, const int *n_depth, const int *max_transl /* ragged, concatenated */,
                               const double *mats, const double *norms, int oper_root, int derivative_order,
                               double build_prec);
void mrx_oper_destroy(mrx_oper *oper);
int mrx_oper_n_terms(const mrx_oper *oper);                      /* MWOperator::size                        */
/* MWOperator::calcBandWidths(prec) + getMaxBandWidth(depth): src/operators/MWOperator.cpp:63-108.
 * widths: [n_terms][max_depth+1][5] (T,C,B,A,max), may be NULL; returns number of depths. */
int mrx_oper_band_widths(mrx_oper *oper, double prec, int *band_max, int band_max_len);
/* packed table introspection (tests): pointer to the 4 blocks / 4 norms of node (term, depth, transl) */
int mrx_oper_node(const mrx_oper *oper, int term, int depth, int transl, double *mats /*4*(k+1)^2*/, double *norms /*4*/);
int mrx_oper_depth(const mrx_oper *oper, int term);
int mrx_oper_max_transl(const mrx_oper *oper, int term, int depth);
/* table introspection (tests): Gauss-Legendre roots / weights on [0, 1] (GaussQuadrature.cpp:157-193 through
 * QuadratureCache.cpp:41-45) and the interpolating scaling functions phi_j of order k or their derivatives
 * (InterpolatingBasis.cpp:62-84), which feed the projection and the ABGV operator construction */
int mrx_quadrature(int n, double *roots, double *weights);
double mrx_interp_scaling(int k, int j, double x, int derivative);
/* PoissonKernel / HelmholtzKernel expansions (tests: size()==26 / 33 KATs) */
int mrx_poisson_kernel(double epsilon, double r_min, double r_max, double *coef, double *expo, int cap);
int mrx_helmholtz_kernel(double mu, double epsilon, double r_min, double r_max, double *coef, double *expo, int cap);

/* ---- the hot path (device) -------------------------------------------------------------------- */
/* mrcpp::apply(prec, out, oper, inp, maxIter, absPrec): src/treebuilders/apply.cpp:68-93.
 * `out` enters with its starting grid (normally empty roots) and no coefficients. The result stays
 * resident in HBM; mrx_tree_to_arrays / mrx_tree_sync_host bring it back. */
int mrx_apply(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec, mrx_apply_stats *stats);
/* ---- multi-GPU: one process per GPU ----------------------------------------------
 * The reference distributes whole trees over MPI ranks (src/utils/parallel.cpp) and runs each apply inside
 * one rank with OpenMP (TreeCalculator.h:39-50). Here ONE apply is sharded: every refinement iteration's
 * work vector is dealt out cyclically (mrx_shard_cyclic); input tree and operator are replicated. Component
 * norms (TreeBuilder's norm bookkeeping and the split decisions, which every rank then takes identically) are
 * all-gathered with NCCL; output coefficient blocks are pushed into the peers' HBM over NVLink by the copy
 * engines (CUDA IPC mapping of a staging buffer; ncclAllGather if IPC is unavailable) while the next iteration
 * runs, so every rank ends with the complete output tree, bit-identical across ranks and to the single-GPU
 * result. Rank 0 creates the id, the host framework broadcasts its 128 bytes (torch.distributed / MPI_Bcast),
 * every rank calls mrx_comm_create. */
int mrx_comm_unique_id(char *id128);
mrx_comm *mrx_comm_create(int rank, int world, const char *id128);
/* Collective, ranks of ONE node: a host arena of `bytes` that every rank maps and registers with CUDA (one anonymous
 * shared-memory file), so that every GPU can write result nodes into the same host pages over its own PCIe link
 * (mrx_tree_set_shared_host_mirror). Returns 0 when all ranks have it, 1 when shared mapping is not possible here (the
 * host mirror of a sharded apply then stays on the calling rank's own link). */
int mrx_comm_host_arena(mrx_comm *comm, long long bytes);
void mrx_comm_destroy(mrx_comm *comm);
int mrx_comm_rank(const mrx_comm *comm);
int mrx_comm_size(const mrx_comm *comm);
void mrx_shard_partition(const long long *cost, int n, int world, int *begin /*[world+1]*/);
/* the distribution mrx_apply_sharded uses: block-cyclic with blocks of B = mrx_shard_block() consecutive items: item i of an
 * iteration's work vector is computed by rank (i / B) % world; exchange buffers are rank-major with equal, padded segments:
 * item i = row ((i / B) % world) * rows + (i / (B world)) * B + i % B, rows = ceil(ceil(n / B) / world) * B */
void mrx_shard_cyclic(int n, int world, int rank, int *count, int *rows);
int mrx_shard_cyclic_row(int i, int n, int world);
int mrx_shard_block(void);
/* mrcpp::apply sharded over the ranks of `comm` (comm == NULL: same as mrx_apply). Collective: every rank
 * calls it with identical arguments. stats->f_applied / gen_nodes are summed over ranks. */
int mrx_apply_sharded(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec,
                      const mrx_comm *comm, mrx_apply_stats *stats);
/* mrcpp::apply_near_field (inside = 1) / apply_far_field (inside = 0) = apply_on_unit_cell(inside, ...):
 * src/treebuilders/apply.cpp:161-188, :294-342, ConvolutionCalculator::fillOperBand src/treebuilders/ConvolutionCalculator.cpp:191-218:
 * on a periodic world, only the input nodes whose (unwrapped) index lies inside / outside the unit cell contribute. The plain
 * mrx_apply on a periodic world is the sum of the two (band clipped to the operator's reach, ConvolutionCalculator.cpp:166-172). */
int mrx_apply_unit_cell(int inside, double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec,
                        mrx_apply_stats *stats);
/* mrcpp::apply(prec, out, oper, inp, precTrees, maxIter, absPrec): src/treebuilders/apply.cpp:214-251. The precision is scaled
 * per output node by 1 / max_i sqrt(maxSquareNorm of prec_trees[i] at the node's index) (makeMaxSquareNorms, MWTree.cpp:536-543;
 * where a precision tree is coarser than the output grid the generated node's own scaled norm, MWNode.h:84), in the screening
 * threshold (ConvolutionCalculator.cpp:241-248) and in the split threshold (WaveletAdaptor.h:51-54). n_prec = 0: factor 1.
 * comm == NULL: one GPU; otherwise sharded like mrx_apply_sharded (collective). */
int mrx_apply_prec_trees(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int n_prec, mrx_tree *const *prec_trees,
                         int max_iter, int abs_prec, const mrx_comm *comm, mrx_apply_stats *stats);
/* mrcpp::apply(out, DerivativeOperator, inp, dir): src/treebuilders/apply.cpp:379-412 */
int mrx_apply_derivative(mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int dir, mrx_apply_stats *stats);
/* MWTree::mwTransform(type, overwrite): src/trees/MWTree.cpp:143-216 (+ norms of touched nodes) */
int mrx_mw_transform(mrx_tree *tree, int type, int overwrite);
/* MWNode::mwTransform(kind): src/trees/MWNode.cpp:557-594 -- in-node compression / reconstruction of the listed nodes (slots in
 * node-store order; n_nodes < 0: every node), in place on the resident node store; component norms are refreshed. */
int mrx_node_mw_transform(mrx_tree *tree, int kind /* MRX_COMPRESSION | MRX_RECONSTRUCTION */, int n_nodes, const int *slots);
/* MWNode::cvTransform(kind): src/trees/MWNode.cpp:448-490 -- scaling coefficients of the children (0/1 representation, i.e. after
 * MRX_RECONSTRUCTION) <-> function values at the children's quadrature points, for the listed nodes, in place. */
int mrx_node_cv_transform(mrx_tree *tree, int kind /* MRX_FORWARD | MRX_BACKWARD */, int n_nodes, const int *slots);
/* project(prec, out, f) (src/treebuilders/project.cpp:85-104) of f(r) = sum_i amp[i] prod_d cos(pi kvec[3 i + d] r_d): a native
 * callback for mrx_project_function (periodic test functions on the unit cell [-1, 1]^3) */
int mrx_project_cosines(mrx_tree *tree, double prec, int n_terms, const double *amp, const double *kvec, int finalize);
/* measurement: cvTransform(Forward) then (Backward) over every node, `reps` times between CUDA events; returns ms per pass
 * (algorithmic traffic 128 K^3 B per node and pass) */
double mrx_bench_cv_transform(mrx_tree *tree, int reps);
/* MWTree::calcSquareNorm: src/trees/MWTree.cpp:109-118 */
double mrx_calc_square_norm(mrx_tree *tree);
/* mrcpp::dot(bra, ket): src/treebuilders/multiply.cpp:286-318 */
double mrx_dot(mrx_tree *bra, mrx_tree *ket);
/* FunctionTree::rescale(c): src/trees/FunctionTree.cpp (coefficient-wise scaling) */
int mrx_tree_rescale(mrx_tree *tree, double c);

/* clear_grid(out) (src/treebuilders/grid.cpp:180-186): keep the grid, drop coefficients and norms. Host only. */
int mrx_tree_clear_grid(mrx_tree *tree);
/* build_grid(out, inp) (src/treebuilders/grid.cpp:144-153): extend the grid of `out` with every node of `inp` (union of the
 * two grids; coefficients of `out` are dropped). Host only. */
int mrx_tree_build_grid_from(mrx_tree *out, const mrx_tree *inp);
/* add(-1.0, out, {(coefs[i], inp[i])}, 0) (src/treebuilders/add.cpp:41-70, AdditionCalculator.h:42-66): sum of the inputs on
 * the grid `out` enters with, no refinement -- the form mrcpp::divergence uses (apply.cpp:527-528). Inputs coarser than the
 * grid contribute their generated (scaling-only) nodes, inputs finer than the grid are truncated, as in the reference. */
int mrx_tree_add(mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp);
/* add(prec, out, {(coefs[i], inp[i])}, max_iter, abs_prec) (add.cpp:41-70): the adaptive form -- starts from the grid `out`
 * enters with (normally empty roots) and refines where the wavelet norm of the sum asks for it (TreeBuilder.cpp:38-86,
 * WaveletAdaptor.h:51-54). prec < 0 or max_iter == 0: same as mrx_tree_add. */
int mrx_tree_add_adaptive(double prec, mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp, int max_iter, int abs_prec);

/* refine_grid(out, prec, absPrec) (scales <= 0) / refine_grid(out, scales) (scales > 0) (src/treebuilders/grid.cpp:271-302): one
 * pass of TreeBuilder::split over the end nodes (scales passes splitting every end node); the new children receive the
 * reconstruction of their parent. Returns the number of new nodes. A grid without coefficients is refined on the host only. */
int mrx_tree_refine_grid(mrx_tree *tree, double prec, int abs_prec, int scales);
/* FunctionTree::add(c, inp) in place on the grid of `tree` (src/trees/FunctionTree.cpp:687-706) */
int mrx_tree_add_inplace(mrx_tree *tree, double c, mrx_tree *inp);
/* multiply(prec, out, {(coefs[i], inp[i])}, max_iter, abs_prec, use_max_norms) (src/treebuilders/multiply.cpp:104-136 with
 * MultiplicationCalculator.h:43-72): point-wise product of the inputs from the grid `out` enters with (normally empty roots),
 * refined where the wavelet norm of the product asks for it (WaveletAdaptor) or, with use_max_norms != 0 and two inputs, where the
 * largest scaling / wavelet norms of the inputs do (makeMaxSquareNorms + MultiplicationAdaptor.h:46-66); prec < 0 or
 * max_iter == 0: no refinement. */
int mrx_tree_multiply(double prec, mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp, int max_iter, int abs_prec,
                      int use_max_norms);

/* power(prec, out, inp, p, max_iter, abs_prec) (src/treebuilders/multiply.cpp:211-234, PowerCalculator.h:43-58): the function
 * values of `inp` raised to the power p, refined like multiply */
int mrx_tree_power(double prec, mrx_tree *out, mrx_tree *inp, double p, int max_iter, int abs_prec);

/* residency control for measurement: host->device / device->host copies of a tree's coefficients */
int mrx_tree_sync_device(mrx_tree *tree); /* upload if the host copy is newer                        */
int mrx_tree_sync_host(mrx_tree *tree);   /* download if the device copy is newer                    */
/* keep the host copy of an apply OUTPUT current: mrx_apply then streams the result into the tree's pinned host chunks while it
 * runs (copy engines, beside the next refinement iteration's kernels) and returns with the tree in host memory, so that a
 * binding that reads coefficients on the host pays a fraction of the final download. In a sharded apply a rank with this
 * flag downloads the whole (replicated) result over its own link. */
int mrx_tree_set_host_mirror(mrx_tree *tree, int on);
/* Sharded apply, collective by convention (every rank calls it on its output tree before mrx_apply_sharded): the tree's
 * host chunks move into the communicator's shared host arena, every rank downloads the chunks it owns (chunk index modulo
 * world size) while the apply runs, and the call returns on every rank with the complete tree in that (shared) host
 * memory: N PCIe links carry the result instead of one. Returns 1 (and behaves like mrx_tree_set_host_mirror(tree, 1))
 * when the communicator has no arena. The host copy is ONE copy that all ranks see: every rank creates and frees its
 * shared-mirror trees in the same order (the arena is a bump allocator; the apply aborts if the ranks disagree on where the
 * tree lies), and the tree must not outlive the communicator. */
int mrx_tree_set_shared_host_mirror(mrx_tree *tree, mrx_comm *comm);
int mrx_tree_drop_device(mrx_tree *tree); /* free the HBM copy (next use uploads again)              */
long long mrx_tree_bytes(const mrx_tree *tree);

/* Host-object handles for the test oracle (oracle/ restates the reference on the same host data
 * model): mrx::Tree<3>* and mrx::Operator*. Not used by the product path. */
void *mrx_tree_host_handle(mrx_tree *tree);
void *mrx_oper_host_handle(mrx_oper *oper);
/* tell the library that the host copy was modified through the handle (device copy becomes stale) */
void mrx_tree_host_modified(mrx_tree *tree);

/* device-side stopwatch on the library's stream (CUDA events): start, run calls, stop -> elapsed ms */
void mrx_timer_start(void);
double mrx_timer_stop_ms(void);

/* micro-benchmarks used by bench.py for the roofline denominators (FP64 tensor pipe, HBM copy) */
double mrx_bench_dmma_tflops(int iters);
double mrx_bench_dfma_tflops(int iters);
double mrx_bench_hbm_gbs(long long bytes, int iters);
/* filter kernels of MWTree::mwTransform alone: the level launches of one TopDown(overwrite) (type MRX_TOP_DOWN) / BottomUp
 * (MRX_BOTTOM_UP) / TopDown(+=) (type 2: the mode mrcpp::apply closes with, apply.cpp:82; it accumulates, so the tree is scratch
 * afterwards) pass repeated `reps` times between two CUDA events; returns ms per pass, *branch_nodes = parent nodes transformed
 * per pass (algorithmic traffic 128 (k+1)^3 bytes -- 192 (k+1)^3 for += , which also reads the children -- and 96 (k+1)^4 flop each) */
double mrx_bench_mw_transform(mrx_tree *tree, int type, int reps, int *branch_nodes);

#ifdef __cplusplus
}
#endif
#endif
