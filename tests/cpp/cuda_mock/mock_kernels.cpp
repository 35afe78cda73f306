// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory): host stand-ins for the kernels the host drivers of
// csrc/cuda/device_tree.cu launch, and for the engine entry points that live in other .cu files (apply, projection, sharding,
// micro-benchmarks). The transform stand-ins use the CPU oracle's restatement of tree_utils::mw_transform(_back) (oracle.cpp is
// included into this translation unit); apply and projection are served by the oracle as a whole. What is exercised for real
// is the driver code: buffer sizes, slot arithmetic, pair lists, residency flags, the refinement loops of add / multiply /
// refine_grid.
#include "../../../oracle/oracle.cpp"

#include "../../../mrcpp_b200/csrc/engine.hpp"
#include "../../../mrcpp_b200/csrc/cuda/kernels.cuh"

namespace mrx {

void launch_norms(const double *coefs, double *norms, const int *slots, int n, int Kd, cudaStream_t, double *normsW) {
    for (int i = 0; i < n; i++) {
        const int node = slots ? slots[i] : i;
        for (int w = 0; w < 8; w++) {
            const double *b = coefs + ((size_t)node * 8 + w) * Kd;
            double s = 0.0;
            for (int e = 0; e < Kd; e++) s += b[e] * b[e];
            norms[(size_t)node * 8 + w] = std::sqrt(s);
            if (normsW) normsW[(size_t)i * 8 + w] = std::sqrt(s);
        }
    }
    launch_counter()++;
}

bool transform_fuses_norms(int) { return false; }

void launch_transform(bool down, bool overwrite, double *coefs, const int *pairs, int cnt, int K, const double *, cudaStream_t, double *) {
    const FilterSet &fs = filter_set(K - 1);
    const int Kd = K * K * K, ncoef = 8 * Kd;
    for (int p = 0; p < cnt; p++) {
        double *parent = coefs + (size_t)pairs[2 * p] * ncoef;
        double *child0 = coefs + (size_t)pairs[2 * p + 1] * ncoef;
        if (down) {
            if (overwrite) std::memset(child0, 0, sizeof(double) * 8 * ncoef); // giveChildrenCoefs(overwrite) zeroes the children
            orc::mw_transform3(fs, K, parent, child0, false, ncoef, overwrite);
        } else {
            std::vector<double> in((size_t)8 * Kd);
            for (int c = 0; c < 8; c++) std::memcpy(in.data() + (size_t)c * Kd, child0 + (size_t)c * ncoef, sizeof(double) * Kd);
            orc::mw_transform_back3(fs, K, in.data(), parent, Kd);
        }
    }
    launch_counter()++;
}

// in-node compression MWNode::mwTransform(Compression): the node's 8 blocks hold its children's scaling coefficients
void launch_compress_nodes(double *coefs, const int *pairs, int cnt, int K, const double *, cudaStream_t, double *) {
    const FilterSet &fs = filter_set(K - 1);
    const int Kd = K * K * K;
    for (int p = 0; p < cnt; p++) {
        double *node = coefs + (size_t)pairs[2 * p] * 8 * Kd;
        std::vector<double> in(node, node + (size_t)8 * Kd);
        orc::mw_transform_back3(fs, K, in.data(), node, Kd);
    }
    launch_counter()++;
}
// in-node reconstruction MWNode::mwTransform(Reconstruction): (s, d) -> the children's scaling blocks, in place
void launch_reconstruct_nodes(double *coefs, const int *pairs, int cnt, int K, const double *, cudaStream_t) {
    const FilterSet &fs = filter_set(K - 1);
    const int Kd = K * K * K;
    for (int p = 0; p < cnt; p++) {
        double *node = coefs + (size_t)pairs[2 * p] * 8 * Kd;
        std::vector<double> in(node, node + (size_t)8 * Kd);
        orc::mw_transform3(fs, K, in.data(), node, false, Kd, true);
    }
    launch_counter()++;
}
// MWNode::cvTransform for the interpolating basis: diagonal map per index and 2^(+-3 (n + 1) / 2)
void launch_cv_transform(double *coefs, const int *items, int cnt, int K, const double *map, bool backward, cudaStream_t) {
    const int Kd = K * K * K;
    for (int p = 0; p < cnt; p++) {
        double *c = coefs + (size_t)items[2 * p] * 8 * Kd;
        const int np1 = items[2 * p + 1] + 1;
        const double two_fac = backward ? std::sqrt(1.0 / std::exp2((double)(3 * np1))) : std::sqrt(std::exp2((double)(3 * np1)));
        for (int o = 0; o < 8 * Kd; o++) {
            const int q = o % Kd;
            c[o] = two_fac * (((c[o] * map[q % K]) * map[(q / K) % K]) * map[q / (K * K)]);
        }
    }
    launch_counter()++;
}
// ProjectionCalculator::calcNode up to cvTransform(Backward): values of the expansion at the expanded child quadrature points,
// times sqrt(w) per dimension and 2^(-3 (n + 1) / 2)  (what project_eval_kernel computes)
void launch_project_eval(double *coefs, const int *slots, const int4 *nodeInfo, int cnt, int K, const GaussTable &G, const double *roots,
                         const double *sqrtw, cudaStream_t) {
    const int Kd = K * K * K;
    for (int b = 0; b < cnt; b++) {
        const int scale = nodeInfo[b].x, l[3] = {nodeInfo[b].y, nodeInfo[b].z, nodeInfo[b].w};
        const double sFac = std::ldexp(1.0, -(scale + 1)), two_fac = std::sqrt(1.0 / std::ldexp(1.0, 3 * (scale + 1)));
        double *out = coefs + (size_t)slots[b] * 8 * Kd;
        for (int o = 0; o < 8 * Kd; o++) {
            const int tt = o / Kd, idx = o % Kd;
            const int j[3] = {idx % K, (idx / K) % K, idx / (K * K)};
            double r[3];
            for (int d = 0; d < 3; d++) r[d] = sFac * (roots[j[d]] + 2.0 * l[d] + (((tt >> d) & 1) ? 1.0 : 0.0));
            double s = 0.0;
            for (int g = 0; g < G.n; g++) {
                double q2 = 0.0, p2 = 1.0;
                for (int d = 0; d < 3; d++) {
                    const double q = r[d] - G.pos[3 * g + d];
                    q2 += G.alpha[g] * q * q;
                    const int pw = G.power[3 * g + d];
                    if (pw == 1) p2 *= q;
                    else if (pw != 0) p2 *= std::pow(q, (double)pw);
                }
                s += (q2 > 746.0) ? 0.0 * G.coef[g] * p2 : G.coef[g] * p2 * std::exp(-q2);
            }
            out[o] = two_fac * (((s * sqrtw[j[0]]) * sqrtw[j[1]]) * sqrtw[j[2]]);
        }
    }
    launch_counter()++;
}
void launch_gen_children(const double *, double *, double *, int, const int *, int, int, const double *, cudaStream_t) {
    MRX_ABORT("mock: launch_gen_children");
}
void launch_reduce_partials(double *, const double *, const int *, int, int, cudaStream_t) { MRX_ABORT("mock: launch_reduce_partials"); }

void launch_dot(const double *a, const double *b, const int *pairs, double *res, int np, int nRoots, int Kd, cudaStream_t) {
    for (int p = 0; p < np; p++) {
        const double *x = a + (size_t)pairs[2 * p] * 8 * Kd, *y = b + (size_t)pairs[2 * p + 1] * 8 * Kd;
        double s = 0.0;
        for (int j = (pairs[2 * p] < nRoots ? 0 : Kd); j < 8 * Kd; j++) s += x[j] * y[j]; // scaling blocks of the roots + every wavelet block
        res[p] = s;
    }
    launch_counter()++;
}

void launch_scale(double *x, size_t n, double c, cudaStream_t) {
    for (size_t i = 0; i < n; i++) x[i] *= c;
    launch_counter()++;
}

// The element-wise kernels of the tree algebra are executed from their REAL source: cpp_build.build_mock_lib() cuts the two
// __global__ functions out of csrc/cuda/kernels.cu into extracted_kernels.inc; with blockIdx / threadIdx as plain variables the
// launchers below run them block by block, thread by thread (no cross-thread communication in either kernel).
namespace simt {
struct Idx {
    int x = 0, y = 0, z = 0;
};
static Idx blockIdx, threadIdx;
#define __global__
#define __launch_bounds__(...)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
using std::exp2;
using std::pow;
using std::sqrt;
#include "extracted_kernels.inc"
#undef __global__
#undef __launch_bounds__
template <typename F> void run(int blocks, int threads, F kernel) {
    for (int b = 0; b < blocks; b++)
        for (int t = 0; t < threads; t++) {
            blockIdx.x = b;
            threadIdx.x = t;
            kernel();
        }
}
} // namespace simt

void launch_axpy_nodes(double *out, const double *in, const int *pairs, int np, int nRoots, int Kd, double c, cudaStream_t) {
    simt::run(np, 256, [&] { simt::axpy_nodes_kernel(out, in, pairs, nRoots, Kd, c); });
    launch_counter()++;
}

void launch_product_values(double *P, const double *S, const int *scale, int nC, int K, const double *map, double c, int mode, cudaStream_t) {
    simt::run(8 * nC, 256, [&] { simt::product_values_kernel(P, S, scale, nC, K, map, c, mode); });
    launch_counter()++;
}

// ---- engine entry points of the other .cu files, served by the oracle on the host copies
static void to_host(mrx_tree &t) {
    if (!t.hostCoefsValid) tree_download(t);
}
static void host_result(mrx_tree &t) { // the host copy is the current one, nothing valid on the "device"
    t.hostCoefsValid = true;
    t.devValid = false;
    t.dev.nNodes = 0;
    t.dev.nGen = 0;
    t.dev.topoNodes = -1;
    t.dev.partial = false;
}
static void host_storage(mrx_tree &t) {
    t.host.allocCoefs = true;
    t.host.ensureCoefStorage();
}
static void fill_stats(const orc::ApplyStats &st, mrx_apply_stats *stats) {
    if (!stats) return;
    std::memset(stats, 0, sizeof(*stats));
    stats->g_nodes = st.gNodes;
    stats->f_applied = st.fApplied;
    stats->f_applied_rank = st.fApplied;
    stats->gen_nodes = st.genUsed;
    stats->iterations = st.iters;
    stats->n_nodes_out = st.nNodesOut;
}
void device_apply(double prec, mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int maxIter, bool absPrec, mrx_apply_stats *stats, const mrx_comm *,
                  const std::vector<mrx_tree *> *precTrees, int unitCell) {
    to_host(inp);
    host_storage(out);
    orc::ApplyStats st;
    std::vector<mrx::Tree<3> *> pt;
    if (precTrees)
        for (mrx_tree *t : *precTrees) {
            to_host(*t);
            host_storage(*t); // the oracle generates nodes (with coefficients) below the leaves of the precision trees
            pt.push_back(&t->host);
        }
    orc::apply(prec, out.host, oper.op, inp.host, maxIter, absPrec, &st, precTrees ? &pt : nullptr, unitCell);
    host_result(out);
    fill_stats(st, stats);
}
void device_apply_derivative(mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int dir, mrx_apply_stats *stats) {
    to_host(inp);
    host_storage(out);
    orc::ApplyStats st;
    orc::apply_derivative(out.host, oper.op, inp.host, dir, &st);
    host_result(out);
    fill_stats(st, stats);
}
int comm_rank(const mrx_comm *) { return 0; }
int comm_world(const mrx_comm *) { return 1; }
bool comm_has_host_arena(const mrx_comm *) { return false; }
void *host_arena_alloc(size_t) { MRX_ABORT("mock: no communicator"); }
void host_arena_free(void *) {}
long long host_arena_offset(const void *) { return 0; }

} // namespace mrx

extern "C" {
int mrx_comm_unique_id(char *) { MRX_ABORT("mock: no communicator"); }
mrx_comm *mrx_comm_create(int, int, const char *) { MRX_ABORT("mock: no communicator"); }
void mrx_comm_destroy(mrx_comm *) {}
int mrx_comm_host_arena(mrx_comm *, long long) { return 1; }
int mrx_comm_rank(const mrx_comm *) { return 0; }
int mrx_comm_size(const mrx_comm *) { return 1; }
void mrx_shard_partition(const long long *, int, int, int *) { MRX_ABORT("mock: sharding"); }
void mrx_shard_cyclic(int, int, int, int *, int *) { MRX_ABORT("mock: sharding"); }
int mrx_shard_cyclic_row(int, int, int) { MRX_ABORT("mock: sharding"); }
int mrx_shard_block(void) { return 1; }
double mrx_bench_dmma_tflops(int) { return 0.0; }
double mrx_bench_dfma_tflops(int) { return 0.0; }
double mrx_bench_hbm_gbs(long long, int) { return 0.0; }
}
