// Internal glue between the C-ABI (cabi.cpp), the host data model (host/) and the CUDA engine (cuda/).
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <memory>
#include <vector>

#include "../../include/mrcpp_b200.h"
#include "host/mrx_host.hpp"

struct mrx_mra {
    mrx::MRA<3> m;
};

namespace mrx {

/// Caching device allocator (cabi.cpp): blocks are rounded up to size classes and never returned to the driver
/// while the library is loaded, so after the first apply every buffer (output coefficients that grow per
/// refinement iteration, tuple lists, partial sums) is served in microseconds. All work runs on the one
/// library stream, so reuse of a freed block is stream-ordered by construction. (cudaMallocAsync pools were
/// measured to stall 10-150 ms per growth step on the B200 boxes.)
void *dev_alloc(size_t bytes);
void dev_free(void *p);
size_t dev_cached_bytes();

/// Growable device buffer. Growth is geometric and preserves contents (device-to-device copy).
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n, bool keep, cudaStream_t st) {
        if (n <= cap) return;
        size_t ncap = n > 2 * cap ? n : 2 * cap;
        T *np = static_cast<T *>(dev_alloc(ncap * sizeof(T)));
        if (keep && p && cap) cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (p) dev_free(p);
        p = np;
        cap = ncap;
    }
    void release() {
        if (p) dev_free(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

/// HBM-resident node store of one function tree: SoA blocks in slot order.
struct DeviceTree {
    DevBuf<double> coefs;    // [nNodes][8][K^3]
    DevBuf<double> norms;    // [nNodes][8] component norms
    DevBuf<double> genCoefs; // [nGen][K^3]   generated (scaling-only) nodes, slot = nReal + i
    DevBuf<double> genNorms; // [nGen]
    int nNodes = 0;          // real nodes with valid device storage
    int nGen = 0;
    // band-walk topology of the tree as an apply INPUT (apply.cu): child0 / depth / node norm per real node; valid while
    // topoNodes == nReal, reset to -1 by everything that changes norms or structure (invalidate_topo)
    DevBuf<int> topoChild0, topoDepth;
    DevBuf<double> topoBound;
    int topoNodes = -1;
    double topoMaxNorm = 0.0;
    // lazy residency of an apply INPUT whose coefficients live in pinned host memory: storage for every node is
    // allocated, but only the nodes the apply actually reads are fetched over PCIe (by a gather kernel reading the host
    // chunks directly); resident[8 n + c]: 0 block c of node n not in HBM, 1 queued for the gather, 2 arrived. `partial` trees are not devValid: every other consumer
    // completes the upload first (tree_upload).
    DevBuf<int> resident;
    DevBuf<const double *> chunkTab; // host chunk base pointers (64 nodes each), device-readable
    bool partial = false;
};

struct DeviceOper {
    DevBuf<double> mats;   // all terms, all nodes: 4*K*K each
    DevBuf<double> norms;  // 4 per node
    DevBuf<int> nodeOff;   // [M][DM] global node index of transl = -maxTransl, -1 if depth absent
    DevBuf<int> maxTransl; // [M][DM]
    DevBuf<int> nodeBase;  // [M][DM] node index of translation 0 (nodeOff + maxTransl), -1 if absent
    DevBuf<int> bw;        // [M][DM][4]   1-D node counts 2 width + 1 per component (per apply: depends on prec)
    DevBuf<int> bsf;       // [M][DM][64]  band size factors (per apply)
    int M = 0, DM = 0;
    int identIdx = 0; // operator block index of the identity block appended to `mats`
    bool tablesValid = false;
    std::vector<size_t> termNodeBase; // host: first global node index of each term
};

} // namespace mrx

struct mrx_tree {
    mrx::Tree<3> host;
    mrx::DeviceTree dev;
    bool hostCoefsValid = true; // host coefficient chunks hold the current values
    bool devValid = false;      // device copy holds the current values
    // keep the host copy current: an apply writing this tree streams the result down while it runs (wavelet blocks of an
    // iteration's nodes are final when the iteration is; copy engines move them beside the next iteration's kernels) and
    // finishes the rest (scaling blocks, branch nodes) behind its closing passes, so that the tree is in host memory when the
    // call returns (mrx_tree_set_host_mirror)
    bool hostMirror = false;
    // sharded apply: the host chunks live in the host arena all ranks of `mirrorComm` have mapped (mrx_comm_host_arena); every
    // rank downloads the chunks it owns (chunk index % world) over ITS PCIe link, and the call returns on every rank with the
    // whole tree in that memory (mrx_tree_set_shared_host_mirror)
    const mrx_comm *mirrorComm = nullptr;
    explicit mrx_tree(const mrx::MRA<3> &m)
            : host(m) {}
};

struct mrx_oper {
    mrx::Operator op;
    mrx::DeviceOper dev;
    std::shared_ptr<void> bandCache; // per-depth band tables of the last (prec, direction) used (apply.cu)
};

namespace mrx {

bool device_enabled();
void require_device(const char *what); // aborts: there is no CPU fallback for the hot path
cudaStream_t stream();
long long &launch_counter();

// device_tree.cu
void tree_upload(mrx_tree &t);
void tree_lazy_begin(mrx_tree &t); // norms + bookkeeping on the device, coefficient blocks fetched on demand by the apply
void tree_download(mrx_tree &t);
void tree_drop_device(mrx_tree &t);
/// whole-tree transform (+ norms of every node). timedReps > 0: the level kernels are run timedReps times between two
/// CUDA events (measurement of the filter kernels alone; TopDown with overwrite and BottomUp are idempotent)
void device_mw_transform(mrx_tree &t, int type, bool overwrite, bool norms = true, int timedReps = 0, double *timedMs = nullptr,
                         int *branchNodes = nullptr);
/// TopDown(+=), BottomUp, norms, square norm: the closing passes of mrcpp::apply. pairsByDepth (optional): (parent, child0)
/// pairs per depth of ALL branch nodes, if the caller already has them
/// topDownDone: the TopDown(+=) steps were already run level by level inside the apply loop (apply.cu): BottomUp + norms only
void device_apply_post(mrx_tree &t, const std::vector<std::vector<int>> *pairsByDepth = nullptr, bool topDownDone = false);
/// pinned staging buffer for host->device uploads of host vectors (a pageable source is staged by the driver on one thread at a
/// fraction of the PCIe rate); contents must be consumed (stream synchronised) before the next call
double *pinned_stage(size_t doubles);
/// fn(begin, end) over [0, n) on a few host threads (std::thread: independent of OMP_NUM_THREADS, which launchers set to 1)
void host_parallel(size_t n, const std::function<void(size_t, size_t)> &fn);
void device_calc_norms_all(mrx_tree &t);                         // norms of every node -> host cnorm/sqn
/// MWNode::mwTransform (what = 0: kind 0 Compression, 1 Reconstruction) / MWNode::cvTransform (what = 1: kind 0 Forward, 1 Backward)
/// of the listed nodes (n < 0: every node) in place
void device_node_transform(mrx_tree &t, int what, int kind, int n, const int *slots, int timedReps = 0, double *timedMs = nullptr);
double device_dot(mrx_tree &bra, mrx_tree &ket);
void device_rescale(mrx_tree &t, double c);
/// add(prec, out, {(c_i, inp_i)}, maxIter, absPrec) from the grid of `out` (add.cpp:41-70); prec < 0 or maxIter = 0: no refinement
void device_add(mrx_tree &out, int n, const double *c, mrx_tree *const *inp, double prec = -1.0, int maxIter = 0, bool absPrec = false);
int device_refine_grid(mrx_tree &t, double prec, bool absPrec, int scales); // refine_grid (grid.cpp:271-302), returns the new nodes
void device_add_inplace(mrx_tree &out, double c, mrx_tree &inp);           // FunctionTree::add(c, inp) (FunctionTree.cpp:687-706)
/// multiply(prec, out, {(c_i, inp_i)}, maxIter, absPrec) from the grid of `out` (multiply.cpp:104-136)
/// power != nullptr: power(prec, out, inp[0], *power) (multiply.cpp:211-234): the values of the single input raised to *power
void device_multiply(mrx_tree &out, int n, const double *c, mrx_tree *const *inp, double prec, int maxIter, bool absPrec,
                     bool useMaxNorms = false, const double *power = nullptr);
void oper_upload(mrx_oper &o);

// project.cu
void device_project_gaussians(mrx_tree &t, double prec, const GaussExp<3> &gexp, int maxIter, bool absPrec);

// comm.cu
int comm_rank(const mrx_comm *c);
int comm_world(const mrx_comm *c);
void comm_allgatherv(const mrx_comm *c, void *base, const size_t *off, const size_t *count, cudaStream_t st);
void comm_allreduce_sum(const mrx_comm *c, double *buf, size_t n, cudaStream_t st);
void comm_allgather(const mrx_comm *c, void *base, size_t bytes, cudaStream_t st); // in place, equal segments
constexpr int kCommStageBufs = 3;
void comm_stage_reserve(mrx_comm *c, size_t bytes, cudaStream_t st); // collective
bool comm_peer_push_enabled(const mrx_comm *c);
char *comm_stage(const mrx_comm *c, int buf);
size_t comm_stage_bytes(const mrx_comm *c);
void comm_push(mrx_comm *c, int buf, size_t off, size_t bytes);
cudaEvent_t comm_ev_reduced(const mrx_comm *c, int buf);
cudaStream_t comm_unpack_stream(mrx_comm *c); // side stream of the row unpack (created on first use, with its events)
cudaEvent_t comm_ev_gathered(const mrx_comm *c);
cudaEvent_t comm_ev_unpacked(const mrx_comm *c, int buf);
cudaEvent_t comm_ev_pushed(const mrx_comm *c, int buf);
bool comm_has_host_arena(const mrx_comm *c); // mrx_comm_host_arena succeeded on all ranks
void *host_arena_alloc(size_t bytes);        // chunk allocator pair over the shared arena (Tree::rebaseChunks)
void host_arena_free(void *p);
long long host_arena_offset(const void *p);  // of an arena chunk from the arena base (the same on every rank for the same chunk)

// apply.cu
/// precTrees != nullptr: apply(prec, out, oper, inp, precTrees, maxIter, absPrec) (apply.cpp:214-251), precision scaled per
/// output node by the largest norms of the precision trees (an empty vector scales by 1)
/// unitCell: 0 plain apply; 1 / 2: apply_near_field / apply_far_field on a periodic world (apply.cpp:294-342): only the band
/// entries inside / outside the unit cell contribute
void device_apply(double prec, mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int maxIter, bool absPrec,
                  mrx_apply_stats *stats, const mrx_comm *comm = nullptr, const std::vector<mrx_tree *> *precTrees = nullptr,
                  int unitCell = 0);
void device_apply_derivative(mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int dir, mrx_apply_stats *stats);

} // namespace mrx
