// mrcpp::apply on the GPU: drivers of the operator application.
//
// Reference control flow replaced (file:line relative to the MRCPP tree):
//   apply<3,double>(prec,out,oper,inp,maxIter,absPrec)       src/treebuilders/apply.cpp:68-93
//   apply(out, DerivativeOperator, inp, dir)                  src/treebuilders/apply.cpp:379-412
//   TreeBuilder::build                                        src/treebuilders/TreeBuilder.cpp:38-86
//   ConvolutionCalculator::{initBandSizes,makeOperBand,calcNode,applyOperComp,applyOperator,
//                           tensorApplyOperComp}              src/treebuilders/ConvolutionCalculator.cpp:105-382
//   WaveletAdaptor::splitNode / tree_utils::split_check       src/treebuilders/WaveletAdaptor.h:51-54, tree_utils.cpp:47-65
//   TreeAdaptor::splitNodeVector                              src/treebuilders/TreeAdaptor.h:41-54
//
// run_apply_pipe (orders with a work-list contraction kernel, K = 4..12; one or many GPUs): per refinement
//   iteration  enumerate band (apply_enum.cu) -> screen / scan / fill / contract / reduce (apply_pipeline.cu) ->
//   [norm all-gather + peer push of the coefficient rows (comm.cu)] -> bookkeeping + split + next work vector
//   (apply_split.cu). The device owns the work vector; the host reads back three small records per iteration and replays
//   the split flags into its topology while the device contracts the next iteration.
// run_apply_legacy (other orders, MRX_LEGACY=1): band enumeration and refinement loop on the host, one CTA per output
//   node (apply_kernels.cu).
// device_apply / device_apply_derivative: residency of the input (whole upload or lazy gathers), the loop, the closing
//   TopDown(+=) / BottomUp passes (device_tree.cu), cleanup of generated nodes.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>

#include "../engine.hpp"
#include "apply_kernels.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace mrx {

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// MRX_PROFILE: report host-side calls that take longer than 3 ms (allocator growth, stalled copies)
struct SlowCall {
    const char *what;
    double t0;
    bool on;
    SlowCall(const char *w, bool enabled) : what(w), t0(enabled ? now_ms() : 0.0), on(enabled) {}
    ~SlowCall() {
        if (!on) return;
        double dt = now_ms() - t0;
        if (dt > 3.0) std::fprintf(stderr, "[mrx] slow host call: %s %.2f ms\n", what, dt);
    }
};

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

// Host mailbox of the refinement loop: three small records per iteration (enumeration counters, tuple / unit counts, split
// result) are published by a one-thread kernel into pinned, mapped host memory and the host spins on a sequence flag -- a few
// microseconds per read-back instead of a cudaMemcpyAsync + cudaStreamSynchronize round trip (MRX_NO_MAILBOX=1: the latter).
struct Mailbox {
    volatile unsigned flagEc, flagHdr, flagRes, pad;
    EnumCounters ec;
    PipeHeader hdr;
    SplitResult res;
};
static Mailbox *mailbox() {
    static Mailbox *mb = nullptr;
    if (!mb) {
        MRX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&mb), sizeof(Mailbox), cudaHostAllocMapped | cudaHostAllocPortable));
        std::memset(const_cast<unsigned *>(&mb->flagEc), 0, sizeof(Mailbox));
    }
    return mb;
}
static unsigned &mailbox_seq() {
    static unsigned seq = 0;
    return seq;
}
template <typename T> static void *mailbox_dev(T *hostPtr) {
    void *d = nullptr;
    MRX_CUDA(cudaHostGetDevicePointer(&d, const_cast<void *>(static_cast<const volatile void *>(hostPtr)), 0));
    return d;
}
// read a record back: mailbox (publish kernel + spin on the host flag) or plain copy + synchronise
template <typename T> static void read_back(T *dst, const T *devSrc, T *mbSlot, volatile unsigned *mbFlag, bool useMailbox, cudaStream_t st) {
    if (useMailbox) {
        const unsigned seq = ++mailbox_seq();
        launch_publish(devSrc, mailbox_dev(mbSlot), (int)(sizeof(T) / 4), static_cast<unsigned *>(mailbox_dev(mbFlag)), seq, st);
        // the publish kernel sits at the end of a dependent chain: if a kernel before it faults, the flag never comes. The spin looks
        // at the stream now and then and fails loudly instead of hanging the process
        unsigned long long spins = 0;
        const double tSpin = now_ms();
        while (*mbFlag != seq) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((++spins & ((1ull << 22) - 1)) == 0) {
                const cudaError_t e = cudaStreamQuery(st);
                if (e != cudaSuccess && e != cudaErrorNotReady) {
                    std::fprintf(stderr, "mrx abort: apply: the device reported '%s' while the host waited for a read-back\n", cudaGetErrorString(e));
                    std::abort();
                }
                if (now_ms() - tSpin > 120000.0) MRX_ABORT("apply: a read-back did not arrive within two minutes");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
        std::memcpy(dst, const_cast<const T *>(mbSlot), sizeof(T));
    } else {
        MRX_CUDA(cudaMemcpyAsync(dst, devSrc, sizeof(T), cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
    }
}

// host mirror of an apply output (mrx_tree::hostMirror): copy-engine stream of the per-iteration downloads
struct MirrorStream {
    cudaStream_t dl = nullptr;
    cudaEvent_t evReduced = nullptr, evCopied = nullptr;
    bool pending = false; // copies enqueued since the last wait
};
static MirrorStream &mirror_stream() {
    static MirrorStream m;
    if (!m.dl) {
        MRX_CUDA(cudaStreamCreateWithFlags(&m.dl, cudaStreamNonBlocking));
        MRX_CUDA(cudaEventCreateWithFlags(&m.evReduced, cudaEventDisableTiming));
        MRX_CUDA(cudaEventCreateWithFlags(&m.evCopied, cudaEventDisableTiming));
    }
    return m;
}

// closing TopDown(+=) folded into the refinement loop: the step parent level -> child level runs on a side stream as soon as the
// children's own coefficients are in the node store (one GPU: after the reduce; sharded: after the unpack), beside the next
// iteration's kernels, instead of level by level after the loop. Same kernel, same level order: bit-identical results.
struct TopDownStream {
    cudaStream_t td = nullptr;
    cudaEvent_t evReady = nullptr, evDone = nullptr;
    cudaEvent_t evBuf[3] = {nullptr, nullptr, nullptr}; // the step that read pair buffer x has finished
    bool pending = false, bufBusy[3] = {false, false, false};
};
static TopDownStream &topdown_stream() {
    static TopDownStream t;
    if (!t.td) {
        MRX_CUDA(cudaStreamCreateWithFlags(&t.td, cudaStreamNonBlocking));
        MRX_CUDA(cudaEventCreateWithFlags(&t.evReady, cudaEventDisableTiming));
        MRX_CUDA(cudaEventCreateWithFlags(&t.evDone, cudaEventDisableTiming));
        for (int x = 0; x < 3; x++) MRX_CUDA(cudaEventCreateWithFlags(&t.evBuf[x], cudaEventDisableTiming));
    }
    return t;
}

// lazy residency of the input tree: a large iteration runs in node sub-ranges, and the gather of sub-range s + 1 (its own stream, one
// CTA per SM: co-resident with the persistent contraction CTAs) runs BESIDE the contraction of sub-range s
struct FetchStream {
    cudaStream_t s = nullptr;
    cudaEvent_t evFilled[kMaxSubRanges] = {}, evFetched[kMaxSubRanges] = {};
    int grid = 0;
};
static FetchStream &fetch_stream() {
    static FetchStream f;
    if (!f.s) {
        int least = 0, greatest = 0, dev = 0;
        MRX_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        MRX_CUDA(cudaStreamCreateWithPriority(&f.s, cudaStreamNonBlocking, greatest));
        for (int x = 0; x < kMaxSubRanges; x++) {
            MRX_CUDA(cudaEventCreateWithFlags(&f.evFilled[x], cudaEventDisableTiming));
            MRX_CUDA(cudaEventCreateWithFlags(&f.evFetched[x], cudaEventDisableTiming));
        }
        MRX_CUDA(cudaGetDevice(&dev));
        MRX_CUDA(cudaDeviceGetAttribute(&f.grid, cudaDevAttrMultiProcessorCount, dev));
        if (getenv("MRX_FETCH_GRID")) f.grid = std::max(1, atoi(getenv("MRX_FETCH_GRID"))); // development switch
    }
    return f;
}

// final node count of the last apply of this process: the next apply reserves its node store for that many nodes up front
static size_t &nodeStoreHint() {
    static size_t h = 0;
    return h;
}

struct Scratch {
    DevBuf<GDesc> gdesc;
    DevBuf<NbrEntry> nbr;
    DevBuf<int> genItems;
    DevBuf<int> gslots;
    DevBuf<unsigned long long> counters;
    DevBuf<GDesc> units;
    DevBuf<int> reduceItems;
    DevBuf<double> partials;
    // work-list pipeline (apply_pipeline.cu)
    DevBuf<unsigned long long> masks;
    DevBuf<unsigned short> cnt64;
    DevBuf<int> segOff;
    DevBuf<int> blockCnt;
    DevBuf<unsigned> blockTupOff;
    DevBuf<int> blockUnitOff;
    DevBuf<PipeHeader> header;
    DevBuf<unsigned long long> pTileTotal, pTileBaseTup;
    DevBuf<int> pTileBaseUnit;
    DevBuf<TupleRec> tuples;
    DevBuf<UnitDesc> units2;
    DevBuf<int> queue;
    // device enumeration (apply_enum.cu)
    DevBuf<int4> gNodes;
    DevBuf<int> pending;
    DevBuf<int> newParents;
    DevBuf<EnumCounters> ecnt;
    DevBuf<int> chunkOff, pNode;
    DevBuf<unsigned long long> chunkPacked, chunkScan, tileTotal, tileBase;
    DevBuf<double> normsW[kCommStageBufs]; // component norms of the iteration's nodes, rank-major (one per staging buffer in flight)
    DevBuf<int> gslotsAll[kCommStageBufs]; // slots of the whole work vector (sharded apply: one per staging buffer in flight)
    // device-side refinement (apply_split.cu)
    DevBuf<int4> gAll[2];            // (depth, l) of the whole work vector, current / next
    DevBuf<unsigned char> isBranch, flags;
    bool hasBranchFlags = false;
    DevBuf<double> scaleFac, state;
    DevBuf<SplitResult> splitRes;
    // locally scaled precision (apply with precTrees, apply_prec.cu)
    DevBuf<int> tdPairs[3];             // (parent, first child) of the nodes that split in an iteration: TopDown(+=) level lists
    DevBuf<double> precAll[2], precLoc; // factors of the whole work vector (current / next) and of this rank's share
    DevBuf<PrecTreeDev> precTreeTab;
    std::vector<std::unique_ptr<DevBuf<double>>> precVReal;
    // lazy residency of the input tree
    DevBuf<int> fetchList, fetchCnt, fetchSnap; // queue, its length, its length after each fill pass of a sub-range
    DevBuf<unsigned long long> fetchTotal;
};

/// input-tree topology on the device for the band enumeration: real nodes + generated nodes in one slot space
struct DevTopo {
    DevBuf<int> child0, depth, flag;
    DevBuf<double> bound;
    void reserve(size_t n, cudaStream_t st) {
        if (n <= child0.cap) return;
        const size_t old = child0.cap;
        child0.reserve(n, true, st);
        depth.reserve(n, true, st);
        bound.reserve(n, true, st);
        flag.reserve(n, true, st);
        // new tail of the flag array must read 0
        MRX_CUDA(cudaMemsetAsync(flag.p + old, 0, sizeof(int) * (flag.cap - old), st));
    }
};

/// Host copy of the per-depth band tables (see DepthInfo). Built lazily for the depths that occur.
struct BandTables {
    double prec = 0.0; // cache key: band widths depend on the apply precision
    int derivDir = -2;
    DevBuf<DepthInfo> d_info;
    DevBuf<int> d_candOff;
    DevBuf<int> d_candTerm;
    DevBuf<unsigned long long> d_candMask;
    std::vector<DepthInfo> info;                // [DM]
    std::vector<char> built;                    // [DM]
    std::vector<int> candOff;                   // concatenated prefix arrays
    std::vector<int> candTerm;
    std::vector<unsigned long long> candMask;
    std::vector<std::vector<std::array<int, 4>>> needed; // per depth: (dx,dy,dz,code) of offsets with candidates
    std::vector<std::vector<double>> maxO;               // per depth, per needed offset: max_{term,combo} |O|^3 * bandSizeFactor
    // the same offsets flattened for the device enumeration (apply_enum.cu)
    std::vector<OffEntry> offs;
    std::vector<int> offStart, offCount; // [DM]
    DevBuf<OffEntry> d_offs;
    DevBuf<int> d_offStart, d_offCount;
    bool dirty = false;
    std::vector<int> bsf; // [M][DM][64] band size factors (ConvolutionCalculator::initBandSizes), mirrored in oper.dev.bsf

    // integer part of the screening: per-term max width (applyOperComp :283), per-dimension band test per
    // component (applyOperator :311-318, OperatorTree::isOutsideBand), T block only at depth 0 (calcNode :261).
    // derivative operators: DerivativeCalculator::applyOperator (:211-249).
    void build(const Operator &op, int depth, int derivDir, const std::vector<int> &bsf, int DM) {
        const int M = op.size();
        const int W = op.getMaxBandWidth(depth);
        info[depth].W = W;
        info[depth].cubeOff = (int)candOff.size();
        built[depth] = 1;
        dirty = true;
        if (W < 0) {
            candOff.push_back((int)candTerm.size());
            return;
        }
        const int cube = 2 * W + 1;
        // per term, per distance a: bitmask over the 64 (gt,ft) combos whose component along a dimension is in band
        // combo bit b = gt*8+ft; component along d: c_d = 2*gt_d + ft_d
        std::vector<unsigned long long> dimMask((size_t)M * 3 * (W + 1), 0ull);
        for (int t = 0; t < M; t++) {
            const OperTerm &ot = op.terms[t];
            for (int d = 0; d < 3; d++)
                for (int a = 0; a <= W; a++) {
                    unsigned long long m = 0ull;
                    for (int b = 0; b < 64; b++) {
                        int gt = b >> 3, ft = b & 7;
                        int c = 2 * ((gt >> d) & 1) + ((ft >> d) & 1);
                        bool ok = a <= ot.width(depth, c);
                        if (derivDir >= 0 && d != derivDir) ok = ok && (a == 0) && (c == 0 || c == 3);
                        if (ok) m |= 1ull << b;
                    }
                    dimMask[((size_t)t * 3 + d) * (W + 1) + a] = m;
                }
        }
        const unsigned long long tExcl = (derivDir < 0 && depth != 0) ? ~1ull : ~0ull; // drop (gt=0,ft=0)
        needed[depth].clear();
        maxO[depth].clear();
        for (int z = -W; z <= W; z++)
            for (int y = -W; y <= W; y++)
                for (int x = -W; x <= W; x++) {
                    const int code = ((z + W) * cube + (y + W)) * cube + (x + W);
                    candOff.push_back((int)candTerm.size());
                    const int ax = std::abs(x), ay = std::abs(y), az = std::abs(z);
                    const int maxD = std::max(ax, std::max(ay, az));
                    const size_t before = candTerm.size();
                    double mo = 0.0;
                    for (int t = 0; t < M; t++) {
                        if (derivDir < 0 && maxD > op.terms[t].maxWidth(depth)) continue;
                        unsigned long long m = dimMask[((size_t)t * 3 + 0) * (W + 1) + ax] &
                                               dimMask[((size_t)t * 3 + 1) * (W + 1) + ay] &
                                               dimMask[((size_t)t * 3 + 2) * (W + 1) + az] & tExcl;
                        if (m) {
                            candTerm.push_back(t);
                            candMask.push_back(m);
                            if (derivDir < 0) {
                                // upper bound of oNorm * bandSizeFactor over the allowed combos (same product order as the kernel)
                                const OperTerm &ot = op.terms[t];
                                const double *n0 = ot.nodeNorms(depth, x), *n1 = ot.nodeNorms(depth, y), *n2 = ot.nodeNorms(depth, z);
                                const int *bs = bsf.data() + ((size_t)t * DM + depth) * 64;
                                unsigned long long mm = m;
                                while (mm) {
                                    int b = __builtin_ctzll(mm);
                                    mm &= mm - 1;
                                    int gt = b >> 3, ft = b & 7;
                                    double o = 1.0;
                                    o *= n0[2 * (gt & 1) + (ft & 1)];
                                    o *= n1[2 * ((gt >> 1) & 1) + ((ft >> 1) & 1)];
                                    o *= n2[2 * ((gt >> 2) & 1) + ((ft >> 2) & 1)];
                                    mo = std::max(mo, o * bs[b]);
                                }
                            }
                        }
                    }
                    if (candTerm.size() != before) {
                        needed[depth].push_back({x, y, z, code});
                        maxO[depth].push_back(mo);
                    }
                }
        candOff.push_back((int)candTerm.size());
        offStart[depth] = (int)offs.size();
        offCount[depth] = (int)needed[depth].size();
        for (size_t q = 0; q < needed[depth].size(); q++) {
            const auto &o = needed[depth][q];
            offs.push_back(OffEntry{o[0], o[1], o[2], o[3], maxO[depth][q]});
        }
    }
};

} // namespace

/// band tables of (operator, prec, direction), cached on the operator; band size factors uploaded once per cache entry
static BandTables &get_band_tables(mrx_oper &oper, double prec, int derivDir, cudaStream_t st) {
    Operator &op = oper.op;
    const int M = op.size(), DM = oper.dev.DM;
    // per-depth band tables are cached on the operator: they depend only on (operator, prec, direction)
    std::shared_ptr<BandTables> btp = std::static_pointer_cast<BandTables>(oper.bandCache);
    if (!btp || btp->prec != prec || btp->derivDir != derivDir || (int)btp->info.size() != DM) {
        btp = std::make_shared<BandTables>();
        btp->prec = prec;
        btp->derivDir = derivDir;
        btp->info.assign(DM, DepthInfo{-1, 0});
        btp->built.assign(DM, 0);
        btp->needed.resize(DM);
        btp->maxO.resize(DM);
        btp->offStart.assign(DM, 0);
        btp->offCount.assign(DM, 0);
        oper.bandCache = btp;
    }
    BandTables &bt = *btp;
    // band size factors depend on the band widths, i.e. on (operator, prec): computed and uploaded once per cache entry
    if (bt.bsf.empty()) {
        std::vector<int> bwTab;
        band_size_factors(op, DM, bt.bsf, bwTab);
        oper.dev.bsf.reserve(bt.bsf.size(), false, st);
        MRX_CUDA(cudaMemcpyAsync(oper.dev.bsf.p, bt.bsf.data(), sizeof(int) * bt.bsf.size(), cudaMemcpyHostToDevice, st));
        // the same factors in separated form: bsf[gt*8+ft] = 64 * prod_d nodes1d[2 gt_d + ft_d] (calcBandSizeFactor :125-139)
        std::vector<int> sep((size_t)M * DM * 4, 1);
        for (int i = 0; i < M; i++)
            for (int depth = 0; depth < DM; depth++)
                for (int c = 0; c < 4; c++) {
                    const int w = bwTab[((size_t)i * DM + depth) * 5 + c];
                    sep[((size_t)i * DM + depth) * 4 + c] = (w < 0) ? 1 : 2 * w + 1;
                }
        oper.dev.bw.reserve(sep.size(), false, st);
        MRX_CUDA(cudaMemcpyAsync(oper.dev.bw.p, sep.data(), sizeof(int) * sep.size(), cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaStreamSynchronize(st));
    }
    return bt;
}

/// (re)upload the band tables after new depths were built
static void upload_band_tables(BandTables &bt, int DM, bool usePipe, cudaStream_t st) {
    if (bt.dirty) {
        bt.d_info.reserve(DM, false, st);
        bt.d_candOff.reserve(bt.candOff.size(), false, st);
        bt.d_candTerm.reserve(std::max<size_t>(bt.candTerm.size(), 1), false, st);
        bt.d_candMask.reserve(std::max<size_t>(bt.candMask.size(), 1), false, st);
        MRX_CUDA(cudaMemcpyAsync(bt.d_info.p, bt.info.data(), sizeof(DepthInfo) * DM, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(bt.d_candOff.p, bt.candOff.data(), sizeof(int) * bt.candOff.size(), cudaMemcpyHostToDevice, st));
        if (!bt.candTerm.empty()) {
            MRX_CUDA(cudaMemcpyAsync(bt.d_candTerm.p, bt.candTerm.data(), sizeof(int) * bt.candTerm.size(),
                                     cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemcpyAsync(bt.d_candMask.p, bt.candMask.data(), sizeof(unsigned long long) * bt.candMask.size(),
                                     cudaMemcpyHostToDevice, st));
        }
        if (usePipe) {
            bt.d_offs.reserve(std::max<size_t>(bt.offs.size(), 1), false, st);
            bt.d_offStart.reserve(DM, false, st);
            bt.d_offCount.reserve(DM, false, st);
            if (!bt.offs.empty())
                MRX_CUDA(cudaMemcpyAsync(bt.d_offs.p, bt.offs.data(), sizeof(OffEntry) * bt.offs.size(), cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemcpyAsync(bt.d_offStart.p, bt.offStart.data(), sizeof(int) * DM, cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemcpyAsync(bt.d_offCount.p, bt.offCount.data(), sizeof(int) * DM, cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaStreamSynchronize(st)); // host vectors may be reallocated by the next build()
        }
        bt.dirty = false;
    }
}

/// pristine band-walk topology of an input tree (cached on the tree)
static void ensure_input_topology(mrx_tree &inp, cudaStream_t st) {
    Tree<3> &f = inp.host;
    const int fRealN = f.nReal;
    // pristine topology (child pointers, depths, node norms) of the input tree: cached on the tree, copied device to device
    // per apply (generated nodes hang new children below real leaves while an apply runs)
    DeviceTree &fd = inp.dev;
    if (fd.topoNodes != fRealN) {
        // built straight into pinned staging memory on a few host threads: [bound: n doubles][child0: n ints][depth: n ints]
        double *stage = pinned_stage((size_t)2 * fRealN + 2);
        double *hBound = stage;
        int *hChild0 = reinterpret_cast<int *>(stage + fRealN), *hDepth = hChild0 + fRealN;
        std::vector<double> mxT(64, 0.0);
        std::atomic<int> slot{0};
        const int rootScale = f.mra.rootScale;
        host_parallel((size_t)fRealN, [&](size_t a, size_t b) {
            double m = 0.0;
            for (size_t n = a; n < b; n++) {
                const double v = std::sqrt(f.sqn[n]);
                hBound[n] = v;
                m = std::max(m, v);
                const int c0 = f.nodes[n].child0;
                hChild0[n] = (c0 >= 0 && c0 < fRealN) ? c0 : -1;
                hDepth[n] = f.nodes[n].scale - rootScale;
            }
            mxT[slot.fetch_add(1) & 63] = m;
        });
        double mx = 0.0;
        for (double v : mxT) mx = std::max(mx, v);
        fd.topoChild0.reserve(fRealN, false, st);
        fd.topoDepth.reserve(fRealN, false, st);
        fd.topoBound.reserve(fRealN, false, st);
        MRX_CUDA(cudaMemcpyAsync(fd.topoBound.p, hBound, sizeof(double) * fRealN, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(fd.topoChild0.p, hChild0, sizeof(int) * fRealN, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(fd.topoDepth.p, hDepth, sizeof(int) * fRealN, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        fd.topoNodes = fRealN;
        fd.topoMaxNorm = mx;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// First-generation path, kept for the orders the work-list pipeline does not cover (k < 3, k > 11) and as
// a development cross-check (MRX_LEGACY=1): band enumeration and the refinement loop on the host, one CTA per output
// node (apply_kernels.cu). Single GPU only.
static void run_apply_legacy(double prec, mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int maxIter, bool absPrec, int derivDir,
                             std::vector<int> workVec, mrx_apply_stats &S) {
    cudaStream_t st = stream();
    const double tEnter = now_ms();
    Operator &op = oper.op;
    const int M = op.size(), DM = oper.dev.DM;
    const bool deriv = derivDir >= 0;
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
    BandTables &bt = get_band_tables(oper, prec, derivDir, st);
    const std::vector<int> &bsf = bt.bsf;

    cudaEvent_t ev0, ev1;
    MRX_CUDA(cudaEventCreate(&ev0));
    MRX_CUDA(cudaEventCreate(&ev1));
    float kernel_ms = 0.f;

    double sNorm = 0.0, wNorm = 0.0;
    int iter = 0;
    const int fRealN = f.nReal;
    std::vector<GDesc> gdesc;
    std::vector<NbrEntry> nbr;
    std::vector<int> newParents;
    std::vector<double> genBound; // per generated node: norm of its real leaf ancestor (upper bound of its own norm)
    ensure_input_topology(inp, st);
    DeviceTree &fd = inp.dev;
    const double fMaxNorm = fd.topoMaxNorm;
    std::vector<double> fNodeNorm; // host enumeration
    fNodeNorm.resize(fRealN);
    for (int n = 0; n < fRealN; n++) fNodeNorm[n] = std::sqrt(f.sqn[n]);
    double tp_enum = 0, tp_phase2 = 0, tp_gen = 0, tp_upload = 0, tp_wait = 0, tp_host = 0, tp_tables = 0;
    const bool profile = getenv("MRX_PROFILE") != nullptr;
    const double tLoop = now_ms();
    while (!workVec.empty()) {
        const int nG = (int)workVec.size();
        double tq = now_ms();
        // ---- band enumeration (makeOperBand/fillOperBand, :142-222, non-periodic) on the topology,
        //      restricted to offsets that at least one (term, gt, ft) can reach
        gdesc.resize(nG);
        nbr.clear();
        newParents.clear();
        double gThrsIter = g.squareNorm;
        if (gThrsIter > 0.0) gThrsIter = prec * 1.0 * std::sqrt(gThrsIter / static_cast<double>(M));
        const bool screenOn = !deriv && gThrsIter >= 0.0;
        // global cut: offsets whose largest operator-norm product cannot lift even the largest input node over gThrs
        // are dropped for every node of this iteration (per depth, computed on first use)
        std::vector<std::vector<int>> live(DM); // indices into bt.needed[depth]
        std::vector<char> liveBuilt(DM, 0);
        auto live_list = [&](int depth) -> const std::vector<int> & {
            if (!liveBuilt[depth]) {
                liveBuilt[depth] = 1;
                const auto &mo = bt.maxO[depth];
                live[depth].reserve(mo.size());
                for (size_t oi = 0; oi < mo.size(); oi++)
                    if (!screenOn || mo[oi] * fMaxNorm * (1.0 + 1e-9) > gThrsIter) live[depth].push_back((int)oi);
            }
            return live[depth];
        };
        for (int i = 0; i < nG; i++) {
            const int dep = g.nodes[workVec[i]].scale - op.operRoot;
            if (dep >= 0 && dep < DM) {
                if (!bt.built[dep]) bt.build(op, dep, derivDir, bsf, DM);
                live_list(dep);
            }
        }
        tp_tables += now_ms() - tq;
        tq = now_ms();
        // phase 1 (parallel over output nodes): enumerate the band on the topology, prune, record the deepest
        // existing input node per surviving offset
        struct Hit {
            int node;
            int code;
            std::array<int, 3> l;
        };
        const int nGH = nG;
        std::vector<std::vector<Hit>> hits(nGH);
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = 0; i < nGH; i++) {
            const auto &nd = g.nodes[workVec[i]];
            const int dep = nd.scale - op.operRoot;
            if (dep < 0 || dep >= DM) continue; // deeper than every operator tree: empty band (:146-151)
            if (bt.info[dep].W < 0) continue;
            int lo[3], hi[3];
            for (int x = 0; x < 3; x++) {
                int nboxes = f.mra.nboxes[x] * (1 << dep);
                lo[x] = f.mra.corner[x] * (1 << dep);
                hi[x] = lo[x] + nboxes - 1;
            }
            const auto &need = bt.needed[dep];
            const auto &mo = bt.maxO[dep];
            auto &out_hits = hits[i];
            int cur = -1; // deepest existing input node found for the previous offset (locality: x runs fastest)
            for (int oi : live[dep]) {
                const auto &o = need[oi];
                std::array<int, 3> l = {nd.l[0] + o[0], nd.l[1] + o[1], nd.l[2] + o[2]};
                if (l[0] < lo[0] || l[0] > hi[0] || l[1] < lo[1] || l[1] > hi[1] || l[2] < lo[2] || l[2] > hi[2]) continue;
                // walk up from the previous hit until an ancestor of l, then down through existing nodes
                while (cur >= 0) {
                    const auto &cn = f.nodes[cur];
                    int sh = nd.scale - cn.scale;
                    if (sh >= 0 && (l[0] >> sh) == cn.l[0] && (l[1] >> sh) == cn.l[1] && (l[2] >> sh) == cn.l[2]) break;
                    cur = cn.parent;
                }
                if (cur < 0) cur = f.rootIndex(nd.scale, l);
                while (f.nodes[cur].scale < nd.scale && f.nodes[cur].child0 >= 0) {
                    int shift = nd.scale - f.nodes[cur].scale - 1;
                    int c = ((l[0] >> shift) & 1) | (((l[1] >> shift) & 1) << 1) | (((l[2] >> shift) & 1) << 2);
                    cur = f.nodes[cur].child0 + c;
                }
                if (screenOn) {
                    // no (term, gt, ft) can pass the norm screening if even the largest operator-norm product times
                    // an upper bound of |f| stays below gThrs. |f_ft| <= |node| for an existing real node, and a
                    // generated node is an orthogonal projection of its real leaf ancestor: |gen| <= |leaf|.
                    double fb = (cur < fRealN) ? fNodeNorm[cur] : genBound[cur - fRealN];
                    if (mo[oi] * fb * (1.0 + 1e-9) <= gThrsIter) continue;
                }
                out_hits.push_back({cur, o[3], l});
            }
        }
        tp_enum += now_ms() - tq;
        tq = now_ms();
        // phase 2 (serial, work-vector order): create the missing generated nodes, emit neighbour entries
        long long nCand = 0;
        int nNbr = 0;
        for (int i = 0; i < nGH; i++) {
            const auto &nd = g.nodes[workVec[i]];
            GDesc &d = gdesc[i];
            d.slot = workVec[i];
            d.depth = nd.scale - op.operRoot;
            d.nbrOff = (int)nbr.size();
            const int *coffG = (d.depth >= 0 && d.depth < DM && bt.built[d.depth]) ? bt.candOff.data() + bt.info[d.depth].cubeOff : nullptr;
            for (const Hit &h : hits[i]) {
                int node = h.node;
                if (f.nodes[node].scale < nd.scale) {
                    double fb = (node < fRealN) ? fNodeNorm[node] : genBound[node - fRealN];
                    node = f.getNodeTopo(nd.scale, h.l, &newParents); // continues below existing nodes
                    genBound.resize(f.size() - fRealN, fb);
                }
                NbrEntry e;
                e.fslot = node;
                e.code = h.code;
                e.g = i;
                e.candBase = (int)nCand;
                nCand += coffG[h.code + 1] - coffG[h.code];
                nbr.push_back(e);
            }
            d.nbrCnt = (int)nbr.size() - d.nbrOff;
            d.partial = -1;
        }
        // ---- work units: nodes with much work are split into chunks of their neighbour list (one CTA each, partial
        //      sums reduced in chunk order afterwards); units are launched in order of decreasing cost (LPT)
        std::vector<GDesc> units;
        std::vector<long long> unitCost;
        std::vector<int> reduceItems; // (slot, firstPartial, nPartials)
        int nPartials = 0;
        if (nCand >= (1ll << 31)) MRX_ABORT("apply: candidate space of one iteration exceeds 2^31");
        std::vector<long long> cost(nG, 0);
        long long total = 0;
        for (int i = 0; i < nG; i++) {
            const GDesc &d = gdesc[i];
            if (d.nbrCnt == 0) continue;
            const int *coff = bt.candOff.data() + bt.info[d.depth].cubeOff;
            long long c = 0;
            for (int q = 0; q < d.nbrCnt; q++) {
                int code = nbr[d.nbrOff + q].code;
                c += coff[code + 1] - coff[code];
            }
            cost[i] = c;
            total += c;
        }
        const long long target = std::max<long long>(total / (148 * 4), 512);
        for (int i = 0; i < nG; i++) {
            const GDesc &d = gdesc[i];
            int nChunks = (int)std::min<long long>((cost[i] + target - 1) / target, std::max(d.nbrCnt, 1));
            if (nChunks <= 1) {
                units.push_back(d);
                unitCost.push_back(cost[i]);
                continue;
            }
            const int *coff = bt.candOff.data() + bt.info[d.depth].cubeOff;
            reduceItems.push_back(d.slot);
            reduceItems.push_back(nPartials);
            int made = 0, q0 = 0;
            long long acc = 0, done = 0;
            for (int q = 0; q < d.nbrCnt; q++) {
                int code = nbr[d.nbrOff + q].code;
                acc += coff[code + 1] - coff[code];
                // close the chunk when its share of the remaining cost is reached
                long long want = (cost[i] - done + (nChunks - made) - 1) / (nChunks - made);
                if ((acc >= want && made < nChunks - 1) || q == d.nbrCnt - 1) {
                    GDesc u = d;
                    u.nbrOff = d.nbrOff + q0;
                    u.nbrCnt = q - q0 + 1;
                    u.partial = nPartials++;
                    units.push_back(u);
                    unitCost.push_back(acc);
                    done += acc;
                    acc = 0;
                    q0 = q + 1;
                    made++;
                }
            }
            reduceItems.push_back(made);
        }
        std::vector<int> order(units.size());
        for (size_t u = 0; u < order.size(); u++) order[u] = (int)u;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return unitCost[a] > unitCost[b]; });
        std::vector<GDesc> sorted(units.size());
        for (size_t u = 0; u < order.size(); u++) sorted[u] = units[order[u]];
        units.swap(sorted);
        tp_phase2 += now_ms() - tq;
        tq = now_ms();
        upload_band_tables(bt, DM, false, st);
        // ---- generated input nodes: parents in creation order; a parent created this iteration must be
        //      filled before its own children -> waves
        if (!newParents.empty()) {
            int nGenTotal = f.size() - fRealN;
            inp.dev.genCoefs.reserve((size_t)nGenTotal * Kd, true, st);
            inp.dev.genNorms.reserve((size_t)nGenTotal, true, st);
            std::vector<int> items;
            size_t pos = 0;
            while (pos < newParents.size()) {
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
        tp_gen += now_ms() - tq;
        tq = now_ms();
        // ---- device storage for the output nodes of this iteration
        {
            SlowCall sc("reserve output coefficients", profile);
            out.dev.coefs.reserve((size_t)g.nReal * ncoef, true, st);
            out.dev.norms.reserve((size_t)g.nReal * 8, true, st);
        }
        out.dev.nNodes = g.nReal;
        SlowCall *scUp = new SlowCall("descriptor uploads", profile);

        scr.units.reserve(std::max<size_t>(units.size(), 1), false, st);
        MRX_CUDA(cudaMemcpyAsync(scr.units.p, units.data(), sizeof(GDesc) * units.size(), cudaMemcpyHostToDevice, st));
        if (nPartials > 0) {
            scr.partials.reserve((size_t)nPartials * ncoef, false, st);
            scr.reduceItems.reserve(reduceItems.size(), false, st);
            MRX_CUDA(cudaMemcpyAsync(scr.reduceItems.p, reduceItems.data(), sizeof(int) * reduceItems.size(), cudaMemcpyHostToDevice, st));
        }
        scr.gdesc.reserve(nG, false, st);
        scr.nbr.reserve(std::max<size_t>(nbr.size(), 1), false, st);
        scr.gslots.reserve(nG, false, st);
        MRX_CUDA(cudaMemcpyAsync(scr.gdesc.p, gdesc.data(), sizeof(GDesc) * nG, cudaMemcpyHostToDevice, st));
        if (!nbr.empty())
            MRX_CUDA(cudaMemcpyAsync(scr.nbr.p, nbr.data(), sizeof(NbrEntry) * nbr.size(), cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(scr.gslots.p, workVec.data(), sizeof(int) * nG, cudaMemcpyHostToDevice, st));
        nNbr = (int)nbr.size();

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
        P.gdesc = scr.units.p;
        P.partials = scr.partials.p;
        P.nbr = scr.nbr.p;
        P.mats = oper.dev.mats.p;
        P.onorms = oper.dev.norms.p;
        P.nodeBase = oper.dev.nodeBase.p;
        P.bsf = oper.dev.bsf.p;
        P.bsfSep = reinterpret_cast<const int4 *>(oper.dev.bw.p);
        P.depthInfo = bt.d_info.p;
        P.candOff = bt.d_candOff.p;
        P.candTerm = bt.d_candTerm.p;
        P.candMask = bt.d_candMask.p;
        P.M = M;
        P.DM = DM;
        P.K = K;
        P.gThrs = gThrs;
        P.counters = scr.counters.p;
        P.derivDir = derivDir;
        P.identIdx = oper.dev.identIdx;

        delete scUp;
        tp_upload += now_ms() - tq;
        tq = now_ms();
        MRX_CUDA(cudaEventRecord(ev0, st));
        launch_apply(P, (int)units.size(), st);
        if (nPartials > 0) launch_reduce_partials(out.dev.coefs.p, scr.partials.p, scr.reduceItems.p, (int)reduceItems.size() / 3, ncoef, st);
        MRX_CUDA(cudaEventRecord(ev1, st));
        // calcNorms of the output nodes (ConvolutionCalculator.cpp:270-272)
        launch_norms(out.dev.coefs.p, out.dev.norms.p, scr.gslots.p, nG, Kd, st);
        // norms back to the host: the slot range covering the work vector
        int lo = 0;
        std::vector<double> range;
        lo = *std::min_element(workVec.begin(), workVec.end());
        int hi = *std::max_element(workVec.begin(), workVec.end());
        range.resize((size_t)(hi - lo + 1) * 8);
        MRX_CUDA(cudaMemcpyAsync(range.data(), out.dev.norms.p + (size_t)lo * 8, sizeof(double) * range.size(),
                                 cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        tp_wait += now_ms() - tq;
        tq = now_ms();
        float ms = 0.f;
        MRX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        kernel_ms += ms;
        if (profile) std::fprintf(stderr, "[mrx] iter %d nG %d nbr %d cand %lld kernels %.3f ms\n", iter, nG, nNbr, nCand, ms);
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
        if (!deriv) {
            // tree_utils::split_check (tree_utils.cpp:47-65) with the wavelet threshold of a scale computed once per
            // iteration (same expression, same operation order: the decision is bit-identical)
            std::vector<double> wThrs(maxScale - g.mra.rootScale + 2, -1.0);
            const bool precOn = prec > 0.0;
            double t_norm = 1.0;
            if (g.squareNorm > 0.0 and not absPrec) t_norm = std::sqrt(g.squareNorm);
            newVec.reserve(workVec.size());
            for (int n : workVec) {
                if (g.isBranch(n)) continue;
                const int scale = g.nodes[n].scale;
                if (scale + 2 > maxScale) continue;
                if (!precOn) continue;
                double &thr = wThrs[scale - g.mra.rootScale];
                if (thr < 0.0) {
                    const double expo = 0.5 * 1.0 * (scale + 1);
                    const double scale_fac = std::pow(2.0, -expo);
                    thr = std::max(2.0 * MachinePrec, prec * t_norm * scale_fac);
                }
                const double w_norm = std::sqrt(g.waveletNorm(n));
                if (w_norm > thr) {
                    int c0 = g.createChildren(n, false);
                    for (int c = 0; c < 8; c++) newVec.push_back(c0 + c);
                }
            }
        }
        workVec.swap(newVec);
        iter++;
        tp_host += now_ms() - tq;
    }
    const double tLoopEnd = now_ms();
    if (profile)
        std::fprintf(stderr, "[mrx] run_apply ms: pre-loop %.2f loop %.2f\n", tLoop - tEnter, tLoopEnd - tLoop);
    if (profile)
        std::fprintf(stderr, "[mrx] host phases ms: tables %.2f enum %.2f phase2 %.2f gen %.2f upload %.2f wait %.2f split %.2f\n", tp_tables,
                     tp_enum, tp_phase2, tp_gen, tp_upload, tp_wait, tp_host);
    S.iterations = iter;
    S.ms_kernel = kernel_ms;
    S.ms_contract = kernel_ms;
    unsigned long long counters[4];
    MRX_CUDA(cudaMemcpyAsync(counters, scr.counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
    MRX_CUDA(cudaStreamSynchronize(st));
    S.f_applied = (long long)counters[0];
    S.f_applied_rank = S.f_applied;
    out.dev.nNodes = g.nReal;
    MRX_CUDA(cudaEventDestroy(ev0));
    MRX_CUDA(cudaEventDestroy(ev1));
    if (profile) std::fprintf(stderr, "[mrx] run_apply ms: post-loop %.2f\n", now_ms() - tLoopEnd);
}


// ---------------------------------------------------------------------------------------------------------------------
// Work-list pipeline with the refinement step on the device. Per iteration the host launches kernels and reads back
// three small records (enumeration counters, tuple/unit counts, split result); it replays the split decisions into its
// own topology AFTER it has launched the next iteration's contraction, i.e. while the device is busy. Nothing on the
// host's critical path loops over nodes (except the one-off set-up of the first work vector).
static void run_apply_pipe(double prec, mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int maxIter, bool absPrec, int derivDir,
                           std::vector<int> workVec, mrx_apply_stats &S, mrx_comm *comm,
                           std::vector<std::vector<int>> *branchPairs = nullptr, const std::vector<mrx_tree *> *precTrees = nullptr,
                           int unitCell = 0, bool *topDownFolded = nullptr) {
    cudaStream_t st = stream();
    const double tEnter = now_ms();
    Operator &op = oper.op;
    const int M = op.size(), DM = oper.dev.DM;
    const bool deriv = derivDir >= 0;
    const int world = comm_world(comm), rank = comm_rank(comm);
    const int shardB = world > 1 ? shard_block() : 1;
    const char *uEnv = getenv("MRX_UNIT_TUPLES");
    const int unitTuples = uEnv ? std::max(8, atoi(uEnv)) : 64; // tuples per contraction work unit
    const bool profile = getenv("MRX_PROFILE") != nullptr;
    const bool useMailbox = getenv("MRX_NO_MAILBOX") == nullptr;
    Mailbox *mb = mailbox();
    // result streamed into the output tree's pinned host chunks while the loop runs (one GPU, convolution apply)
    const bool mirror = out.hostMirror && derivDir < 0 && out.host.coefsPinned() && !getenv("MRX_NO_MIRROR_STREAM");
    // mirrored output, one GPU: iterations of at least 2 x subRangeMinTiles x 256 nodes run in up to subRangesMax sub-ranges
    const int subRangesMax = getenv("MRX_SUB_RANGES") ? atoi(getenv("MRX_SUB_RANGES")) : 4;
    const int subRangeMinTiles = getenv("MRX_SUB_MIN_TILES") ? std::max(1, atoi(getenv("MRX_SUB_MIN_TILES"))) : 4;
    MirrorStream *ms_ = mirror ? &mirror_stream() : nullptr;
    // shared mirror (mrx_tree_set_shared_host_mirror): host chunk c is downloaded by rank c % world over that rank's PCIe link
    const int shareW = (mirror && out.mirrorComm && out.mirrorComm == comm) ? world : 1, shareR = rank;
    // TopDown(+=) inside the loop: only when every branch node of the output comes from this apply's own splits (bare roots)
    const bool fold = topDownFolded != nullptr && branchPairs != nullptr && derivDir < 0 && !getenv("MRX_NO_TDFOLD");
    TopDownStream *td_ = fold ? &topdown_stream() : nullptr;
    if (topDownFolded) *topDownFolded = fold;
    int tdCount = 0, tdBuf = 0; // pairs of the step that becomes runnable when the CURRENT iteration's nodes are in the store
    Tree<3> &g = out.host;
    Tree<3> &f = inp.host;
    const int K = g.K, Kd = g.Kd, ncoef = g.ncoef;
    const int maxScale = g.mra.maxScale();
    g.allocCoefs = false; // the output lives in HBM until somebody asks for it
    out.hostCoefsValid = false;
    out.devValid = true;
    const double *filt = device_filters(g.k);
    const int fRealN = f.nReal;

    Scratch scr;
    scr.counters.reserve(4, false, st);
    MRX_CUDA(cudaMemsetAsync(scr.counters.p, 0, 4 * sizeof(unsigned long long), st));
    BandTables &bt = get_band_tables(oper, prec, derivDir, st);
    const std::vector<int> &bsf = bt.bsf;
    cudaEvent_t ev0, ev1, ev2, ev3;
    MRX_CUDA(cudaEventCreate(&ev0));
    MRX_CUDA(cudaEventCreate(&ev1));
    MRX_CUDA(cudaEventCreate(&ev2));
    MRX_CUDA(cudaEventCreate(&ev3));
    float kernel_ms = 0.f, contract_ms = 0.f;
    long long tuplesTotal = 0;

    if (inp.dev.topoNodes != fRealN) S.h2d_bytes += 16ll * fRealN; // child pointer, depth, norm bound per node
    ensure_input_topology(inp, st);
    DeviceTree &fd = inp.dev;
    const double fMaxNorm = fd.topoMaxNorm;
    DevTopo topo;
    int fTotal = fRealN; // real + generated input nodes known to the device
    topo.reserve((size_t)fRealN + 4096, st);
    MRX_CUDA(cudaMemcpyAsync(topo.child0.p, fd.topoChild0.p, sizeof(int) * fRealN, cudaMemcpyDeviceToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(topo.depth.p, fd.topoDepth.p, sizeof(int) * fRealN, cudaMemcpyDeviceToDevice, st));
    MRX_CUDA(cudaMemcpyAsync(topo.bound.p, fd.topoBound.p, sizeof(double) * fRealN, cudaMemcpyDeviceToDevice, st));
    MRX_CUDA(cudaMemsetAsync(topo.flag.p, 0, sizeof(int) * topo.flag.cap, st));

    // input tree in pinned host memory (DeviceTree::partial): nodes are gathered over PCIe when the apply first reads them
    const bool lazy = inp.dev.partial;
    // gather of sub-range s + 1 beside the contraction of sub-range s (k = 7 kernel: the one with unit sub-ranges)
    const bool overlapFetch = lazy && K == 8 && !getenv("MRX_NO_FETCH_OVERLAP");
    FetchStream *fs_ = overlapFetch ? &fetch_stream() : nullptr;
    if (lazy) {
        scr.fetchCnt.reserve(1, false, st);
        scr.fetchSnap.reserve(kMaxSubRanges + 1, false, st);
        scr.fetchTotal.reserve(1, false, st);
        MRX_CUDA(cudaMemsetAsync(scr.fetchSnap.p, 0, (kMaxSubRanges + 1) * sizeof(int), st));
        MRX_CUDA(cudaMemsetAsync(scr.fetchTotal.p, 0, sizeof(unsigned long long), st));
    }
    // sharded apply: the iteration whose rows are still travelling / not yet unpacked into the node store. The unpack runs on a
    // side stream beside the next iteration's kernels; whatever it reads or writes (ring slot `buf` of the staging buffer, the
    // norm rows and the slot list; the node store) is protected by an event the main stream waits for before reusing it.
    struct {
        bool active = false;
        int buf = 0, nG = 0, rows = 0;
        int slot0 = -1; // first slot of the iteration's (contiguous) nodes; -1: the first work vector (slots of the caller's grid)
        int tdCount = 0, tdBuf = 0; // TopDown(+=) step whose children are this iteration's nodes
    } pend;
    // host mirror: the nodes of one iteration, whose wavelet blocks are final once the iteration is in the node store (the closing
    // TopDown(+=) touches scaling blocks, BottomUp rewrites branch nodes only), on the download stream, per 64-node host chunk;
    // `after` = event on another stream the copies have to wait for
    auto mirror_copy = [&](cudaEvent_t after, int slot0, int cnt, int storeNodes) {
        out.host.ensureCoefStorageFor((size_t)storeNodes);
        MRX_CUDA(cudaStreamWaitEvent(ms_->dl, after, 0));
        // whole nodes, contiguous per 64-node host chunk: one large DMA transfer each (strided 7-of-8-block copies cost a DMA
        // descriptor per 28 KB row and ran at a third of the PCIe rate); the scaling block that travels with them is not final
        // yet and is overwritten by the closing push
        auto copy_run = [&](int s0, int c) {
            while (c > 0) {
                const int inChunk = std::min(c, 64 - (s0 & 63));
                if (shareW == 1 || (s0 >> 6) % shareW == shareR)
                    MRX_CUDA(cudaMemcpyAsync(out.host.coef(s0), out.dev.coefs.p + (size_t)s0 * out.host.ncoef,
                                             (size_t)inChunk * out.host.ncoef * sizeof(double), cudaMemcpyDeviceToHost, ms_->dl));
                s0 += inChunk;
                c -= inChunk;
            }
        };
        if (slot0 < 0) { // first work vector: runs of consecutive slots
            size_t a = 0;
            while (a < workVec.size()) {
                size_t b2 = a + 1;
                while (b2 < workVec.size() && workVec[b2] == workVec[b2 - 1] + 1) b2++;
                copy_run(workVec[a], (int)(b2 - a));
                a = b2;
            }
        } else {
            copy_run(slot0, cnt);
        }
        MRX_CUDA(cudaEventRecord(ms_->evCopied, ms_->dl));
        ms_->pending = true;
    };
    // TopDown(+=) step whose children are the nodes of the iteration that just reached the node store; `after` = the event that
    // says so. The mirror copies of those nodes follow it on the download stream (their scaling blocks are final then).
    auto topdown_step = [&](cudaEvent_t after, int cnt, int buf, int pairOff = 0) {
        if (cnt <= 0) return after;
        MRX_CUDA(cudaStreamWaitEvent(td_->td, after, 0));
        launch_transform(true, false, out.dev.coefs.p, scr.tdPairs[buf].p + 2 * (size_t)pairOff, cnt, K, filt, td_->td,
                         transform_fuses_norms(K) ? out.dev.norms.p : nullptr);
        MRX_CUDA(cudaEventRecord(td_->evDone, td_->td));
        MRX_CUDA(cudaEventRecord(td_->evBuf[buf], td_->td));
        td_->pending = true;
        td_->bufBusy[buf] = true;
        return td_->evDone;
    };
    bool unpackInFlight[kCommStageBufs] = {false, false, false};
    cudaStream_t ust = (world > 1) ? comm_unpack_stream(comm) : nullptr;
    auto wait_unpack = [&](int slot) { // main stream: the unpack that used ring slot `slot` has finished
        if (world > 1 && unpackInFlight[slot]) {
            MRX_CUDA(cudaStreamWaitEvent(st, comm_ev_unpacked(comm, slot), 0));
            unpackInFlight[slot] = false;
        }
    };
    auto wait_all_unpacks = [&]() {
        for (int x = 0; x < kCommStageBufs; x++) wait_unpack(x);
    };
    auto unpack_async = [&]() { // rows of `pend` have landed (the caller has ordered that on the main stream)
        MRX_CUDA(cudaEventRecord(comm_ev_gathered(comm), st));
        MRX_CUDA(cudaStreamWaitEvent(ust, comm_ev_gathered(comm), 0));
        launch_unpack_nodes(out.dev.coefs.p, reinterpret_cast<double *>(comm_stage(comm, pend.buf)), scr.gslotsAll[pend.buf].p, pend.nG,
                            world, pend.rows, shardB, ncoef, scr.normsW[pend.buf].p, out.dev.norms.p, ust);
        MRX_CUDA(cudaEventRecord(comm_ev_unpacked(comm, pend.buf), ust));
        unpackInFlight[pend.buf] = true;
        cudaEvent_t ready = comm_ev_unpacked(comm, pend.buf);
        if (fold) ready = topdown_step(ready, pend.tdCount, pend.tdBuf);
        if (mirror) mirror_copy(ready, pend.slot0, pend.nG, out.dev.nNodes);
    };
    auto flush_pending = [&]() {
        if (pend.active) {
            // own push done -> tiny all-reduce: behind it every peer's push is done as well -> unpack
            MRX_CUDA(cudaStreamWaitEvent(st, comm_ev_pushed(comm, pend.buf), 0));
            comm_allreduce_sum(comm, reinterpret_cast<double *>(scr.counters.p + 2), 1, st);
            unpack_async();
            pend.active = false;
        }
        wait_all_unpacks();
    };

    // ---- precision trees (apply.cpp:214-234): per real node the value getMaxSquareNorm() answers after makeMaxSquareNorms
    //      (stored maximum over the node and its descendants if positive, else the node's own scaled square norm), computed from
    //      the host's norms (children have larger slots than their parent) and uploaded; node stores and topologies resident
    const bool usePrec = precTrees != nullptr;
    PrecParams PR{};
    if (usePrec) {
        std::vector<PrecTreeDev> tab;
        for (mrx_tree *pt : *precTrees) {
            Tree<3> &t = pt->host;
            if (!(t.mra == g.mra)) MRX_ABORT("Incompatible MRA");
            ensure_input_topology(*pt, st);
            std::vector<double> v(t.nReal);
            for (int n = t.nReal - 1; n >= 0; n--) {
                const double own = std::pow(2.0, 3 * t.nodes[n].scale) * t.sqn[n];
                double mx = own;
                if (t.isBranch(n) && t.nodes[n].child0 < t.nReal)
                    for (int c = 0; c < 8; c++) mx = std::max(mx, v[t.nodes[n].child0 + c]);
                v[n] = mx;
            }
            // getMaxSquareNorm (MWNode.h:84): a stored maximum that is not positive falls back to the node's own scaled norm,
            // which is then 0 as well: v[n] as computed covers both cases
            auto buf = std::make_unique<DevBuf<double>>();
            buf->reserve(std::max(t.nReal, 1), false, st);
            MRX_CUDA(cudaMemcpyAsync(buf->p, v.data(), sizeof(double) * t.nReal, cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaStreamSynchronize(st)); // v is a local
            tab.push_back(PrecTreeDev{pt->dev.topoChild0.p, pt->dev.coefs.p, buf->p});
            scr.precVReal.push_back(std::move(buf));
        }
        scr.precTreeTab.reserve(std::max<size_t>(tab.size(), 1), false, st);
        if (!tab.empty()) {
            MRX_CUDA(cudaMemcpyAsync(scr.precTreeTab.p, tab.data(), sizeof(PrecTreeDev) * tab.size(), cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaStreamSynchronize(st));
        }
        PR.depthShift = op.operRoot - g.mra.rootScale;
        PR.trees = scr.precTreeTab.p;
        PR.nTrees = (int)tab.size();
        for (int x = 0; x < 3; x++) {
            PR.corner[x] = g.mra.corner[x];
            PR.nboxes[x] = g.mra.nboxes[x];
        }
        PR.K = K;
        PR.rootScale = g.mra.rootScale;
        PR.filters = filt;
    }

    // ---- first work vector: from the host topology
    int nG = (int)workVec.size();
    int minDep = 1 << 30, maxDep = -(1 << 30);
    {
        std::vector<int4> gN(std::max(nG, 1));
        std::vector<unsigned char> br(std::max(nG, 1), 0);
        bool anyBranch = false;
        for (int i = 0; i < nG; i++) {
            const auto &nd = g.nodes[workVec[i]];
            const int dep = nd.scale - op.operRoot;
            gN[i] = make_int4(dep, nd.l[0], nd.l[1], nd.l[2]);
            br[i] = g.isBranch(workVec[i]) ? 1 : 0;
            anyBranch = anyBranch || br[i];
            minDep = std::min(minDep, dep);
            maxDep = std::max(maxDep, dep);
            if (dep >= 0 && dep < DM && !bt.built[dep]) bt.build(op, dep, derivDir, bsf, DM);
        }
        scr.gAll[0].reserve(std::max(nG, 1), false, st);
        scr.gslotsAll[0].reserve(std::max(nG, 1), false, st);
        scr.isBranch.reserve(std::max(nG, 1), false, st);
        if (nG > 0) {
            MRX_CUDA(cudaMemcpyAsync(scr.gAll[0].p, gN.data(), sizeof(int4) * nG, cudaMemcpyHostToDevice, st));
            MRX_CUDA(cudaMemcpyAsync(scr.gslotsAll[0].p, workVec.data(), sizeof(int) * nG, cudaMemcpyHostToDevice, st));
            if (anyBranch) MRX_CUDA(cudaMemcpyAsync(scr.isBranch.p, br.data(), nG, cudaMemcpyHostToDevice, st));
        }
        scr.hasBranchFlags = anyBranch;
        MRX_CUDA(cudaStreamSynchronize(st)); // gN / br are locals
    }
    upload_band_tables(bt, DM, true, st);
    // split_check's scale factors as the host's pow() gives them; TreeBuilder state
    {
        const int nS = g.mra.maxDepth + 3;
        std::vector<double> sf(nS);
        for (int d = 0; d < nS; d++) {
            const double expo = 0.5 * 1.0 * (g.mra.rootScale + d + 1);
            sf[d] = std::pow(2.0, -expo);
        }
        scr.scaleFac.reserve(nS, false, st);
        scr.state.reserve(4, false, st);
        scr.splitRes.reserve(1, false, st);
        MRX_CUDA(cudaMemcpyAsync(scr.scaleFac.p, sf.data(), sizeof(double) * nS, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemsetAsync(scr.state.p, 0, 4 * sizeof(double), st));
        MRX_CUDA(cudaMemsetAsync(scr.splitRes.p, 0, sizeof(SplitResult), st));
        MRX_CUDA(cudaStreamSynchronize(st));
    }
    auto prep_local = [&](int cur, int slotBuf, int nGgiven) {
        const int cap = (nGgiven >= 0) ? nGgiven : 8 * nG; // upper bound of the next work vector
        const int capL = shard_rows(cap, world, shardB) + 1;
        scr.gNodes.reserve(capL, false, st);
        scr.gslots.reserve(capL, false, st);
        scr.chunkOff.reserve(capL + 1, false, st);
        PrepParams PP{};
        if (usePrec) {
            // precision factors of the work vector in buffer `cur` (item count on the device when the vector was just built)
            scr.precAll[cur].reserve(std::max(cap, 1), false, st);
            scr.precLoc.reserve(capL, false, st);
            PR.gNodesAll = scr.gAll[cur].p;
            PR.nG = nGgiven;
            PR.nGptr = (nGgiven >= 0) ? nullptr : &scr.splitRes.p->nNext;
            PR.precFacAll = scr.precAll[cur].p;
            launch_prec_factor(PR, cap, st);
            PP.precAll = scr.precAll[cur].p;
            PP.precLoc = scr.precLoc.p;
        }
        PP.nG = nGgiven;
        PP.world = world;
        PP.rank = rank;
        PP.shardB = shardB;
        PP.gNodesAll = scr.gAll[cur].p;
        PP.slotsAll = scr.gslotsAll[slotBuf].p;
        PP.offCount = bt.d_offCount.p;
        PP.depthInfo = bt.d_info.p;
        PP.DM = DM;
        PP.gNodesLoc = scr.gNodes.p;
        PP.slotsLoc = scr.gslots.p;
        PP.chunkOffLoc = scr.chunkOff.p;
        PP.res = scr.splitRes.p;
        launch_prep_local(PP, st);
    };
    SplitResult res{};
    prep_local(0, 0, nG);
    read_back(&res, scr.splitRes.p, &mb->res, &mb->flagRes, useMailbox, st);

    // host replay of the split decisions (deferred: runs while the device contracts the next iteration)
    // split flags: split_kernel stores them straight into a pinned, mapped host arena (no copy-engine transfer in the loop: a small
    // device-to-host copy would queue behind the megabytes a mirrored output streams down on the same DMA engine and stall the
    // main stream); the host reads them one iteration later, after it has seen a later mailbox flag (the kernel has completed)
    struct Replay {
        size_t off;   // offset of the flags in the pinned arena
        int n;        // items of the iteration
        bool split;   // false: the iteration took no split decisions (all flags zero)
        int slotBase;
    };
    static unsigned char *flagArena = nullptr;
    static size_t flagArenaCap = 0;
    size_t flagArenaUsed = 0;
    auto flag_arena_reserve = [&](size_t need) {
        if (need <= flagArenaCap) return;
        MRX_CUDA(cudaStreamSynchronize(st)); // copies into the old arena have landed; pending replays still read it: copy over
        size_t ncap = std::max<size_t>(need * 2, (size_t)1 << 20);
        unsigned char *np = nullptr;
        MRX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&np), ncap, cudaHostAllocMapped | cudaHostAllocPortable));
        if (flagArena) {
            std::memcpy(np, flagArena, flagArenaUsed);
            MRX_CUDA(cudaFreeHost(flagArena));
        }
        flagArena = np;
        flagArenaCap = ncap;
    };
    std::vector<Replay> replay; // one per finished iteration, processed in order
    size_t replayed = 0;
    std::vector<int> hostVec = workVec; // host work vector of iteration `replayed`
    auto replay_pending = [&]() {
        for (; replayed < replay.size(); replayed++) {
            const Replay &R = replay[replayed];
            const unsigned char *rflags = flagArena + R.off;
            std::vector<int> next;
            int expect = R.slotBase;
            for (size_t i = 0; i < hostVec.size(); i++) {
                g.nodes[hostVec[i]].flags |= FlagHasCoefs;
                if (R.split && (int)i < R.n && rflags[i]) {
                    const int c0 = g.createChildren(hostVec[i], false);
                    if (c0 != expect) MRX_ABORT("apply: host replay of the device split decisions lost its slot order");
                    expect += 8;
                    if (branchPairs) { // level lists of the closing transforms, for free while the device is busy
                        const size_t d = (size_t)(g.nodes[hostVec[i]].scale - g.mra.rootScale);
                        if (branchPairs->size() <= d) branchPairs->resize(d + 1);
                        (*branchPairs)[d].push_back(hostVec[i]);
                        (*branchPairs)[d].push_back(c0);
                    }
                    for (int c = 0; c < 8; c++) next.push_back(c0 + c);
                }
            }
            hostVec.swap(next);
        }
    };

    int nRealDev = g.nReal; // slots of the output tree as the device sees them (the host topology lags by one iteration)
    int iter = 0, cur = 0;
    double tp_gen = 0, tp_wait = 0, tp_split = 0, tp_replay = 0;
    const double tLoop = now_ms();
    while (nG > 0) {
        double tq = now_ms();
        const int b = iter % kCommStageBufs;
        const int nL = res.nLoc, nChunks = res.nChunksLoc;
        const long long nbrCap = res.nbrCapLoc;
        const int rowsPerRank = shard_rows(nG, world, shardB);
        if (nbrCap >= (1ll << 31) || (long long)nChunks * 32 >= (1ll << 31)) MRX_ABORT("apply: band of one iteration exceeds 2^31 entries");
        double gThrs = g.squareNorm; // ConvolutionCalculator.cpp:241-248
        if (gThrs > 0.0) gThrs = prec * 1.0 * std::sqrt(gThrs / static_cast<double>(M));
        const bool screenOn = !deriv && gThrs >= 0.0;
        // ---- band enumeration + generated input nodes (apply_enum.cu)
        scr.gdesc.reserve(std::max(nL, 1), false, st);
        scr.nbr.reserve(std::max<long long>(nbrCap, 1), false, st);
        scr.pending.reserve(std::max<long long>(nbrCap, 1), false, st);
        scr.ecnt.reserve(1, false, st);
        scr.pNode.reserve(std::max<size_t>((size_t)nChunks * 32, 1), false, st);
        scr.chunkPacked.reserve(std::max(nChunks, 1), false, st);
        scr.chunkScan.reserve(std::max(nChunks, 1), false, st);
        scr.tileTotal.reserve(nChunks / 2048 + 2, false, st);
        scr.tileBase.reserve(nChunks / 2048 + 3, false, st);
        MRX_CUDA(cudaMemsetAsync(scr.ecnt.p, 0, sizeof(EnumCounters), st));
        EnumParams E{};
        E.gNodes = scr.gNodes.p;
        E.gSlots = scr.gslots.p;
        E.nG = nL;
        E.chunkOff = scr.chunkOff.p;
        E.nChunks = nChunks;
        E.pNode = scr.pNode.p;
        E.chunkPacked = scr.chunkPacked.p;
        E.chunkScan = scr.chunkScan.p;
        E.tileTotal = scr.tileTotal.p;
        E.tileBase = scr.tileBase.p;
        E.depthShift = op.operRoot - f.mra.rootScale;
        E.offStart = bt.d_offStart.p;
        E.offCount = bt.d_offCount.p;
        E.offs = bt.d_offs.p;
        E.depthInfo = bt.d_info.p;
        E.candOff = bt.d_candOff.p;
        E.DM = DM;
        E.fChild0 = topo.child0.p;
        E.fDepth = topo.depth.p;
        E.fBound = topo.bound.p;
        E.fFlag = topo.flag.p;
        for (int x = 0; x < 3; x++) {
            E.corner[x] = f.mra.corner[x];
            E.nboxes[x] = f.mra.nboxes[x];
        }
        E.gThrs = gThrs;
        E.fMaxNorm = fMaxNorm;
        E.screenOn = screenOn ? 1 : 0;
        E.periodic = f.mra.periodic ? 1 : 0;
        E.reach = op.operReach;
        E.unitCell = unitCell;
        // apply with precTrees: gThrs = prec * precFac(node) * sqrt(|g|^2 / M) (ConvolutionCalculator.cpp:241-248)
        const double sqrtTerm = (g.squareNorm > 0.0) ? std::sqrt(g.squareNorm / static_cast<double>(M)) : g.squareNorm;
        if (usePrec) {
            E.precFac = scr.precLoc.p;
            E.prec = prec;
            E.sqrtTerm = sqrtTerm;
        }
        E.gdesc = scr.gdesc.p;
        E.nbr = scr.nbr.p;
        E.pending = scr.pending.p;
        E.cnt = scr.ecnt.p;
        launch_enum(E, st);
        EnumCounters ec;
        read_back(&ec, scr.ecnt.p, &mb->ec, &mb->flagEc, useMailbox, st);
        const int nNbr = ec.nNbr;
        const long long nCand = (long long)ec.nCand;
        if (nCand >= (1ll << 31)) MRX_ABORT("apply: candidate space of one iteration exceeds 2^31");
        const int nPending = ec.nPending;
        while (nPending > 0) { // generated input nodes, one round per missing level
            scr.newParents.reserve(nPending, false, st);
            scr.genItems.reserve((size_t)2 * nPending, false, st);
            E.newParents = scr.newParents.p;
            E.genItems = scr.genItems.p;
            MRX_CUDA(cudaMemsetAsync(&scr.ecnt.p->nUnresolved, 0, 2 * sizeof(int), st)); // nUnresolved, nNewParents
            launch_enum_resolve(E, nPending, st);
            read_back(&ec, scr.ecnt.p, &mb->ec, &mb->flagEc, useMailbox, st);
            if (ec.nUnresolved == 0) break;
            const int nNew = ec.nNewParents;
            topo.reserve((size_t)fTotal + 8 * (size_t)nNew, st);
            E.fChild0 = topo.child0.p;
            E.fDepth = topo.depth.p;
            E.fBound = topo.bound.p;
            E.fFlag = topo.flag.p;
            inp.dev.genCoefs.reserve((size_t)(fTotal - fRealN + 8 * nNew) * Kd, true, st);
            inp.dev.genNorms.reserve((size_t)(fTotal - fRealN + 8 * nNew), true, st);
            launch_enum_create(E, nNew, fTotal, st);
            if (lazy) { // real leaves that get generated children are read by the generation kernel
                scr.fetchList.reserve((size_t)8 * std::max(nNew, 1), false, st);
                MRX_CUDA(cudaMemsetAsync(scr.fetchCnt.p, 0, sizeof(int), st));
                launch_fetch_mark(scr.newParents.p, nNew, fRealN, inp.dev.resident.p, scr.fetchList.p, scr.fetchCnt.p, st);
                launch_fetch_nodes(inp.dev.coefs.p, inp.dev.chunkTab.p, scr.fetchList.p, nullptr, scr.fetchCnt.p, ncoef, scr.fetchTotal.p,
                                   inp.dev.resident.p, st);
            }
            launch_gen_children(inp.dev.coefs.p, inp.dev.genCoefs.p, inp.dev.genNorms.p, fRealN, scr.genItems.p, nNew, K, filt, st);
            fTotal += 8 * nNew;
            S.gen_nodes += 8 * (long long)nNew;
            inp.dev.nGen = fTotal - fRealN;
        }
        tp_gen += now_ms() - tq;
        tq = now_ms();
        // ---- device storage for the output nodes of this iteration (a growing store is copied: nothing may be writing it)
        if ((size_t)nRealDev * ncoef > out.dev.coefs.cap || (size_t)nRealDev * 8 > out.dev.norms.cap) {
            wait_all_unpacks();
            if (fold && td_->pending) { // the TopDown stream still updates the store that is about to be replaced
                MRX_CUDA(cudaStreamWaitEvent(st, td_->evDone, 0));
                td_->pending = false;
            }
            if (mirror && ms_->pending) { // the download stream still reads the store that is about to be replaced
                MRX_CUDA(cudaStreamWaitEvent(st, ms_->evCopied, 0));
                ms_->pending = false;
            }
            // room for the whole tree if an earlier apply of this process told how large it gets: no copy, no regrowth
            const size_t want = std::max<size_t>((size_t)nRealDev, std::min<size_t>(nodeStoreHint(), (size_t)64 * nRealDev + 4096));
            out.dev.coefs.reserve(want * ncoef, true, st);
            out.dev.norms.reserve(want * 8, true, st);
        }
        out.dev.nNodes = nRealDev;
        wait_unpack(b); // ring slot b (staging rows, norm rows) is written again by this iteration

        ApplyParams P{};
        P.fReal = inp.dev.coefs.p;
        P.fGen = inp.dev.genCoefs.p;
        P.fNorms = inp.dev.norms.p;
        P.fGenNorms = inp.dev.genNorms.p;
        P.nRealF = fRealN;
        P.gCoefs = out.dev.coefs.p;
        P.gdesc = scr.gdesc.p; // node-level descriptors: balancing happens on the device
        P.partials = nullptr;
        P.nbr = scr.nbr.p;
        P.mats = oper.dev.mats.p;
        P.onorms = oper.dev.norms.p;
        P.nodeBase = oper.dev.nodeBase.p;
        P.bsf = oper.dev.bsf.p;
        P.bsfSep = reinterpret_cast<const int4 *>(oper.dev.bw.p);
        P.depthInfo = bt.d_info.p;
        P.candOff = bt.d_candOff.p;
        P.candTerm = bt.d_candTerm.p;
        P.candMask = bt.d_candMask.p;
        P.M = M;
        P.DM = DM;
        P.K = K;
        P.gThrs = gThrs;
        if (usePrec) {
            P.precFac = scr.precLoc.p;
            P.prec = prec;
            P.sqrtTerm = sqrtTerm;
        }
        P.counters = scr.counters.p;
        P.derivDir = derivDir;
        P.identIdx = oper.dev.identIdx;

        MRX_CUDA(cudaEventRecord(ev0, st));
        scr.masks.reserve(std::max<long long>(nCand, 1), false, st);
        scr.cnt64.reserve(std::max<size_t>((size_t)nNbr * 64, 1), false, st);
        scr.segOff.reserve(std::max<size_t>((size_t)nNbr * 64, 1), false, st);
        scr.blockCnt.reserve((size_t)nL * 8 + 8, false, st);
        scr.blockTupOff.reserve((size_t)nL * 8 + 9, false, st);
        scr.blockUnitOff.reserve((size_t)nL * 8 + 9, false, st);
        DevBuf<double> &normsBuf = scr.normsW[world > 1 ? b : 0];
        normsBuf.reserve((size_t)world * rowsPerRank * 8 + 8, false, st);
        scr.header.reserve(1, false, st);
        scr.queue.reserve(1, false, st);
        PipeBuffers B{};
        B.masks = scr.masks.p;
        B.cnt64 = scr.cnt64.p;
        B.segOff = scr.segOff.p;
        B.blockCnt = scr.blockCnt.p;
        B.blockTupOff = scr.blockTupOff.p;
        B.blockUnitOff = scr.blockUnitOff.p;
        scr.pTileTotal.reserve((size_t)nL * 8 / 2048 + 2, false, st);
        scr.pTileBaseTup.reserve((size_t)nL * 8 / 2048 + 2, false, st);
        scr.pTileBaseUnit.reserve((size_t)nL * 8 / 2048 + 2, false, st);
        B.tileTotal = scr.pTileTotal.p;
        B.tileBaseTup = scr.pTileBaseTup.p;
        B.tileBaseUnit = scr.pTileBaseUnit.p;
        B.header = scr.header.p;
        B.queue = scr.queue.p;
        // A large iteration runs in node sub-ranges (k = 7 kernel), boundaries on tiles of the block scan (256 nodes); same units,
        // same unit order: bit-identical. Two uses. Host mirror on one GPU: contraction, reduce, TopDown step and download per
        // sub-range, so that an iteration's first nodes travel while its last ones are still contracted (the two largest
        // iterations come late: their download was the tail of the call). Input gathered from host memory: fill pass and gather
        // per sub-range, the gather of sub-range s + 1 on its own stream beside the contraction of sub-range s; the first
        // sub-range, whose gather nothing hides, is a small one. Work vectors after the first hold the children of the nodes
        // that split, 8 per (parent, child0) pair.
        const bool subMirror = mirror && world == 1 && K == 8;
        int nSub = 1, subNode[kMaxSubRanges + 1] = {0};
        if ((subMirror || overlapFetch) && iter > 0 && subRangesMax > 1) {
            const int nTiles = (nL * 8 + kSubRangeBlocks - 1) / kSubRangeBlocks;
            const int S = std::max(1, std::min(std::min(subRangesMax, kMaxSubRanges - 1), nTiles / subRangeMinTiles));
            if (S > 1) {
                const int first = overlapFetch ? std::max(1, nTiles / 16) : 0; // tiles of the small leading sub-range
                nSub = 0;
                if (first > 0) B.subTile[nSub++] = 0;
                for (int x = 0; x < S; x++) B.subTile[nSub++] = first + (int)((long long)x * (nTiles - first) / S);
                for (int x = 0; x < nSub; x++) subNode[x] = B.subTile[x] * (kSubRangeBlocks / 8);
            }
        }
        subNode[nSub] = nL;
        B.nSub = nSub;
        launch_pipe_screen(P, B, nNbr, st);
        launch_pipe_scan(P, B, nL, unitTuples, st);
        PipeHeader hdr;
        read_back(&hdr, scr.header.p, &mb->hdr, &mb->flagHdr, useMailbox, st);
        if (hdr.totalTuples >= (1ull << 32)) MRX_ABORT("apply: tuple list of one iteration exceeds 2^32 records");
        scr.tuples.reserve(std::max<size_t>((size_t)hdr.totalTuples, 1), false, st);
        scr.units2.reserve(std::max<size_t>((size_t)hdr.nUnits, 1), false, st);
        scr.partials.reserve(std::max<size_t>((size_t)hdr.nUnits * Kd, 1), false, st);
        B.tuples = scr.tuples.p;
        B.units = scr.units2.p;
        B.partials = scr.partials.p;
        if (lazy) {
            scr.fetchList.reserve((size_t)8 * std::max(nNbr, 1), false, st);
            MRX_CUDA(cudaMemsetAsync(scr.fetchCnt.p, 0, sizeof(int), st));
            B.resident = inp.dev.resident.p;
            B.fetchList = scr.fetchList.p;
            B.fetchCnt = scr.fetchCnt.p;
        }
        launch_pipe_units(B, nL, st);
        const bool subFetch = overlapFetch && nSub > 1;
        if (subFetch) {
            // all fill passes up front (each leaves the queue length behind it in fetchSnap[x + 1]; fetchSnap[0] stays 0), the
            // gathers behind them on the gather stream; the contraction of sub-range x below waits for gather x only
            for (int x = 0; x < nSub; x++) {
                B.fillLo = subNode[x];
                B.fillHi = subNode[x + 1];
                launch_pipe_fill(P, B, nNbr, st);
                MRX_CUDA(cudaMemcpyAsync(scr.fetchSnap.p + x + 1, scr.fetchCnt.p, sizeof(int), cudaMemcpyDeviceToDevice, st));
                MRX_CUDA(cudaEventRecord(fs_->evFilled[x], st));
                MRX_CUDA(cudaStreamWaitEvent(fs_->s, fs_->evFilled[x], 0));
                launch_fetch_nodes(inp.dev.coefs.p, inp.dev.chunkTab.p, scr.fetchList.p, scr.fetchSnap.p + x, scr.fetchSnap.p + x + 1, ncoef,
                                   scr.fetchTotal.p, inp.dev.resident.p, fs_->s, fs_->grid);
                MRX_CUDA(cudaEventRecord(fs_->evFetched[x], fs_->s));
            }
        } else {
            B.fillLo = 0;
            B.fillHi = std::max(nL, 0) + 1;
            launch_pipe_fill(P, B, nNbr, st);
            if (lazy)
                launch_fetch_nodes(inp.dev.coefs.p, inp.dev.chunkTab.p, scr.fetchList.p, nullptr, scr.fetchCnt.p, ncoef, scr.fetchTotal.p,
                                   inp.dev.resident.p, st);
        }
        MRX_CUDA(cudaEventRecord(ev2, st));
        double *normsMine = normsBuf.p + (size_t)rank * rowsPerRank * 8;
        if (nSub > 1) {
            for (int x = 0; x < nSub; x++) {
                if (subFetch) MRX_CUDA(cudaStreamWaitEvent(st, fs_->evFetched[x], 0));
                const int uEnd = (x + 1 < nSub) ? hdr.subUnit[x + 1] : hdr.nUnits;
                launch_pipe_contract(P, B, uEnd, st, hdr.subUnit[x]);
                if (world > 1) continue; // sharded: one reduce into the staging rows below
                launch_pipe_reduce(P, B, scr.gslots.p, out.dev.norms.p, normsMine, subNode[x + 1], st, nullptr, subNode[x]);
                if (fold || mirror) {
                    cudaEvent_t ready = fold ? td_->evReady : ms_->evReduced;
                    MRX_CUDA(cudaEventRecord(ready, st));
                    const int cntN = subNode[x + 1] - subNode[x];
                    if (fold) ready = topdown_step(ready, std::min(tdCount - subNode[x] / 8, cntN / 8), tdBuf, subNode[x] / 8);
                    if (mirror) mirror_copy(ready, nRealDev - nG + subNode[x], cntN, nRealDev);
                }
            }
        } else {
            launch_pipe_contract(P, B, hdr.nUnits, st);
        }
        MRX_CUDA(cudaEventRecord(ev3, st));
        // partial sums in unit order + calcNorms of the output nodes (ConvolutionCalculator.cpp:270-272)
        if (world == 1 && nSub > 1) {
            MRX_CUDA(cudaEventRecord(ev1, st));
        } else if (world == 1) {
            launch_pipe_reduce(P, B, scr.gslots.p, out.dev.norms.p, normsMine, nL, st);
            MRX_CUDA(cudaEventRecord(ev1, st));
            if (fold || mirror) {
                cudaEvent_t ready = fold ? td_->evReady : ms_->evReduced;
                MRX_CUDA(cudaEventRecord(ready, st));
                if (fold) ready = topdown_step(ready, tdCount, tdBuf);
                if (mirror) mirror_copy(ready, iter == 0 ? -1 : nRealDev - nG, nG, nRealDev);
            }
        } else {
            // ---- exchange over NVLink. Output blocks: the reduce kernel writes this rank's rows of a rank-major staging
            //      buffer; copy engines push them into every peer's HBM (CUDA IPC mapping) on a second stream while
            //      the next iteration already runs, and the whole iteration is unpacked into the node store one
            //      iteration later. Norms (8 doubles per node, the input of the split decision that every rank takes
            //      identically) go through one in-place ncclAllGather, which is also the only cross-rank
            //      synchronisation: a rank enters it only after its previous push has completed, so whoever leaves it
            //      knows that the previous iteration's rows of all peers have landed.
            const size_t rowBytes = (size_t)ncoef * sizeof(double);
            const size_t segBytes = (size_t)rowsPerRank * rowBytes;
            if ((size_t)world * segBytes > comm_stage_bytes(comm)) {
                flush_pending();
                comm_stage_reserve(comm, (size_t)world * segBytes, st);
            }
            const bool push = comm_peer_push_enabled(comm);
            double *stageB = reinterpret_cast<double *>(comm_stage(comm, b));
            launch_pipe_reduce(P, B, scr.gslots.p, out.dev.norms.p, normsMine, nL, st,
                               reinterpret_cast<double *>(comm_stage(comm, b) + (size_t)rank * segBytes));
            MRX_CUDA(cudaEventRecord(ev1, st));
            if (push) {
                MRX_CUDA(cudaEventRecord(comm_ev_reduced(comm, b), st));
                comm_push(comm, b, (size_t)rank * segBytes, (size_t)nL * rowBytes);
                if (pend.active) MRX_CUDA(cudaStreamWaitEvent(st, comm_ev_pushed(comm, pend.buf), 0));
                comm_allgather(comm, normsBuf.p, (size_t)rowsPerRank * 8 * sizeof(double), st);
                if (pend.active) unpack_async(); // beside the next iteration's kernels
                pend.active = true;
                pend.buf = b;
                pend.nG = nG;
                pend.rows = rowsPerRank;
                pend.slot0 = (iter == 0) ? -1 : nRealDev - nG;
                pend.tdCount = tdCount;
                pend.tdBuf = tdBuf;
            } else {
                comm_allgather(comm, normsBuf.p, (size_t)rowsPerRank * 8 * sizeof(double), st);
                comm_allgather(comm, stageB, segBytes, st);
                launch_unpack_nodes(out.dev.coefs.p, stageB, scr.gslotsAll[b].p, nG, world, rowsPerRank, shardB, ncoef, normsBuf.p,
                                    out.dev.norms.p, st);
                if (fold || mirror) { // this iteration's nodes are in the store: its TopDown step and mirror copies
                    cudaEvent_t ready = fold ? td_->evReady : ms_->evReduced;
                    MRX_CUDA(cudaEventRecord(ready, st));
                    if (fold) ready = topdown_step(ready, tdCount, tdBuf);
                    if (mirror) mirror_copy(ready, iter == 0 ? -1 : nRealDev - nG, nG, nRealDev);
                }
            }
        }
        // ---- the device is busy for a while: replay earlier split decisions into the host topology, build the band
        //      tables of the depths the next work vector can hold
        double tr = now_ms();
        replay_pending();
        const bool doSplit = !deriv && !(iter >= maxIter && maxIter >= 0);
        if (doSplit) {
            for (int dep = std::max(minDep + 1, 0); dep <= maxDep + 1 && dep < DM; dep++)
                if (!bt.built[dep]) bt.build(op, dep, derivDir, bsf, DM);
            upload_band_tables(bt, DM, true, st); // synchronises only when a depth is new for this (operator, prec)
        }
        tp_replay += now_ms() - tr;
        // ---- TreeBuilder bookkeeping, split decisions and the next work vector on the device (apply_split.cu)
        const int nb = (iter + 1) % kCommStageBufs;
        wait_unpack(nb); // the slot list of ring slot nb is rewritten by the split kernel
        scr.gAll[cur ^ 1].reserve((size_t)8 * nG, false, st);
        scr.gslotsAll[nb].reserve((size_t)8 * nG, false, st);
        scr.flags.reserve(nG, false, st);
        SplitParams SP{};
        SP.normRows = normsBuf.p;
        SP.nG = nG;
        SP.world = world;
        SP.rows = rowsPerRank;
        SP.shardB = shardB;
        SP.gNodesAll = scr.gAll[cur].p;
        SP.isBranch = (iter == 0 && scr.hasBranchFlags) ? scr.isBranch.p : nullptr;
        SP.operRoot = op.operRoot;
        SP.rootScale = g.mra.rootScale;
        SP.maxScale = maxScale;
        SP.scaleFac = scr.scaleFac.p;
        SP.prec = prec;
        SP.precFacAll = usePrec ? scr.precAll[cur].p : nullptr;
        SP.absPrec = absPrec ? 1 : 0;
        SP.iter = iter;
        SP.doSplit = doSplit ? 1 : 0;
        SP.slotBase = nRealDev;
        SP.state = scr.state.p;
        SP.gNodesNext = scr.gAll[cur ^ 1].p;
        SP.slotsNext = scr.gslotsAll[nb].p;
        if (fold) { // (parent, first child) pairs of the nodes that split now: the TopDown step of the NEXT iteration's nodes
            const int x = (iter + 1) % 3;
            if (td_->bufBusy[x]) { // the step that read this buffer three iterations ago (long finished; the wait is free)
                MRX_CUDA(cudaStreamWaitEvent(st, td_->evBuf[x], 0));
                td_->bufBusy[x] = false;
            }
            scr.tdPairs[x].reserve((size_t)2 * std::max(nG, 1), false, st);
            SP.slotsCur = scr.gslotsAll[b].p;
            SP.pairsNext = scr.tdPairs[x].p;
        }
        // flags of this iteration: straight into the host arena (device pointer of the mapped allocation)
        flag_arena_reserve(flagArenaUsed + (size_t)nG);
        SP.flags = static_cast<unsigned char *>(mailbox_dev(flagArena)) + flagArenaUsed;
        SP.res = scr.splitRes.p;
        launch_split(SP, st);
        prep_local(cur ^ 1, nb, -1);
        Replay R;
        R.off = flagArenaUsed;
        R.n = nG;
        R.split = doSplit;
        R.slotBase = nRealDev;
        flagArenaUsed += (size_t)nG;
        read_back(&res, scr.splitRes.p, &mb->res, &mb->flagRes, useMailbox, st);
        tp_wait += now_ms() - tq;
        tq = now_ms();
        replay.push_back(R);
        float ms = 0.f, msc = 0.f;
        MRX_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        MRX_CUDA(cudaEventElapsedTime(&msc, ev2, ev3));
        kernel_ms += ms;
        contract_ms += msc;
        tuplesTotal += (long long)hdr.totalTuples;
        if (profile)
            std::fprintf(stderr, "[mrx] iter %d nG %d nbr %d cand %lld tuples %llu kernels %.3f ms (contract %.3f ms, %.2f TF/s) split %d\n", iter,
                         nG, nNbr, nCand, hdr.totalTuples, ms, msc,
                         msc > 0 ? hdr.totalTuples * 6.0 * K * K * K * K / (msc * 1e-3) / 1e12 : 0.0, res.nSplit);
        S.g_nodes += nG;
        tdCount = res.nSplit; // pairs written by this iteration's split: runnable when the next iteration's nodes exist
        tdBuf = (iter + 1) % 3;
        g.squareNorm = res.squareNorm; // TreeBuilder.cpp:56-66
        nRealDev += res.nNext;
        nG = res.nNext;
        minDep += 1;
        maxDep += 1;
        cur ^= 1;
        iter++;
        tp_split += now_ms() - tq;
    }
    if (world > 1) flush_pending();
    if (fold && td_->pending) { // the closing BottomUp (main stream) reads what the TopDown stream wrote
        MRX_CUDA(cudaStreamWaitEvent(st, td_->evDone, 0));
        td_->pending = false;
    }
    if (fold)
        for (int x = 0; x < 3; x++) td_->bufBusy[x] = false;
    MRX_CUDA(cudaStreamSynchronize(st)); // the last iteration's flags have landed; events are complete
    replay_pending();
    if (g.nReal != nRealDev) MRX_ABORT("apply: host topology and device node store disagree");
    nodeStoreHint() = (size_t)nRealDev + (size_t)nRealDev / 8;
    const double tLoopEnd = now_ms();
    if (profile) {
        std::fprintf(stderr, "[mrx] run_apply_pipe ms: pre-loop %.2f loop %.2f\n", tLoop - tEnter, tLoopEnd - tLoop);
        std::fprintf(stderr, "[mrx] host phases ms: enum+gen %.2f screen..split (device wait) %.2f of which replay/tables %.2f, bookkeeping %.2f\n",
                     tp_gen, tp_wait, tp_replay, tp_split);
    }
    S.iterations = iter;
    S.ms_kernel = kernel_ms;
    S.ms_contract = contract_ms;
    S.f_applied = tuplesTotal;
    S.f_applied_rank = S.f_applied;
    if (lazy) {
        unsigned long long fetched = 0;
        MRX_CUDA(cudaMemcpyAsync(&fetched, scr.fetchTotal.p, sizeof(fetched), cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        S.h2d_bytes += (long long)fetched * Kd * (long long)sizeof(double); // coefficient blocks gathered
    }
    if (world > 1) {
        double h[2] = {(double)S.f_applied, (double)S.gen_nodes};
        double *dsum = reinterpret_cast<double *>(scr.counters.p + 2);
        MRX_CUDA(cudaMemcpyAsync(dsum, h, sizeof(h), cudaMemcpyHostToDevice, st));
        comm_allreduce_sum(comm, dsum, 2, st);
        MRX_CUDA(cudaMemcpyAsync(h, dsum, sizeof(h), cudaMemcpyDeviceToHost, st));
        MRX_CUDA(cudaStreamSynchronize(st));
        S.f_applied = (long long)(h[0] + 0.5);
        S.gen_nodes = (long long)(h[1] + 0.5);
    }
    out.dev.nNodes = g.nReal;
    MRX_CUDA(cudaEventDestroy(ev0));
    MRX_CUDA(cudaEventDestroy(ev1));
    MRX_CUDA(cudaEventDestroy(ev2));
    MRX_CUDA(cudaEventDestroy(ev3));
}

static bool use_pipeline(const mrx_tree &out) {
    const char *legacy = getenv("MRX_LEGACY");
    return pipe_supports_order(out.host.K) && !(legacy && legacy[0] == '1');
}

void device_apply(double prec, mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int maxIter, bool absPrec,
                  mrx_apply_stats *stats, const mrx_comm *comm, const std::vector<mrx_tree *> *precTrees, int unitCell) {
    require_device("device_apply");
    if (out.host.mra.periodic) {
        // apply on a periodic world: operator rooted at the world's root scale (no nodes above the root: touchParentNodes,
        // ConvolutionCalculator.cpp:384-398, is the negative-scale variant this path does not build), with a reach
        if (oper.op.operRoot != out.host.mra.rootScale) MRX_ABORT("periodic apply: operators rooted above the world are not supported");
        if (oper.op.operReach < 0) MRX_ABORT("periodic apply: the operator was built without a reach (use the *_create_reach constructors)");
        if (!use_pipeline(out)) MRX_ABORT("periodic apply is implemented for the work-list pipeline (orders 3..11) only");
    } else if (unitCell != 0) {
        MRX_ABORT("apply_near_field / apply_far_field need a periodic world");
    }
    mrx_apply_stats S{};
    long long launches0 = launch_counter();
    double t0 = now_ms();
    // precision trees (apply.cpp:214-251) are read node by node on the device: whole trees resident (the input tree itself is
    // often one of them; it is then resident before its own residency decision below)
    if (precTrees) {
        if (!use_pipeline(out)) MRX_ABORT("apply with precision trees is implemented for the work-list pipeline (orders 3..11) only");
        for (mrx_tree *pt : *precTrees)
            if (!pt->devValid) {
                tree_upload(*pt);
                S.h2d_bytes += (long long)pt->host.nReal * (pt->host.ncoef + 8) * (long long)sizeof(double);
            }
    }
    // ---- residency: input tree + operator tables in HBM. A tree whose coefficients sit in pinned host memory is not copied
    //      as a whole: the apply gathers the nodes it reads (DeviceTree::partial); MRX_EAGER_UPLOAD=1 forces the full copy
    if (!inp.devValid) {
        const long long n8 = (long long)inp.host.nReal * 8 * (long long)sizeof(double);
        if (use_pipeline(out) && inp.hostCoefsValid && inp.host.coefsPinned() && !getenv("MRX_EAGER_UPLOAD")) {
            if (!inp.dev.partial) S.h2d_bytes += n8; // norms
            tree_lazy_begin(inp);
        } else {
            tree_upload(inp);
            S.h2d_bytes += n8 + (long long)inp.host.nReal * inp.host.ncoef * (long long)sizeof(double);
        }
    }
    oper_upload(oper);
    oper.op.calcBandWidths(prec);
    S.ms_upload = now_ms() - t0;

    double tb = now_ms();
    std::vector<int> workVec;
    out.host.nodeTable(workVec); // getInitialWorkVector: ALL nodes of `out` (ConvolutionCalculator.cpp:400-405)
    // an output tree that starts from bare roots gets all its branch nodes from this apply: their (parent, child0) pairs
    // are collected per depth during the loop and handed to the closing transforms
    std::vector<std::vector<int>> branchPairs;
    const bool bareRoots = out.host.nReal == out.host.nRoots;
    const bool pipe = use_pipeline(out);
    bool tdFolded = false;
    if (pipe) run_apply_pipe(prec, out, oper, inp, maxIter, absPrec, -1, workVec, S, const_cast<mrx_comm *>(comm), bareRoots ? &branchPairs : nullptr, precTrees, unitCell, &tdFolded);
    else {
        if (comm_world(comm) > 1) MRX_ABORT("sharded apply is implemented for the work-list pipeline (orders 3..11) only");
        run_apply_legacy(prec, out, oper, inp, maxIter, absPrec, -1, workVec, S);
    }
    S.ms_build = now_ms() - tb;

    // ---- post: TopDown(+=), BottomUp, square norm, cleanup (apply.cpp:81-87)
    double tp = now_ms();
    oper.op.clearBandWidths();
    const bool prof = getenv("MRX_PROFILE") != nullptr;
    device_apply_post(out, (pipe && bareRoots) ? &branchPairs : nullptr, pipe && bareRoots && tdFolded);
    if (pipe && out.hostMirror && out.host.coefsPinned() && !getenv("MRX_NO_MIRROR_STREAM")) {
        // ---- host mirror: what the per-iteration copies could not take -- scaling blocks of every node, all blocks of the
        //      branch nodes -- pushed by the SMs into the host chunks behind the streamed copies of the same nodes
        cudaStream_t st = stream();
        Tree<3> &g = out.host;
        MirrorStream &M = mirror_stream();
        g.ensureCoefStorageFor((size_t)g.nReal);
        // shared mirror: every rank finishes the chunks it owns; a barrier at the end makes the whole tree visible to all
        const int shareW = (out.mirrorComm && out.mirrorComm == comm) ? comm_world(comm) : 1, shareR = comm_rank(comm);
        std::vector<int> items;
        items.reserve(g.nReal / shareW + 64);
        // TopDown(+=) folded into the loop: a leaf node was copied after the step that finished its scaling block, so only the
        // branch nodes (rewritten by BottomUp) are left; otherwise the scaling block of every node has changed since its copy
        const bool leavesDone = bareRoots && tdFolded;
        for (int n = 0; n < g.nReal; n++) {
            if (shareW > 1 && (n >> 6) % shareW != shareR) continue;
            if (g.isBranch(n)) items.push_back((int)((unsigned)n | 0x80000000u));
            else if (!leavesDone) items.push_back(n);
        }
        const int nItems = (int)items.size();
        DevBuf<int> dItems;
        dItems.reserve(std::max(nItems, 1), false, st);
        const auto &chunks = g.coefChunks();
        out.dev.chunkTab.reserve(std::max<size_t>(chunks.size(), 1), false, st);
        MRX_CUDA(cudaMemcpyAsync(dItems.p, items.data(), sizeof(int) * nItems, cudaMemcpyHostToDevice, st));
        MRX_CUDA(cudaMemcpyAsync(out.dev.chunkTab.p, chunks.data(), sizeof(double *) * chunks.size(), cudaMemcpyHostToDevice, st));
        double tm0 = now_ms();
        if (prof) {
            MRX_CUDA(cudaStreamSynchronize(st));
            std::fprintf(stderr, "[mrx] mirror: closing passes drained after %.2f ms\n", now_ms() - tm0);
            tm0 = now_ms();
            MRX_CUDA(cudaStreamSynchronize(M.dl));
            std::fprintf(stderr, "[mrx] mirror: streamed copies drained after another %.2f ms\n", now_ms() - tm0);
            tm0 = now_ms();
        }
        if (M.pending) {
            MRX_CUDA(cudaStreamWaitEvent(st, M.evCopied, 0));
            M.pending = false;
        }
        launch_push_nodes(out.dev.coefs.p, const_cast<double *const *>(reinterpret_cast<const double *const *>(out.dev.chunkTab.p)), dItems.p,
                          nItems, g.ncoef, st);
        std::vector<double> offs;
        DevBuf<double> bar;
        if (shareW > 1) {
            // the peers' pushes into the shared arena are ordered before this rank's return; the same all-reduce carries every
            // rank's arena offset of the tree's first chunk: the ranks must have placed the tree at the same spot
            offs.assign(shareW, 0.0);
            offs[shareR] = (double)host_arena_offset(chunks[0]);
            bar.reserve(shareW, false, st);
            MRX_CUDA(cudaMemcpyAsync(bar.p, offs.data(), sizeof(double) * shareW, cudaMemcpyHostToDevice, st));
            comm_allreduce_sum(comm, bar.p, shareW, st);
            MRX_CUDA(cudaMemcpyAsync(offs.data(), bar.p, sizeof(double) * shareW, cudaMemcpyDeviceToHost, st));
        }
        MRX_CUDA(cudaStreamSynchronize(st));
        for (int r = 0; r < (int)offs.size(); r++)
            if (offs[r] != offs[shareR])
                MRX_ABORT("shared host mirror: the ranks placed the output tree at different arena offsets (every rank must create and "
                          "free its shared-mirror trees in the same order)");
        if (prof) std::fprintf(stderr, "[mrx] mirror: push of the remainder %.2f ms\n", now_ms() - tm0);
        for (int n = 0; n < g.nReal; n++) g.nodes[n].flags |= FlagHasCoefs;
        out.hostCoefsValid = true;
        long long mine = 0; // nodes THIS rank's link carried
        for (int n = 0; n < g.nReal; n++) mine += (shareW == 1 || (n >> 6) % shareW == shareR) ? 1 : 0;
        S.d2h_bytes = mine * g.ncoef * (long long)sizeof(double);
    }
    inp.host.deleteGenerated();
    inp.dev.nGen = 0;
    S.ms_post = now_ms() - tp;
    (void)prof;
    S.n_nodes_out = out.host.nReal;
    S.kernel_launches = launch_counter() - launches0;
    if (getenv("MRX_PROFILE"))
        std::fprintf(stderr, "[mrx] device_apply ms: upload %.2f build %.2f post %.2f total %.2f\n", S.ms_upload, S.ms_build, S.ms_post,
                     now_ms() - t0);
    if (stats) *stats = S;
}

void device_apply_derivative(mrx_tree &out, mrx_oper &oper, mrx_tree &inp, int dir, mrx_apply_stats *stats) {
    require_device("device_apply_derivative");
    mrx_apply_stats S{};
    long long launches0 = launch_counter();
    double t0 = now_ms();
    if (!inp.devValid) {
        if (use_pipeline(out) && inp.hostCoefsValid && inp.host.coefsPinned() && !getenv("MRX_EAGER_UPLOAD")) tree_lazy_begin(inp);
        else tree_upload(inp);
    }
    oper_upload(oper);
    Operator &op = oper.op;
    op.calcBandWidths(1.0); // fixed 0 or 1 for derivatives (apply.cpp:389)
    S.ms_upload = now_ms() - t0;
    double tb = now_ms();
    Tree<3> &g = out.host;
    Tree<3> &f = inp.host;
    const int maxScale = g.mra.maxScale();
    int bw[3] = {0, 0, 0};
    bw[dir] = op.getMaxBandWidth();
    g.allocCoefs = false;
    // grid: CopyAdaptor(inp, maxScale, bw) + DefaultCalculator, maxIter = -1 (apply.cpp:391-393, CopyAdaptor.cpp:56-72)
    {
        std::vector<int> workVec;
        g.endNodeTable(workVec);
        while (!workVec.empty()) {
            std::vector<int> newVec;
            for (int n : workVec) {
                if (g.isBranch(n)) continue;
                if (g.nodes[n].scale + 2 > maxScale) continue;
                // CopyAdaptor::splitNode asks, for every child c of the node, dimension d and offset |b| <= bw[d], whether the input
                // tree has a node at the child's index shifted by b along d. A node of scale + 1 exists iff its parent is a branch,
                // so the 8 x 3 x (2 bw + 1) look-ups collapse to: is one of the input nodes at THIS scale with index l, or l shifted
                // along d by j in [(2 l_d - bw) >> 1, (2 l_d + 1 + bw) >> 1], a branch node?
                bool split = false;
                const auto idx0 = g.nodes[n];
                auto branch_at = [&](const std::array<int, 3> &l) {
                    const int m = f.findNode(idx0.scale, l);
                    return m >= 0 && f.isBranch(m) && f.nodes[m].child0 < f.nReal;
                };
                split = branch_at(idx0.l);
                for (int d = 0; d < 3 && !split; d++) {
                    if (bw[d] == 0) continue;
                    const int lo = (2 * idx0.l[d] - bw[d]) >> 1, hi = (2 * idx0.l[d] + 1 + bw[d]) >> 1;
                    for (int j = lo; j <= hi && !split; j++) {
                        if (j == idx0.l[d]) continue;
                        std::array<int, 3> l = idx0.l;
                        l[d] = j;
                        split = branch_at(l);
                    }
                }
                if (split) {
                    int c0 = g.createChildren(n, false);
                    for (int c = 0; c < 8; c++) newVec.push_back(c0 + c);
                }
            }
            workVec.swap(newVec);
        }
    }
    // DerivativeCalculator on the end nodes, one pass (DerivativeCalculator.cpp:277-279, apply.cpp:396-398)
    std::vector<int> workVec;
    g.endNodeTable(workVec);
    // branch nodes are never computed: zero their device storage so BottomUp starts from defined memory
    cudaStream_t st = stream();
    out.dev.coefs.reserve((size_t)g.nReal * g.ncoef, false, st);
    out.dev.norms.reserve((size_t)g.nReal * 8, false, st);
    MRX_CUDA(cudaMemsetAsync(out.dev.coefs.p, 0, sizeof(double) * (size_t)g.nReal * g.ncoef, st));
    if (use_pipeline(out)) run_apply_pipe(-1.0, out, oper, inp, 0, false, dir, workVec, S, nullptr);
    else run_apply_legacy(-1.0, out, oper, inp, 0, false, dir, workVec, S);
    S.ms_build = now_ms() - tb;
    double tp = now_ms();
    op.clearBandWidths();
    device_mw_transform(out, MRX_BOTTOM_UP, true);
    out.host.calcSquareNorm();
    inp.host.deleteGenerated();
    inp.dev.nGen = 0;
    S.ms_post = now_ms() - tp;
    S.n_nodes_out = g.nReal;
    S.kernel_launches = launch_counter() - launches0;
    if (stats) *stats = S;
}

} // namespace mrx
