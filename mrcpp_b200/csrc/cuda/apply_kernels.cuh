// Parameter blocks of the operator-application kernels (apply_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace mrx {

/// one output node of the current work vector
struct GDesc {
    int slot;   // slot in the output tree
    int depth;  // operator depth = scale - operator root
    int nbrOff; // first entry of this node's neighbour list in `nbr`
    int nbrCnt; // number of neighbour entries
    int partial; // -1: this CTA owns the whole node and writes g directly; >= 0: slot in the partial-sum buffer
};

/// neighbour entry: input-node slot and the code of its translation offset inside the depth's band cube
struct NbrEntry {
    int fslot;
    int code;     // (dz+W)*(2W+1)^2 + (dy+W)*(2W+1) + (dx+W)
    int g;        // index of the output node in this iteration's work vector
    int candBase; // first candidate (neighbour, term) index of this entry in the iteration's candidate space
};

/// per-depth band table: for every offset of the band cube, the separation terms that can reach it and
/// the (gt,ft) combinations whose per-dimension band tests pass (integer part of applyOperator,
/// ConvolutionCalculator.cpp:311-318, hoisted out of the node loop because it only depends on depth/offset)
struct DepthInfo {
    int W;       // band_max(depth), -1 = no operator at this depth
    int cubeOff; // offset of this depth's (2W+1)^3+1 prefix array in candOff
};

struct ApplyParams {
    // input tree (real nodes: 8 blocks; generated nodes: scaling block only, slot >= nRealF)
    const double *fReal;
    const double *fGen;
    const double *fNorms;    // [nRealF][8]
    const double *fGenNorms; // [nGen]
    int nRealF;
    // output tree
    double *gCoefs;
    double *partials; // [nPartials][8][K^3]: per-chunk partial sums of nodes split across CTAs
    const GDesc *gdesc;
    const NbrEntry *nbr;
    // operator tables
    const double *mats;   // node-major: 4 blocks of K*K (column-major p[i + K m], i = input, m = output)
    const double *onorms; // 4 per node
    const int *nodeBase;  // [M][DM] node index of translation 0 (= nodeOff + maxTransl), -1 if depth absent
    const int *bsf;       // [M][DM][64]
    const int4 *bsfSep;   // [M][DM]: 1-D node counts (2 width + 1, or 1) of the components T, C, B, A: bsf = 64 * product over d
    const DepthInfo *depthInfo;         // [DM]
    const int *candOff;                 // prefix arrays, all depths
    const int *candTerm;                // candidate term ids
    const unsigned long long *candMask; // band-allowed (gt*8+ft) bits per candidate
    int M, DM, K;
    double gThrs;
    // locally scaled precision (apply with precTrees, apply.cpp:214-251): gThrs of output node g = prec * precFac[g] * sqrtTerm
    // (ConvolutionCalculator.cpp:241-248, same operation order) while the tree norm is known (sqrtTerm >= 0)
    const double *precFac; // [nG local] or nullptr
    double prec, sqrtTerm;
    unsigned long long *counters; // [0] tuples applied
    int derivDir;                 // -1 for convolution operators
    int identIdx;                 // operator block index of the K x K identity (derivative apply: dimensions != derivDir)
};

void launch_apply(const ApplyParams &P, int nG, cudaStream_t st);

// ---- band enumeration on the device (apply_enum.cu) ----------------------------------------------------
/// one reachable offset of a depth's band cube
struct OffEntry {
    int dx, dy, dz;
    int code;    // offset code inside the band cube
    double maxO; // max over (term, gt, ft) of |O|^3 * bandSizeFactor: upper bound used for the early-out
};
struct EnumCounters {
    int nNbr;
    unsigned nCand;
    int nPending;
    int nUnresolved;
    int nNewParents;
    int pad;
};
struct EnumParams {
    const int4 *gNodes; // [nG] (operator depth, lx, ly, lz) of the work-vector nodes
    const int *gSlots;  // [nG]
    int nG;
    // lane-per-offset probing: node j owns the chunks (32 offsets each) chunkOff[j] .. chunkOff[j+1]
    const int *chunkOff; // [nG+1]
    int nChunks;
    int *pNode;            // [nChunks*32] probe result: node (| pending bit) or -1
    unsigned long long *chunkPacked; // [nChunks] surviving offsets (low 32 bits) and their candidates (high 32 bits)
    unsigned long long *chunkScan;   // [nChunks] exclusive scan inside the chunk's scan tile
    unsigned long long *tileTotal;   // [nTiles]
    unsigned long long *tileBase;    // [nTiles+1] exclusive scan of the tile totals
    int depthShift; // function-tree depth = operator depth + depthShift
    const int *offStart, *offCount; // [DM] into offs
    const OffEntry *offs;
    const DepthInfo *depthInfo;
    const int *candOff;
    int DM;
    // input-tree topology (real + generated nodes, unified slot space)
    int *fChild0;
    int *fDepth;
    double *fBound; // upper bound of the node norm
    int *fFlag;
    int corner[3], nboxes[3];
    double gThrs, fMaxNorm;
    int screenOn;
    const double *precFac; // [nG] per-node precision factor (apply with precTrees) or nullptr: gThrs = prec * precFac * sqrtTerm
    double prec, sqrtTerm;
    // periodic world (ConvolutionCalculator.cpp:166-172, :191-218; periodic_utils.cpp:35-73): the band is clipped to `reach`
    // cells around the world, the input node is looked up at the index wrapped into the unit cell [-2^n, 2^n), the neighbour
    // entry keeps the unwrapped offset (it selects the operator blocks). unitCell 1 / 2: keep only indices inside / outside
    // the unit cell (apply_near_field / apply_far_field).
    int periodic, reach, unitCell;
    // outputs
    GDesc *gdesc;
    NbrEntry *nbr;
    int *pending;
    int *newParents;
    int *genItems;
    EnumCounters *cnt;
};
void launch_enum(const EnumParams &E, cudaStream_t st);
void launch_enum_resolve(const EnumParams &E, int nPending, cudaStream_t st);
void launch_enum_create(const EnumParams &E, int nNew, int firstSlot, cudaStream_t st);
/// lazy residency: flag the real nodes of `list` that are not in HBM yet and queue them; gather the queued nodes from the
/// pinned host chunks (64 nodes each) into the node store; *total accumulates the number of nodes fetched
void launch_fetch_mark(const int *list, int n, int nRealF, int *resident, int *fetchList, int *fetchCnt, cudaStream_t st);
/// host mirror of an apply output: blocks of the listed nodes from the node store straight into the pinned host chunks (64 nodes
/// each). items[i] = slot, bit 31 set: all eight blocks (branch node), clear: the scaling block only
void launch_push_nodes(const double *coefs, double *const *chunkTab, const int *items, int n, int ncoef, cudaStream_t st);
/// gathers queue entries [*begin, *end) (begin == nullptr: from 0); resident[blk] becomes 2 when block blk has arrived;
/// grid > 0: that many CTAs (a gather that runs beside a contraction kernel: one CTA per SM is co-resident), 0: the default
void launch_fetch_nodes(double *coefs, const double *const *chunkTab, const int *list, const int *begin, const int *end, int ncoef,
                        unsigned long long *total, int *resident, cudaStream_t st, int grid = 0);

// ---- work-list pipeline (apply_pipeline.cu): screen -> scan -> fill -> contract -> reduce -------------------
/// one surviving (g, f, ft, gt, term) tuple: indices of the source block and of the three 1-D operator blocks
struct TupleRec {
    int fblk; // < 8*nRealF: real block fslot*8+ft; otherwise generated scaling block (fblk - 8*nRealF)
    int o0, o1, o2; // operator block index (node*4 + component) per dimension
};
/// contraction work unit: `cnt` consecutive tuples of one output block (g, gt)
struct UnitDesc {
    unsigned t0; // first tuple in the iteration's tuple list
    int cnt;
};
constexpr int kMaxSubRanges = 8;
constexpr int kSubRangeBlocks = 2048; // sub-range boundaries fall on tiles of the block scan (256 output nodes)
struct PipeHeader {
    unsigned long long totalTuples;
    int nUnits;
    int U; // tuples per unit
    int subUnit[kMaxSubRanges]; // first unit of sub-range s (s >= nSub: nUnits)
};
struct PipeBuffers {
    unsigned long long *masks; // [nCand] surviving (gt,ft) bits per candidate
    unsigned short *cnt64;     // [nNbr][64] tuples per (neighbour, gt, ft)
    int *segOff;               // [nNbr][64] offset of the (neighbour, gt, ft) segment inside its block's tuple run
    int *blockCnt;             // [nG*8]
    unsigned *blockTupOff;     // [nG*8+1]
    int *blockUnitOff;         // [nG*8+1]
    unsigned long long *tileTotal;   // [nTiles] block-scan tiles (2048 blocks): tuples | units << 36
    unsigned long long *tileBaseTup; // [nTiles]
    int *tileBaseUnit;               // [nTiles]
    PipeHeader *header;
    TupleRec *tuples;
    UnitDesc *units;
    double *partials; // [nUnits][K^3]
    int *queue;       // dynamic unit counter of the contraction kernel
    // lazy residency of the input tree (nullptr: everything is resident): pipe_fill queues the nodes the tuples read
    int *resident;
    int *fetchList;
    int *fetchCnt;
    int fillLo, fillHi; // pipe_fill pass: nodes [fillLo, fillHi) of the rank's items
    // contraction / reduce / download of one iteration in nSub node sub-ranges: sub-range s starts at block tile subTile[s]
    int nSub;
    int subTile[kMaxSubRanges];
};
bool pipe_supports_order(int K); // orders with a work-list contraction kernel (K = k + 1)
int pipe_contract_warps(); // persistent warps of the contraction kernel on this device
void launch_pipe_screen(const ApplyParams &P, const PipeBuffers &B, int nNbr, cudaStream_t st);
void launch_pipe_scan(const ApplyParams &P, const PipeBuffers &B, int nG, int unitHint, cudaStream_t st);
void launch_pipe_units(const PipeBuffers &B, int nG, cudaStream_t st); // global block offsets + unit descriptors
/// tuple records (and the gather queue of a lazily resident input) of the nodes [B.fillLo, B.fillHi)
void launch_pipe_fill(const ApplyParams &P, const PipeBuffers &B, int nNbr, cudaStream_t st);
/// units [uBase, nUnits) of the iteration (uBase != 0: k = 7 kernel only)
void launch_pipe_contract(const ApplyParams &P, const PipeBuffers &B, int nUnits, cudaStream_t st, int uBase = 0);
/// gslots / gNormsW: the rank's nodes in local order. stageRows != nullptr: output blocks go to the rank's rows of the
/// exchange staging buffer (row j = local node j) instead of the node store; nodes [gBase, nG) of the rank's items
void launch_pipe_reduce(const ApplyParams &P, const PipeBuffers &B, const int *gslots, double *gNorms, double *gNormsW, int nG,
                        cudaStream_t st, double *stageRows = nullptr, int gBase = 0);
// ---- distribution of a work vector over the ranks of a sharded apply: block-cyclic, blocks of B consecutive items (B = 1: plain
//      cyclic). Consecutive items are siblings and spatial neighbours: they cost about the same (so dealing them out balances the
//      tuple counts) and they read nearly the same input nodes (so keeping a few together lets a rank gather fewer of them from
//      host memory). Item i -> rank (i / B) % world, local index (i / (B world)) B + i % B; rows = per-rank capacity.
__host__ __device__ inline int shard_rows(int n, int world, int B) { return (((n + B - 1) / B + world - 1) / world) * B; }
__host__ __device__ inline int shard_row(int i, int world, int rows, int B) {
    const int q = i / B;
    return (q % world) * rows + (q / world) * B + i % B;
}
__host__ __device__ inline int shard_count(int n, int world, int rank, int B) {
    const int full = n / B, rem = n % B;
    return ((full + world - 1 - rank) / world) * B + ((rem && full % world == rank) ? rem : 0);
}
__host__ __device__ inline int shard_item(int rank, int j, int world, int B) { return ((j / B) * world + rank) * B + j % B; }
int shard_block(); // B of this process (MRX_SHARD_BLOCK; comm.cu)

/// sharded apply: rank-major staging buffer (row shard_row(i) = work-vector item i) -> node store,
/// coefficient blocks and (normRows, same layout, 8 per row) component norms
void launch_unpack_nodes(double *coefs, const double *stage, const int *gslotsAll, int nG, int world, int rowsPerRank, int shardB, int ncoef,
                         const double *normRows, double *gNorms, cudaStream_t st);

// ---- refinement step on the device (apply_split.cu) ------------------------------------------------------
struct SplitResult {
    double squareNorm, sNorm, wNorm; // TreeBuilder bookkeeping after this iteration
    int nSplit, nNext;               // nodes that split; size of the next work vector (8 nSplit)
    int nLoc, nChunksLoc;            // prep_local: this rank's items and probe chunks of the (next) work vector
    long long nbrCapLoc;             // upper bound of its neighbour list
};
struct SplitParams {
    const double *normRows; // component norms of the iteration, rank-major rows of 8 (item i = row shard_row(i, world, rows, shardB))
    int nG, world, rows, shardB;
    const int4 *gNodesAll;          // [nG] (operator depth, lx, ly, lz) in work-vector order
    const unsigned char *isBranch;  // [nG] or nullptr (no item is a branch node)
    int operRoot, rootScale, maxScale;
    const double *scaleFac;         // [depth] 2^{-(scale+1)/2} as the host's std::pow gives it (split_check)
    double prec;
    const double *precFacAll;       // [nG] per-node precision factor (apply with precTrees, WaveletAdaptor.h:51-54) or nullptr
    int absPrec, iter, doSplit;
    int slotBase;                   // slot of the first child created by this iteration
    double *state;                  // [3] sNorm, wNorm, squareNorm carried across iterations
    int4 *gNodesNext;               // [8 nSplit] next work vector
    int *slotsNext;
    const int *slotsCur;            // [nG] slots of this iteration's items (for the pair list) or nullptr
    int *pairsNext;                 // [2 nSplit] (parent slot, first child slot) of the nodes that split: the level list of the
                                    // closing TopDown(+=) step parent -> children, run as soon as the children exist (or nullptr)
    unsigned char *flags;           // [nG] split decisions (host replays them into its topology)
    SplitResult *res;
};
struct PrepParams {
    int nG;                  // >= 0: given by the host; < 0: read res->nNext (written by split_kernel just before)
    int world, rank, shardB;
    const int4 *gNodesAll;
    const int *slotsAll;
    const int *offCount;     // [DM] reachable offsets per depth (band tables)
    const DepthInfo *depthInfo;
    int DM;
    int4 *gNodesLoc;         // this rank's items (i = shard_item(rank, j, world, shardB))
    int *slotsLoc;
    const double *precAll;   // per-item precision factors of the whole work vector (or nullptr) ...
    double *precLoc;         // ... and this rank's share
    int *chunkOffLoc;        // [nLoc+1]
    SplitResult *res;
};
void launch_split(const SplitParams &S, cudaStream_t st);
/// host mailbox: copy `words` 32-bit words from device memory to (mapped, pinned) host memory, fence, then store `seq` into the
/// host flag -- one tiny kernel at the end of a dependent chain, so that the host can spin on the flag instead of paying a
/// cudaMemcpyAsync + cudaStreamSynchronize round trip per read-back
void launch_publish(const void *src, void *hostDst, int words, unsigned *hostFlag, unsigned seq, cudaStream_t st);

// ---- locally scaled precision (apply_prec.cu) ---------------------------------------------------------------
/// one precision tree on the device: real-node topology, node store, and per real node the value getMaxSquareNorm() answers
/// (stored maximum over the node and its descendants if positive, else the node's own scaled square norm; MWNode.h:84)
struct PrecTreeDev {
    const int *child0;
    const double *coefs;
    const double *vReal;
};
struct PrecParams {
    const int4 *gNodesAll; // (operator depth, lx, ly, lz) of the work vector
    int nG;                // number of items, or ...
    const int *nGptr;      // ... read from the device (next work vector, written by split_kernel) when not null
    int depthShift;        // function-tree depth = operator depth + depthShift
    const PrecTreeDev *trees;
    int nTrees;
    int corner[3], nboxes[3];
    int K, rootScale;
    const double *filters;
    double *precFacAll;    // [nG] out: 1 / max_i sqrt(maxSquareNorm_i(idx))
};
void launch_prec_factor(const PrecParams &P, int gridCap, cudaStream_t st);
void launch_prep_local(const PrepParams &P, cudaStream_t st);

} // namespace mrx
