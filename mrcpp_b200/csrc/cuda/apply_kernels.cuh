// Parameter blocks of the operator-application kernels (apply_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace mrx {

/// one output node of the current work vector
struct GDesc {
    int slot;   // slot in the output tree
    int depth;  // operator depth = scale - operator root
    int l[3];   // translation
    int s[3];   // first translation of the (clipped) input band
    int nb[3];  // band extent per dimension (0 = empty band)
    int nbrOff; // offset of this node's neighbour slots in `nbr` (x fastest)
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
    const GDesc *gdesc;
    const int *nbr;
    // operator tables
    const double *mats;   // node-major: 4 blocks of K*K (column-major p[i + K m], i = input, m = output)
    const double *onorms; // 4 per node
    const int *nodeOff;   // [M][DM]
    const int *maxTransl; // [M][DM]
    const int *bw;        // [M][DM][5]
    const int *bsf;       // [M][DM][64]
    int M, DM, K;
    double gThrs;
    unsigned long long *counters; // [0] tuples applied
    // derivative apply
    int derivDir; // -1 for convolution operators
};

void launch_apply(const ApplyParams &P, int nG, cudaStream_t st);

} // namespace mrx
