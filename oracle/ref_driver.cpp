// C interface to the REAL reference (MRCPP sources compiled in place into oracle/_ref/, see oracle/build_ref.sh) for the
// tests: build trees and operators through the reference's own public API, run mrcpp::apply / project / mwTransform, and
// export nodes keyed by (scale, translation). TEST INFRASTRUCTURE only. Everything computed here is computed by the
// reference's code; the only foreign part is the dense linear-algebra stand-in for Eigen (oracle/eigen_shim).
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "MRCPP/Gaussians"
#include "MRCPP/MWFunctions"
#include "MRCPP/MWOperators"
#include "MRCPP/Printer"
#include "MRCPP/Timer"
#include "treebuilders/add.h"
#include "treebuilders/apply.h"
#include "treebuilders/grid.h"
#include "treebuilders/multiply.h"
#include "treebuilders/project.h"
#include "trees/FunctionTree.h"
#include "trees/MWNode.h"
#include "utils/tree_utils.h"

using namespace mrcpp;

namespace {
struct RefTree {
    FunctionTree<3, double> tree;
    explicit RefTree(const MultiResolutionAnalysis<3> &mra) : tree(mra) {}
};
GaussExp<3> make_exp(int n, const double *coef, const double *alpha, const double *pos, const int *power) {
    GaussExp<3> g;
    for (int i = 0; i < n; i++) {
        Coord<3> p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        std::array<int, 3> pw{0, 0, 0};
        if (power)
            for (int d = 0; d < 3; d++) pw[d] = power[3 * i + d];
        GaussFunc<3> f(alpha[i], coef[i], p, pw);
        g.append(f);
    }
    return g;
}
} // namespace

extern "C" {

void ref_init() {
    static bool done = false;
    if (!done) {
        Printer::init(-1);
        done = true;
    }
}

void *ref_mra_create(int order, int root_scale, const int *corner, const int *nboxes, int max_depth) {
    ref_init();
    std::array<int, 3> c{corner[0], corner[1], corner[2]}, b{nboxes[0], nboxes[1], nboxes[2]};
    BoundingBox<3> world(root_scale, c, b);
    InterpolatingBasis basis(order);
    return new MultiResolutionAnalysis<3>(world, basis, max_depth);
}
/// periodic world (BoundingBox(n, l, nb, sf, pbc = true), src/trees/BoundingBox.cpp:95-117): unit cell [-1, 1]^3, scaling factor 1
void *ref_mra_create_periodic(int order, int max_depth) {
    ref_init();
    std::array<int, 3> c{-1, -1, -1}, b{2, 2, 2};
    std::array<double, 3> sf{1.0, 1.0, 1.0};
    BoundingBox<3> world(0, c, b, sf, true);
    InterpolatingBasis basis(order);
    return new MultiResolutionAnalysis<3>(world, basis, max_depth);
}
/// project(prec, out, f) (src/treebuilders/project.cpp:85-104) of f(r) = sum_i amp[i] prod_d cos(pi k[3 i + d] r_d)
void ref_project_cosines(void *t, double prec, int n, const double *amp, const double *k) {
    auto &tree = static_cast<RefTree *>(t)->tree;
    std::function<double(const Coord<3> &)> f = [n, amp, k](const Coord<3> &r) {
        double s = 0.0;
        for (int i = 0; i < n; i++) {
            double p = amp[i];
            for (int d = 0; d < 3; d++) p *= std::cos(pi * k[3 * i + d] * r[d]);
            s += p;
        }
        return s;
    };
    project<3, double>(prec, tree, f);
}
/// operators for periodic worlds: PoissonOperator(mra, prec, root, reach) (PoissonOperator.cpp:56-77), HelmholtzOperator likewise
void *ref_poisson_create_reach(void *mra, double prec, int root, int reach) {
    return new PoissonOperator(*static_cast<MultiResolutionAnalysis<3> *>(mra), prec, root, reach);
}
void *ref_helmholtz_create_reach(void *mra, double mu, double prec, int root, int reach) {
    return new HelmholtzOperator(*static_cast<MultiResolutionAnalysis<3> *>(mra), mu, prec, root, reach);
}
/// apply_near_field / apply_far_field (src/treebuilders/apply.cpp:294-342)
void ref_apply_unit_cell(int inside, double prec, void *out, void *oper, void *inp, int max_iter, int abs_prec) {
    auto &o = static_cast<RefTree *>(out)->tree;
    auto &i = static_cast<RefTree *>(inp)->tree;
    auto &P = *static_cast<ConvolutionOperator<3> *>(oper);
    if (inside) apply_near_field<3, double>(prec, o, P, i, max_iter, abs_prec != 0);
    else apply_far_field<3, double>(prec, o, P, i, max_iter, abs_prec != 0);
}
void ref_mra_destroy(void *m) { delete static_cast<MultiResolutionAnalysis<3> *>(m); }

void *ref_tree_create(void *mra) { return new RefTree(*static_cast<MultiResolutionAnalysis<3> *>(mra)); }
void ref_tree_destroy(void *t) { delete static_cast<RefTree *>(t); }
int ref_tree_n_nodes(void *t) { return static_cast<RefTree *>(t)->tree.getNNodes(); }
double ref_tree_square_norm(void *t) { return static_cast<RefTree *>(t)->tree.getSquareNorm(); }
double ref_tree_evalf(void *t, const double *r, int precise) {
    Coord<3> x{r[0], r[1], r[2]};
    auto &tree = static_cast<RefTree *>(t)->tree;
    return precise ? tree.evalf_precise(x) : tree.evalf(x);
}
/// FunctionTree::saveTreeTXT / loadTreeTXT (src/trees/FunctionTree.cpp:240-372): the text interchange format
void ref_tree_save_txt(void *t, const char *path) { static_cast<RefTree *>(t)->tree.saveTreeTXT(path); }
void ref_tree_load_txt(void *t, const char *path) { static_cast<RefTree *>(t)->tree.loadTreeTXT(path); }
double ref_tree_integrate(void *t) { return static_cast<RefTree *>(t)->tree.integrate(); }
/// build_grid alone (src/treebuilders/grid.cpp:106-123)
void ref_build_grid_gaussians(void *t, int n, const double *coef, const double *alpha, const double *pos, const int *power) {
    GaussExp<3> g = make_exp(n, coef, alpha, pos, power);
    build_grid(static_cast<RefTree *>(t)->tree, g);
}

/// build_grid + project of a Gaussian expansion (src/treebuilders/grid.cpp:106-123, project.cpp:85-104)
void ref_project_gaussians(void *t, double prec, int n, const double *coef, const double *alpha, const double *pos, const int *power,
                           int do_build_grid) {
    auto &tree = static_cast<RefTree *>(t)->tree;
    GaussExp<3> g = make_exp(n, coef, alpha, pos, power);
    if (do_build_grid) build_grid(tree, g);
    project<3, double>(prec, tree, g);
}

void *ref_poisson_create(void *mra, double prec) { return new PoissonOperator(*static_cast<MultiResolutionAnalysis<3> *>(mra), prec); }
void *ref_helmholtz_create(void *mra, double mu, double prec) {
    return new HelmholtzOperator(*static_cast<MultiResolutionAnalysis<3> *>(mra), mu, prec);
}
void *ref_abgv_create(void *mra, double a, double b) { return new ABGVOperator<3>(*static_cast<MultiResolutionAnalysis<3> *>(mra), a, b); }
int ref_oper_n_terms(void *o) { return static_cast<ConvolutionOperator<3> *>(o)->size(); }
void ref_conv_destroy(void *o) { delete static_cast<ConvolutionOperator<3> *>(o); }
void ref_deriv_destroy(void *o) { delete static_cast<DerivativeOperator<3> *>(o); }

/// mrcpp::apply(prec, out, oper, inp, maxIter, absPrec) (src/treebuilders/apply.cpp:68-93); returns seconds
double ref_apply(double prec, void *out, void *oper, void *inp, int max_iter, int abs_prec) {
    Timer t;
    apply(prec, static_cast<RefTree *>(out)->tree, *static_cast<ConvolutionOperator<3> *>(oper), static_cast<RefTree *>(inp)->tree, max_iter,
          abs_prec != 0);
    t.stop();
    return t.elapsed();
}
/// mrcpp::apply(out, DerivativeOperator, inp, dir) (apply.cpp:379-412)
/// apply(prec, out, oper, inp, precTrees, maxIter, absPrec) (src/treebuilders/apply.cpp:214-251)
double ref_apply_prec_trees(double prec, void *out, void *oper, void *inp, int n, void **precTrees, int maxIter, int absPrec) {
    FunctionTreeVector<3, double> vec;
    for (int i = 0; i < n; i++) vec.push_back(std::make_tuple(1.0, &static_cast<RefTree *>(precTrees[i])->tree));
    apply(prec, static_cast<RefTree *>(out)->tree, *static_cast<ConvolutionOperator<3> *>(oper), static_cast<RefTree *>(inp)->tree, vec, maxIter,
          absPrec != 0);
    return static_cast<RefTree *>(out)->tree.getSquareNorm();
}
void ref_apply_derivative(void *out, void *oper, void *inp, int dir) {
    apply(static_cast<RefTree *>(out)->tree, *static_cast<DerivativeOperator<3> *>(oper), static_cast<RefTree *>(inp)->tree, dir);
}
double ref_dot(void *a, void *b) { return dot(static_cast<RefTree *>(a)->tree, static_cast<RefTree *>(b)->tree); }
/// build_grid(out, inp): extend the grid of `out` with the nodes of `inp` (src/treebuilders/grid.cpp:144-153)
void ref_build_grid_tree(void *out, void *inp) { build_grid(static_cast<RefTree *>(out)->tree, static_cast<RefTree *>(inp)->tree); }
/// add(-1.0, out, {(c_i, inp_i)}, 0): addition on the grid `out` enters with (src/treebuilders/add.cpp:41-70)
void ref_add(void *out, int n, const double *coefs, void **inp) {
    FunctionTreeVector<3, double> vec;
    for (int i = 0; i < n; i++) vec.push_back(std::make_tuple(coefs[i], &static_cast<RefTree *>(inp[i])->tree));
    add(-1.0, static_cast<RefTree *>(out)->tree, vec, 0);
}
/// add(prec, out, {(c_i, inp_i)}, maxIter, absPrec): the adaptive form
void ref_add_adaptive(double prec, void *out, int n, const double *coefs, void **inp, int maxIter, int absPrec) {
    FunctionTreeVector<3, double> vec;
    for (int i = 0; i < n; i++) vec.push_back(std::make_tuple(coefs[i], &static_cast<RefTree *>(inp[i])->tree));
    add(prec, static_cast<RefTree *>(out)->tree, vec, maxIter, absPrec != 0);
}
/// multiply(prec, out, {(c_i, inp_i)}, maxIter, absPrec) (src/treebuilders/multiply.cpp:104-136)
void ref_multiply(double prec, void *out, int n, const double *coefs, void **inp, int maxIter, int absPrec, int useMaxNorms) {
    FunctionTreeVector<3, double> vec;
    for (int i = 0; i < n; i++) vec.push_back(std::make_tuple(coefs[i], &static_cast<RefTree *>(inp[i])->tree));
    multiply(prec, static_cast<RefTree *>(out)->tree, vec, maxIter, absPrec != 0, useMaxNorms != 0);
}
int ref_refine_grid(void *t, double prec, int absPrec, int scales) {
    auto &tree = static_cast<RefTree *>(t)->tree;
    return scales > 0 ? refine_grid(tree, scales) : refine_grid(tree, prec, absPrec != 0);
}
void ref_add_inplace(void *out, double c, void *inp) { static_cast<RefTree *>(out)->tree.add(c, static_cast<RefTree *>(inp)->tree); }
void ref_clear_grid(void *t) { clear_grid(static_cast<RefTree *>(t)->tree); }
void ref_power(double prec, void *out, void *inp, double p, int maxIter, int absPrec) {
    power(prec, static_cast<RefTree *>(out)->tree, static_cast<RefTree *>(inp)->tree, p, maxIter, absPrec != 0);
}
/// divergence(out, oper, {inp_x, inp_y, inp_z}) (src/treebuilders/apply.cpp:514-530)
void ref_divergence(void *out, void *oper, void **inp) {
    FunctionTreeVector<3, double> vec;
    for (int d = 0; d < 3; d++) vec.push_back(std::make_tuple(1.0, &static_cast<RefTree *>(inp[d])->tree));
    divergence(static_cast<RefTree *>(out)->tree, *static_cast<DerivativeOperator<3> *>(oper), vec);
}
void *ref_ph_create(void *mra, int order) {
    return static_cast<DerivativeOperator<3> *>(new PHOperator<3>(*static_cast<MultiResolutionAnalysis<3> *>(mra), order));
}
void *ref_bs_create(void *mra, int order) {
    return static_cast<DerivativeOperator<3> *>(new BSOperator<3>(*static_cast<MultiResolutionAnalysis<3> *>(mra), order));
}
void ref_copy_grid(void *out, void *inp) { copy_grid(static_cast<RefTree *>(out)->tree, static_cast<RefTree *>(inp)->tree); }
void ref_mw_transform(void *t, int type, int overwrite) { static_cast<RefTree *>(t)->tree.mwTransform(type, overwrite != 0); }

/// every node (depth by depth, table order): scale, translation, is-branch flag, 8 (k+1)^3 coefficients, 8 component norms
int ref_tree_export(void *t, int *scale, int *transl, int *branch, double *coefs, double *norms) {
    auto &tree = static_cast<RefTree *>(t)->tree;
    std::vector<MWNodeVector<3, double>> table;
    tree_utils::make_node_table(tree, table);
    int n = 0;
    for (auto &level : table)
        for (MWNode<3, double> *nd : level) {
            if (scale) scale[n] = nd->getScale();
            if (transl)
                for (int d = 0; d < 3; d++) transl[3 * n + d] = nd->getNodeIndex()[d];
            if (branch) branch[n] = nd->isBranchNode() ? 1 : 0;
            if (coefs) std::memcpy(coefs + (size_t)n * nd->getNCoefs(), nd->getCoefs(), sizeof(double) * nd->getNCoefs());
            if (norms)
                for (int c = 0; c < 8; c++) norms[8 * n + c] = nd->getComponentNorm(c);
            n++;
        }
    return n;
}
int ref_num_threads() { return mrcpp_get_num_threads(); }
}
