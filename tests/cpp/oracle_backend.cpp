// TEST INFRASTRUCTURE (CPU suite only, never shipped): device entry points of the C ABI re-defined on top of the CPU oracle
// (oracle/_build/liboracle.so, loaded with dlopen) and the library's own host-side entry points, compiled INTO a test
// executable so that the program logic of tests/cpp/*.cpp and the plumbing of the C++ mirror (include/MRCPP/) can be checked
// where no GPU exists. The executable's definitions take precedence over the library's; the product library itself is
// untouched and still aborts on every hot-path call without a device.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mrcpp_b200.h"
#include "../../mrcpp_b200/csrc/host/mrx_host.hpp"

namespace {
struct orc_stats {
    long long gNodes, fApplied, genUsed;
    int iters, nNodesOut;
    double t_band, t_calc, t_post, t_total;
};
struct Oracle {
    void (*apply)(double, void *, void *, void *, int, int, orc_stats *);
    void (*apply_derivative)(void *, void *, void *, int, orc_stats *);
    void (*down)(void *, int);
    void (*up)(void *);
    void (*sqnorm)(void *);
    double (*dot)(void *, void *);
    void (*set_tables)(const char *);
    void (*add)(double, void *, int, const double *, void **, int, int);
    void (*multiply)(double, void *, int, const double *, void **, int, int, int);
    int (*refine_grid)(void *, double, int, int);
    void (*power)(double, void *, void *, double, int, int);
    void (*add_inplace)(void *, double, void *);
    void (*apply_unit_cell)(int, double, void *, void *, void *, int, int, orc_stats *);
    void (*apply_prec_trees)(double, void *, void *, void *, int, void **, int, int, orc_stats *);
};
Oracle &oracle() {
    static Oracle o = [] {
        const char *path = std::getenv("MRX_TEST_ORACLE");
        void *h = dlopen(path ? path : "oracle/_build/liboracle.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            std::fprintf(stderr, "oracle_backend: cannot load the oracle (%s)\n", dlerror());
            std::abort();
        }
        Oracle r;
        r.apply = reinterpret_cast<decltype(r.apply)>(dlsym(h, "orc_apply"));
        r.apply_derivative = reinterpret_cast<decltype(r.apply_derivative)>(dlsym(h, "orc_apply_derivative"));
        r.down = reinterpret_cast<decltype(r.down)>(dlsym(h, "orc_mw_transform_down"));
        r.up = reinterpret_cast<decltype(r.up)>(dlsym(h, "orc_mw_transform_up"));
        r.sqnorm = reinterpret_cast<decltype(r.sqnorm)>(dlsym(h, "orc_calc_square_norm"));
        r.dot = reinterpret_cast<decltype(r.dot)>(dlsym(h, "orc_dot"));
        r.set_tables = reinterpret_cast<decltype(r.set_tables)>(dlsym(h, "orc_set_table_path"));
        r.add = reinterpret_cast<decltype(r.add)>(dlsym(h, "orc_add"));
        r.multiply = reinterpret_cast<decltype(r.multiply)>(dlsym(h, "orc_multiply"));
        r.refine_grid = reinterpret_cast<decltype(r.refine_grid)>(dlsym(h, "orc_refine_grid"));
        r.power = reinterpret_cast<decltype(r.power)>(dlsym(h, "orc_power"));
        r.add_inplace = reinterpret_cast<decltype(r.add_inplace)>(dlsym(h, "orc_add_inplace"));
        r.apply_unit_cell = reinterpret_cast<decltype(r.apply_unit_cell)>(dlsym(h, "orc_apply_unit_cell"));
        r.apply_prec_trees = reinterpret_cast<decltype(r.apply_prec_trees)>(dlsym(h, "orc_apply_prec_trees"));
        if (const char *t = std::getenv("MRX_TABLES")) r.set_tables(t);
        return r;
    }();
    return o;
}
template <typename F> F next(const char *name) { return reinterpret_cast<F>(dlsym(RTLD_NEXT, name)); }
} // namespace

extern "C" {

int mrx_mw_transform(mrx_tree *tree, int type, int overwrite) {
    if (type == MRX_BOTTOM_UP) oracle().up(mrx_tree_host_handle(tree));
    else oracle().down(mrx_tree_host_handle(tree), overwrite);
    mrx_tree_host_modified(tree);
    return 0;
}
double mrx_calc_square_norm(mrx_tree *tree) {
    oracle().sqnorm(mrx_tree_host_handle(tree));
    return mrx_tree_square_norm(tree);
}
int mrx_project_gaussians_device(mrx_tree *tree, double prec, int n, const double *coef, const double *alpha, const double *pos,
                                 const int *power, int build_grid) {
    mrx_project_gaussians(tree, prec, n, coef, alpha, pos, power, build_grid, /*finalize=*/0);
    mrx_mw_transform(tree, MRX_BOTTOM_UP, 1);
    mrx_calc_square_norm(tree);
    return 0;
}
int mrx_project_function(mrx_tree *tree, double prec, mrx_func3 f, void *user, int threads_ok, int finalize) {
    static auto real = next<int (*)(mrx_tree *, double, mrx_func3, void *, int, int)>("mrx_project_function");
    real(tree, prec, f, user, threads_ok, 0);
    if (finalize) {
        mrx_mw_transform(tree, MRX_BOTTOM_UP, 1);
        mrx_calc_square_norm(tree);
    }
    return 0;
}
int mrx_apply(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec, mrx_apply_stats *stats) {
    orc_stats st{};
    oracle().apply(prec, mrx_tree_host_handle(out), mrx_oper_host_handle(oper), mrx_tree_host_handle(inp), max_iter, abs_prec, &st);
    mrx_tree_host_modified(out);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->g_nodes = st.gNodes;
        stats->f_applied = st.fApplied;
        stats->gen_nodes = st.genUsed;
        stats->iterations = st.iters;
        stats->n_nodes_out = st.nNodesOut;
    }
    return 0;
}
static void copy_stats(const orc_stats &st, mrx_apply_stats *stats) {
    if (!stats) return;
    std::memset(stats, 0, sizeof(*stats));
    stats->g_nodes = st.gNodes;
    stats->f_applied = st.fApplied;
    stats->gen_nodes = st.genUsed;
    stats->iterations = st.iters;
    stats->n_nodes_out = st.nNodesOut;
}
int mrx_apply_unit_cell(int inside, double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int max_iter, int abs_prec,
                        mrx_apply_stats *stats) {
    orc_stats st{};
    oracle().apply_unit_cell(inside, prec, mrx_tree_host_handle(out), mrx_oper_host_handle(oper), mrx_tree_host_handle(inp), max_iter, abs_prec, &st);
    mrx_tree_host_modified(out);
    copy_stats(st, stats);
    return 0;
}
int mrx_apply_prec_trees(double prec, mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int n_prec, mrx_tree *const *prec_trees, int max_iter,
                         int abs_prec, const mrx_comm *, mrx_apply_stats *stats) {
    orc_stats st{};
    std::vector<void *> h(n_prec > 0 ? n_prec : 0);
    for (int i = 0; i < n_prec; i++) h[i] = mrx_tree_host_handle(prec_trees[i]);
    oracle().apply_prec_trees(prec, mrx_tree_host_handle(out), mrx_oper_host_handle(oper), mrx_tree_host_handle(inp), n_prec, h.data(), max_iter,
                              abs_prec, &st);
    mrx_tree_host_modified(out);
    copy_stats(st, stats);
    return 0;
}
int mrx_apply_derivative(mrx_tree *out, mrx_oper *oper, mrx_tree *inp, int dir, mrx_apply_stats *stats) {
    orc_stats st{};
    oracle().apply_derivative(mrx_tree_host_handle(out), mrx_oper_host_handle(oper), mrx_tree_host_handle(inp), dir, &st);
    mrx_tree_host_modified(out);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->g_nodes = st.gNodes;
        stats->f_applied = st.fApplied;
        stats->n_nodes_out = st.nNodesOut;
    }
    return 0;
}
int mrx_tree_add_adaptive(double prec, mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp, int max_iter, int abs_prec) {
    std::vector<void *> h(n);
    for (int i = 0; i < n; i++) h[i] = mrx_tree_host_handle(inp[i]);
    oracle().add(prec, mrx_tree_host_handle(out), n, coefs, h.data(), max_iter, abs_prec);
    mrx_tree_host_modified(out);
    return 0;
}
int mrx_tree_power(double prec, mrx_tree *out, mrx_tree *inp, double p, int max_iter, int abs_prec) {
    oracle().power(prec, mrx_tree_host_handle(out), mrx_tree_host_handle(inp), p, max_iter, abs_prec);
    mrx_tree_host_modified(out);
    return 0;
}
int mrx_tree_refine_grid(mrx_tree *tree, double prec, int abs_prec, int scales) {
    const int n = oracle().refine_grid(mrx_tree_host_handle(tree), prec, abs_prec, scales);
    mrx_tree_host_modified(tree);
    return n;
}
int mrx_tree_add_inplace(mrx_tree *tree, double c, mrx_tree *inp) {
    oracle().add_inplace(mrx_tree_host_handle(tree), c, mrx_tree_host_handle(inp));
    mrx_tree_host_modified(tree);
    return 0;
}
int mrx_tree_multiply(double prec, mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp, int max_iter, int abs_prec,
                      int use_max_norms) {
    std::vector<void *> h(n);
    for (int i = 0; i < n; i++) h[i] = mrx_tree_host_handle(inp[i]);
    oracle().multiply(prec, mrx_tree_host_handle(out), n, coefs, h.data(), max_iter, abs_prec, use_max_norms);
    mrx_tree_host_modified(out);
    return 0;
}
int mrx_tree_add(mrx_tree *out, int n, const double *coefs, mrx_tree *const *inp) {
    return mrx_tree_add_adaptive(-1.0, out, n, coefs, inp, 0, 0);
}
double mrx_dot(mrx_tree *bra, mrx_tree *ket) { return oracle().dot(mrx_tree_host_handle(bra), mrx_tree_host_handle(ket)); }
int mrx_tree_rescale(mrx_tree *tree, double c) {
    auto *h = static_cast<mrx::Tree<3> *>(mrx_tree_host_handle(tree));
    for (int n = 0; n < h->nReal; n++) {
        double *p = h->coef(n);
        for (int i = 0; i < h->ncoef; i++) p[i] *= c;
        h->calcNorms(n);
    }
    h->calcSquareNorm();
    mrx_tree_host_modified(tree);
    return 0;
}
}
