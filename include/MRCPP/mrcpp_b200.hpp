// C++ host mirror of the MRCPP API for the B200 operator-application path, header-only, over the C ABI of
// libmrcpp_b200.so (include/mrcpp_b200.h). Same names, argument meaning and error behaviour (print + abort) as the
// reference classes it stands in for, so that a program written against MRCPP's public headers
//
//     #include "MRCPP/Gaussians"   #include "MRCPP/MWFunctions"   #include "MRCPP/MWOperators"
//     #include "MRCPP/Printer"     #include "MRCPP/Timer"
//
// (e.g. the reference's examples/poisson.cpp and examples/projection.cpp, unmodified) compiles against -Iinclude and
// links with -lmrcpp_b200. Only what the path needs is here: D = 3, T = double, interpolating basis; periodic worlds as the unit
// cell [-1, 1]^3 without scaling factors.
// Anything else is a compile-time error (static_assert) or aborts with a message, never a silent CPU fallback: the
// arithmetic of every call below runs in the library (CUDA); this file only holds handles and formats output.
//
// Reference interfaces mirrored (file:line under the MRCPP source tree):
//   BoundingBox<D>               src/trees/BoundingBox.h:53-60
//   InterpolatingBasis           src/core/InterpolatingBasis.h:43
//   MultiResolutionAnalysis<D>   src/trees/MultiResolutionAnalysis.h:51-54
//   GaussFunc<D>, GaussExp<D>    src/functions/GaussFunc.h:56, GaussExp.h:54-118
//   FunctionTree<D, T>           src/trees/FunctionTree.h, MWTree.h:97-179
//   ConvolutionOperator<D>, PoissonOperator, HelmholtzOperator, DerivativeOperator<D>, ABGVOperator<D>, PHOperator<D>, BSOperator<D>
//                                src/operators/{ConvolutionOperator,PoissonOperator,HelmholtzOperator,ABGVOperator}.h
//   build_grid, copy_grid, clear_grid   src/treebuilders/grid.h:35-43
//   project                      src/treebuilders/project.h:33-34
//   apply (convolution, derivative), gradient, divergence   src/treebuilders/apply.h:41,49,51,55
//   add, multiply, square        src/treebuilders/add.h, multiply.h
//   FunctionTreeVector           src/trees/FunctionTreeVector.h
//   dot                          src/treebuilders/multiply.h
//   Printer, print::*, Timer     src/utils/Printer.h:61-133, src/utils/Timer.h:42-50
#pragma once

#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "../mrcpp_b200.h"

namespace mrcpp {

// ---- api/constants.h
const double MachinePrec = 1.0e-15;
const double MachineZero = 1.0e-14;
const int MaxOrder = 41;
const int MaxDepth = 30;
const int MaxScale = 31;
const int MinScale = -31;
enum FuncType { Legendre, Interpol };
enum Traverse { TopDown, BottomUp };
const double pi = 3.1415926535897932384626433832795;
const double root_pi = 1.7724538509055160273;

template <int D> using Coord = std::array<double, D>;

// ---- error convention: print + abort (src/utils/Printer.h:165-190)
#define MRCPP_B200_ABORT(X)                                                                                                    \
    {                                                                                                                          \
        std::cerr << "Error: " << __FILE__ << ": " << __func__ << "(), line " << __LINE__ << ": " << X << std::endl;           \
        std::abort();                                                                                                          \
    }

// ---- library / device selection -------------------------------------------------------------------------------------
namespace b200 {
/// Selects the CUDA device of this process (one process per GPU) and loads the filter tables. Called implicitly by the
/// first object that needs the library: device = $MRCPP_B200_DEVICE if set, else 0 when a GPU is visible, else -1
/// (host-only: construction helpers work, every hot-path call aborts). An MPI program calls init(local_rank) first.
inline int &device_ref() {
    static int dev = -2; // -2: not initialised
    return dev;
}
inline void init(int device, const char *tables = nullptr) {
    if (device_ref() != -2) {
        if (device_ref() != device) MRCPP_B200_ABORT("library already initialised on device " << device_ref());
        return;
    }
    if (tables == nullptr) tables = std::getenv("MRCPP_B200_TABLES"); // else: data/mwtables.bin next to the library
    if (mrx_init(tables ? tables : "", device) != 0) MRCPP_B200_ABORT("mrx_init failed for device " << device);
    device_ref() = device;
}
inline void ensure_init() {
    if (device_ref() != -2) return;
    int device = mrx_device_count() > 0 ? 0 : -1;
    if (const char *env = std::getenv("MRCPP_B200_DEVICE")) device = std::atoi(env);
    init(device);
}
} // namespace b200

// ---- Timer (src/utils/Timer.h:42-50) -----------------------------------------------------------------------------------
class Timer final {
public:
    explicit Timer(bool start_timer = true) {
        if (start_timer) start();
    }
    void start() {
        clock_start = now();
        time_used = 0.0;
        running = true;
    }
    void resume() {
        if (running) std::cerr << "Warning: timer already running" << std::endl;
        clock_start = now();
        running = true;
    }
    void stop() {
        if (!running) std::cerr << "Warning: timer not running" << std::endl;
        time_used += diff(now(), clock_start);
        running = false;
    }
    double elapsed() const { return running ? diff(now(), clock_start) : time_used; }

private:
    using timeT = std::chrono::time_point<std::chrono::high_resolution_clock>;
    bool running{false};
    double time_used{0.0};
    timeT clock_start;
    static timeT now() { return std::chrono::high_resolution_clock::now(); }
    static double diff(timeT t2, timeT t1) { return std::chrono::duration<double>(t2 - t1).count(); }
};

// ---- BoundingBox / basis / MRA ----------------------------------------------------------------------------------------
template <int D> class BoundingBox {
public:
    explicit BoundingBox(int n = 0, const std::array<int, D> &l = {}, const std::array<int, D> &nb = {},
                         const std::array<double, D> &sf = {}, bool pbc = false)
            : scale(n)
            , corner(l)
            , boxes(nb)
            , periodic(pbc) {
        for (int d = 0; d < D; d++) {
            if (boxes[d] <= 0) boxes[d] = 1; // BoundingBox.cpp: zero means one box
            if (sf[d] != 0.0 && sf[d] != 1.0) MRCPP_B200_ABORT("scaling factors are not supported on the B200 path");
        }
        // periodic worlds (BoundingBox.cpp:95-117): the unit cell [-1, 1]^D in box units; mrx_mra_set_periodic checks the shape
    }
    int getScale() const { return scale; }
    int size(int d) const { return boxes[d]; }
    int size() const {
        int n = 1;
        for (int d = 0; d < D; d++) n *= boxes[d];
        return n;
    }
    const std::array<int, D> &getCornerIndex() const { return corner; }
    double getUnitLength(int) const { return std::pow(2.0, -scale); }
    double getBoxLength(int d) const { return getUnitLength(d) * boxes[d]; }
    double getLowerBound(int d) const { return getUnitLength(d) * corner[d]; }
    double getUpperBound(int d) const { return getUnitLength(d) * (corner[d] + boxes[d]); }
    bool isPeriodic() const { return periodic; }
    bool operator==(const BoundingBox<D> &o) const {
        return scale == o.scale && corner == o.corner && boxes == o.boxes && periodic == o.periodic;
    }
    bool operator!=(const BoundingBox<D> &o) const { return !(*this == o); }

private:
    int scale;
    std::array<int, D> corner;
    std::array<int, D> boxes;
    bool periodic{false};
};

class ScalingBasis {
public:
    ScalingBasis(int k, int t)
            : type(t)
            , order(k) {
        if (order < 0) MRCPP_B200_ABORT("Invalid scaling order");
    }
    virtual ~ScalingBasis() = default;
    int getScalingType() const { return type; }
    int getScalingOrder() const { return order; }
    int getQuadratureOrder() const { return order + 1; }

protected:
    const int type;
    const int order;
};

class InterpolatingBasis final : public ScalingBasis {
public:
    InterpolatingBasis(int k)
            : ScalingBasis(k, Interpol) {}
};

template <int D> class MultiResolutionAnalysis final {
public:
    MultiResolutionAnalysis(const BoundingBox<D> &bb, const ScalingBasis &sb, int depth = MaxDepth)
            : world(bb)
            , order(sb.getScalingOrder())
            , maxDepth(depth) {
        if (sb.getScalingType() != Interpol) MRCPP_B200_ABORT("only the interpolating basis is supported on the B200 path");
        setup();
    }
    MultiResolutionAnalysis(const BoundingBox<D> &bb, int k, int depth = MaxDepth)
            : world(bb)
            , order(k)
            , maxDepth(depth) {
        setup();
    }
    int getOrder() const { return order; }
    int getMaxDepth() const { return maxDepth; }
    int getMaxScale() const { return world.getScale() + maxDepth; }
    int getRootScale() const { return world.getScale(); }
    const BoundingBox<D> &getWorldBox() const { return world; }
    bool operator==(const MultiResolutionAnalysis<D> &o) const { return world == o.world && order == o.order && maxDepth == o.maxDepth; }
    bool operator!=(const MultiResolutionAnalysis<D> &o) const { return !(*this == o); }
    const mrx_mra *handle() const { return h.get(); }

private:
    BoundingBox<D> world;
    int order;
    int maxDepth;
    std::shared_ptr<mrx_mra> h;
    void setup() {
        static_assert(D == 3, "the B200 path implements 3-dimensional trees only");
        b200::ensure_init();
        int c[3], nb[3];
        for (int d = 0; d < 3; d++) {
            c[d] = world.getCornerIndex()[d];
            nb[d] = world.size(d);
        }
        h = std::shared_ptr<mrx_mra>(mrx_mra_create(order, world.getScale(), c, nb, maxDepth), mrx_mra_destroy);
        if (world.isPeriodic()) mrx_mra_set_periodic(h.get(), 1);
    }
};

// ---- analytic functions -----------------------------------------------------------------------------------------------
template <int D, typename T = double> class RepresentableFunction {
public:
    virtual ~RepresentableFunction() = default;
    virtual T evalf(const Coord<D> &r) const = 0;
};

template <int D> class Gaussian : public RepresentableFunction<D, double> {
public:
    Gaussian(double a, double c, const Coord<D> &r, const std::array<int, D> &p)
            : coef(c)
            , power(p)
            , pos(r) {
        alpha.fill(a);
    }
    double getCoef() const { return coef; }
    const std::array<double, D> &getExp() const { return alpha; }
    const std::array<int, D> &getPower() const { return power; }
    const Coord<D> &getPos() const { return pos; }
    void setCoef(double c) { coef = c; }
    void setPos(const Coord<D> &r) { pos = r; }

protected:
    double coef;
    std::array<int, D> power;
    std::array<double, D> alpha;
    Coord<D> pos;
};

/// coef * prod_d (x_d - pos_d)^pow_d * exp(-beta |x - pos|^2)   (src/functions/GaussFunc.cpp:47-66)
template <int D> class GaussFunc : public Gaussian<D> {
public:
    GaussFunc(double beta, double alpha, const Coord<D> &pos = {}, const std::array<int, D> &pow = {})
            : Gaussian<D>(beta, alpha, pos, pow) {}
    double evalf(const Coord<D> &r) const override {
        double q2 = 0.0, p2 = 1.0;
        for (int d = 0; d < D; d++) {
            const double q = r[d] - this->pos[d];
            q2 += this->alpha[d] * q * q;
            if (this->power[d] == 1) p2 *= q;
            else if (this->power[d] != 0) p2 *= std::pow(q, this->power[d]);
        }
        return this->coef * p2 * std::exp(-q2);
    }
    /// src/functions/GaussFunc.cpp:210-237: sqrt(4 a / pi) F_0(a R^2), a = p q / (p + q); both Gaussians normalised to unit charge
    double calcCoulombEnergy(const GaussFunc<D> &gf) const {
        static_assert(D == 3, "calcCoulombEnergy: 3-dimensional Gaussians only (as in the reference)");
        const double p = this->alpha[0], q = gf.alpha[0];
        const double a = p * q / (p + q);
        double R2 = 0.0;
        for (int d = 0; d < D; d++) R2 += (this->pos[d] - gf.pos[d]) * (this->pos[d] - gf.pos[d]);
        const double x = a * R2;
        const double boys = x < 1.0e-14 ? 1.0 : 0.5 * std::sqrt(pi / x) * std::erf(std::sqrt(x));
        return std::sqrt(4.0 * a / pi) * boys;
    }
};

template <int D> class GaussExp : public RepresentableFunction<D, double> {
public:
    GaussExp(int nTerms = 0) { funcs.reserve(nTerms); }
    double evalf(const Coord<D> &r) const override {
        double v = 0.0;
        for (const auto &f : funcs) v += f.evalf(r);
        return v;
    }
    int size() const { return (int)funcs.size(); }
    void append(const GaussFunc<D> &g) { funcs.push_back(g); }
    void append(const GaussExp<D> &g) { funcs.insert(funcs.end(), g.funcs.begin(), g.funcs.end()); }
    GaussFunc<D> &getFunc(int i) { return funcs[i]; }
    const GaussFunc<D> &getFunc(int i) const { return funcs[i]; }

private:
    std::vector<GaussFunc<D>> funcs;
};

namespace b200 {
/// flat arrays of a Gaussian expansion as the C ABI takes them
template <int D> struct GaussArrays {
    std::vector<double> coef, expo, pos;
    std::vector<int> power;
    void add(const GaussFunc<D> &g) {
        for (int d = 1; d < D; d++)
            if (g.getExp()[d] != g.getExp()[0]) MRCPP_B200_ABORT("anisotropic Gaussians are not supported on the B200 path");
        coef.push_back(g.getCoef());
        expo.push_back(g.getExp()[0]);
        for (int d = 0; d < D; d++) {
            pos.push_back(g.getPos()[d]);
            power.push_back(g.getPower()[d]);
        }
    }
    int size() const { return (int)coef.size(); }
};
/// Gaussian content of a RepresentableFunction, if it is one (GaussFunc or GaussExp)
template <int D> bool gauss_arrays(const RepresentableFunction<D, double> &f, GaussArrays<D> &out) {
    if (auto *g = dynamic_cast<const GaussFunc<D> *>(&f)) {
        out.add(*g);
        return true;
    }
    if (auto *e = dynamic_cast<const GaussExp<D> *>(&f)) {
        for (int i = 0; i < e->size(); i++) out.add(e->getFunc(i));
        return true;
    }
    return false;
}
} // namespace b200

// ---- FunctionTree -----------------------------------------------------------------------------------------------------
template <int D, typename T = double> class MWTree {
public:
    virtual ~MWTree() = default;
    MWTree(const MWTree &) = delete;
    MWTree &operator=(const MWTree &) = delete;

    const MultiResolutionAnalysis<D> &getMRA() const { return MRA; }
    int getOrder() const { return MRA.getOrder(); }
    int getKp1() const { return MRA.getOrder() + 1; }
    int getKp1_d() const {
        int n = 1;
        for (int d = 0; d < D; d++) n *= getKp1();
        return n;
    }
    int getRootScale() const { return MRA.getRootScale(); }
    double getSquareNorm() const { return mrx_tree_square_norm(h); }
    void calcSquareNorm() { mrx_calc_square_norm(h); }
    int getNNodes() const { return mrx_tree_n_nodes(h); }
    int getNEndNodes() const { return mrx_tree_n_end_nodes(h); }
    int getSizeNodes() const { return (int)(mrx_tree_bytes(h) / 1024); } // kB, MWTree.cpp:120-130
    void mwTransform(int type, bool overwrite = true) { mrx_mw_transform(h, type == TopDown ? MRX_TOP_DOWN : MRX_BOTTOM_UP, overwrite ? 1 : 0); }
    mrx_tree *handle() const { return h; }

protected:
    explicit MWTree(const MultiResolutionAnalysis<D> &mra)
            : MRA(mra)
            , h(mrx_tree_create(mra.handle())) {}
    const MultiResolutionAnalysis<D> MRA;
    mrx_tree *h;
};

template <int D, typename T = double> class FunctionTree final : public MWTree<D, T>, public RepresentableFunction<D, T> {
    static_assert(D == 3 && std::is_same<T, double>::value, "the B200 path implements FunctionTree<3, double> only");

public:
    explicit FunctionTree(const MultiResolutionAnalysis<D> &mra)
            : MWTree<D, T>(mra) {}
    ~FunctionTree() override { mrx_tree_destroy(this->h); }
    T integrate() const { return mrx_tree_integrate(this->h); }
    /// FunctionTree::add(c, inp): in place, on this tree's grid (src/trees/FunctionTree.cpp:687-706)
    void add(T c, FunctionTree<D, T> &inp) {
        if (this->getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
        mrx_tree_add_inplace(this->h, c, inp.handle());
    }
    /// FunctionTree::saveTreeTXT / loadTreeTXT (src/trees/FunctionTree.cpp:240-372): the text interchange format
    void saveTreeTXT(const std::string &file) { mrx_tree_save_txt(this->h, file.c_str()); }
    void loadTreeTXT(const std::string &file) { mrx_tree_load_txt(this->h, file.c_str()); }
    int getNGenNodes() const { return 0; } // generated nodes never outlive the call that made them
    void deleteGenerated() {}
    void rescale(T c) { mrx_tree_rescale(this->h, c); }
    void normalize() {
        const double sq = this->getSquareNorm();
        if (sq < 0.0) MRCPP_B200_ABORT("Normalizing uninitialized function");
        rescale(1.0 / std::sqrt(sq));
    }
    void clear() { mrx_tree_clear(this->h); }
    /// FunctionTree::evalf / evalf_precise (src/trees/FunctionTree.cpp:374-436): values from the downloaded tree
    T evalf(const Coord<D> &r) const override {
        double v = 0.0;
        mrx_tree_evalf(this->h, 1, r.data(), &v, 0);
        return v;
    }
    T evalf_precise(const Coord<D> &r) {
        double v = 0.0;
        mrx_tree_evalf(this->h, 1, r.data(), &v, 1);
        return v;
    }
};

// ---- FunctionTreeVector (src/trees/FunctionTreeVector.h): (coefficient, tree) pairs, trees not owned ------------------
template <int D, typename T = double> using CoefsFunctionTree = std::tuple<T, FunctionTree<D, T> *>;
template <int D, typename T = double> using FunctionTreeVector = std::vector<CoefsFunctionTree<D, T>>;
template <int D, typename T> T get_coef(const FunctionTreeVector<D, T> &fs, int i) { return std::get<0>(fs[i]); }
template <int D, typename T> FunctionTree<D, T> &get_func(FunctionTreeVector<D, T> &fs, int i) { return *std::get<1>(fs[i]); }
template <int D, typename T> const FunctionTree<D, T> &get_func(const FunctionTreeVector<D, T> &fs, int i) { return *std::get<1>(fs[i]); }
/// clear(fs, dealloc): FunctionTreeVector.h -- optionally deletes the trees, then empties the vector
template <int D, typename T> void clear(FunctionTreeVector<D, T> &fs, bool dealloc = false) {
    if (dealloc)
        for (auto &t : fs) delete std::get<1>(t);
    fs.clear();
}

// ---- operators ---------------------------------------------------------------------------------------------------------
class MWOperatorBase {
public:
    virtual ~MWOperatorBase() {
        if (h) mrx_oper_destroy(h);
    }
    MWOperatorBase(const MWOperatorBase &) = delete;
    MWOperatorBase &operator=(const MWOperatorBase &) = delete;
    int size() const { return mrx_oper_n_terms(h); } // separation rank (MWOperator::size)
    mrx_oper *handle() const { return h; }

protected:
    MWOperatorBase() = default;
    mrx_oper *h = nullptr;
};

template <int D> class ConvolutionOperator : public MWOperatorBase {
    static_assert(D == 3, "the B200 path implements 3-dimensional operators only");

public:
    /// ConvolutionOperator(mra, kernel, prec): src/operators/ConvolutionOperator.cpp:50-62
    ConvolutionOperator(const MultiResolutionAnalysis<D> &mra, GaussExp<1> &kernel, double prec)
            : MRA(mra)
            , buildPrec(prec) {
        std::vector<double> c, e;
        for (int i = 0; i < kernel.size(); i++) {
            c.push_back(kernel.getFunc(i).getCoef());
            e.push_back(kernel.getFunc(i).getExp()[0]);
        }
        h = mrx_convolution_create(mra.handle(), (int)c.size(), c.data(), e.data(), prec);
    }
    double getBuildPrec() const { return buildPrec; }
    const MultiResolutionAnalysis<D> &getMRA() const { return MRA; }
    bool isPeriodic() const { return false; }

protected:
    ConvolutionOperator(const MultiResolutionAnalysis<D> &mra, double prec)
            : MRA(mra)
            , buildPrec(prec) {}
    const MultiResolutionAnalysis<D> MRA;
    double buildPrec;
};

/// src/operators/PoissonOperator.cpp:40-55
class PoissonOperator final : public ConvolutionOperator<3> {
public:
    PoissonOperator(const MultiResolutionAnalysis<3> &mra, double prec)
            : ConvolutionOperator<3>(mra, prec) {
        h = mrx_poisson_create(mra.handle(), prec);
    }
    /// periodic worlds: src/operators/PoissonOperator.cpp:56-77
    PoissonOperator(const MultiResolutionAnalysis<3> &mra, double prec, int root, int reach)
            : ConvolutionOperator<3>(mra, prec) {
        h = mrx_poisson_create_reach(mra.handle(), prec, root, reach);
    }
};

/// src/operators/HelmholtzOperator.cpp:44-59
class HelmholtzOperator final : public ConvolutionOperator<3> {
public:
    HelmholtzOperator(const MultiResolutionAnalysis<3> &mra, double m, double prec)
            : ConvolutionOperator<3>(mra, prec)
            , mu(m) {
        h = mrx_helmholtz_create(mra.handle(), m, prec);
    }
    /// periodic worlds: src/operators/HelmholtzOperator.cpp:60-81
    HelmholtzOperator(const MultiResolutionAnalysis<3> &mra, double m, double prec, int root, int reach)
            : ConvolutionOperator<3>(mra, prec)
            , mu(m) {
        h = mrx_helmholtz_create_reach(mra.handle(), m, prec, root, reach);
    }
    double getMu() const { return mu; }

private:
    double mu;
};

/// IdentityConvolution(mra, prec): src/operators/IdentityConvolution.cpp:40-56 with IdentityKernel (IdentityKernel.h:40-47): one
/// narrow normalised Gaussian, exponent sqrt(1 / (prec / 10))
template <int D> class IdentityConvolution final : public ConvolutionOperator<D> {
public:
    IdentityConvolution(const MultiResolutionAnalysis<D> &mra, double prec)
            : ConvolutionOperator<D>(mra, prec) {
        const double expo = std::sqrt(1.0 / (prec / 10.0));
        const double coef = std::pow(expo / pi, D / 2.0);
        this->h = mrx_convolution_create(mra.handle(), 1, &coef, &expo, prec);
    }
};

template <int D> class DerivativeOperator : public MWOperatorBase {
    static_assert(D == 3, "the B200 path implements 3-dimensional operators only");

public:
    int getOrder() const { return order; }

protected:
    explicit DerivativeOperator(int ord)
            : order(ord) {}
    int order;
};

/// src/operators/ABGVOperator.cpp:46-74
template <int D> class ABGVOperator final : public DerivativeOperator<D> {
public:
    ABGVOperator(const MultiResolutionAnalysis<D> &mra, double a, double b)
            : DerivativeOperator<D>(1) {
        this->h = mrx_abgv_create(mra.handle(), a, b);
    }
};

/// src/operators/PHOperator.cpp:40-69 (order 1 or 2)
template <int D> class PHOperator final : public DerivativeOperator<D> {
public:
    PHOperator(const MultiResolutionAnalysis<D> &mra, int order)
            : DerivativeOperator<D>(order) {
        this->h = mrx_ph_create(mra.handle(), order);
    }
};

/// src/operators/BSOperator.cpp:40-66 (order 1, 2 or 3)
template <int D> class BSOperator final : public DerivativeOperator<D> {
public:
    BSOperator(const MultiResolutionAnalysis<D> &mra, int order)
            : DerivativeOperator<D>(order) {
        this->h = mrx_bs_create(mra.handle(), order);
    }
};

// ---- tree builders -----------------------------------------------------------------------------------------------------
/// build_grid(out, GaussExp | Gaussian): src/treebuilders/grid.cpp:78-123
template <int D> void build_grid(FunctionTree<D> &out, const GaussExp<D> &inp, int maxIter = -1) {
    b200::GaussArrays<D> a;
    b200::gauss_arrays<D>(inp, a);
    mrx_build_grid_gaussians(out.handle(), a.size(), a.coef.data(), a.expo.data(), a.pos.data(), a.power.data(), maxIter);
}
template <int D, typename T> void build_grid(FunctionTree<D, T> &out, const RepresentableFunction<D, T> &inp, int maxIter = -1) {
    b200::GaussArrays<D> a;
    if (!b200::gauss_arrays<D>(inp, a)) MRCPP_B200_ABORT("build_grid: only Gaussian functions know where they are visible on the B200 path");
    mrx_build_grid_gaussians(out.handle(), a.size(), a.coef.data(), a.expo.data(), a.pos.data(), a.power.data(), maxIter);
}
/// copy_grid / clear_grid: src/treebuilders/grid.cpp:150-166, :180-186
template <int D, typename T> void copy_grid(FunctionTree<D, T> &out, FunctionTree<D, T> &inp) { mrx_tree_copy_grid(out.handle(), inp.handle()); }
/// refine_grid(out, prec, absPrec) / refine_grid(out, scales): src/treebuilders/grid.cpp:271-302; returns the number of new nodes
template <int D, typename T> int refine_grid(FunctionTree<D, T> &out, double prec, bool absPrec = false) {
    return mrx_tree_refine_grid(out.handle(), prec, absPrec ? 1 : 0, 0);
}
template <int D, typename T> int refine_grid(FunctionTree<D, T> &out, int scales) {
    return scales > 0 ? mrx_tree_refine_grid(out.handle(), -1.0, 0, scales) : 0;
}
/// clear_grid(out): src/treebuilders/grid.cpp:180-186
template <int D, typename T> void clear_grid(FunctionTree<D, T> &out) { mrx_tree_clear_grid(out.handle()); }
/// build_grid(out, tree): extend the grid of `out` with the nodes of `inp` (src/treebuilders/grid.cpp:144-153)
template <int D, typename T> void build_grid(FunctionTree<D, T> &out, FunctionTree<D, T> &inp, int maxIter = -1) {
    if (maxIter >= 0) MRCPP_B200_ABORT("build_grid(out, tree, maxIter >= 0) is not on the B200 path");
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_tree_build_grid_from(out.handle(), inp.handle());
}

/// build_grid(out, FunctionTreeVector): union with the grids of all inputs (src/treebuilders/grid.cpp:170-178)
template <int D, typename T> void build_grid(FunctionTree<D, T> &out, FunctionTreeVector<D, T> &inp, int maxIter = -1) {
    if (maxIter >= 0) MRCPP_B200_ABORT("build_grid(out, trees, maxIter >= 0) is not on the B200 path");
    for (auto &t : inp) mrx_tree_build_grid_from(out.handle(), std::get<1>(t)->handle());
}

namespace b200 {
template <int D> double trampoline(const double r[3], void *user) {
    const auto &f = *static_cast<const std::function<double(const Coord<D> &)> *>(user);
    return f(Coord<D>{r[0], r[1], r[2]});
}
} // namespace b200

/// project(prec, out, f): src/treebuilders/project.cpp:85-104. The callback may be invoked from several OpenMP threads at
/// once, as in the reference (ProjectionCalculator runs inside TreeCalculator's parallel loop).
template <int D, typename T = double>
void project(double prec, FunctionTree<D, T> &out, std::function<T(const Coord<D> &r)> func, int maxIter = -1, bool absPrec = false) {
    if (maxIter >= 0 || absPrec) MRCPP_B200_ABORT("project: maxIter / absPrec variants are not on the B200 path");
    mrx_project_function(out.handle(), prec, &b200::trampoline<D>, &func, /*threads_ok=*/1, /*finalize=*/1);
}
/// Gaussian functions are projected on the device (quadrature, cvTransform, compression and norms as CUDA kernels); any other
/// RepresentableFunction goes through the callback projection above.
template <int D, typename T = double>
void project(double prec, FunctionTree<D, T> &out, RepresentableFunction<D, T> &inp, int maxIter = -1, bool absPrec = false) {
    if (maxIter >= 0 || absPrec) MRCPP_B200_ABORT("project: maxIter / absPrec variants are not on the B200 path");
    b200::GaussArrays<D> a;
    if (b200::gauss_arrays<D>(inp, a)) {
        mrx_project_gaussians_device(out.handle(), prec, a.size(), a.coef.data(), a.expo.data(), a.pos.data(), a.power.data(),
                                     /*build_grid=*/0);
        return;
    }
    std::function<T(const Coord<D> &)> f = [&inp](const Coord<D> &r) { return inp.evalf(r); };
    mrx_project_function(out.handle(), prec, &b200::trampoline<D>, &f, 1, 1);
}

namespace b200 {
/// work counters of the last apply of this thread (OperatorStatistics, src/operators/OperatorStatistics.cpp:83-106)
inline mrx_apply_stats &last_apply_stats() {
    static thread_local mrx_apply_stats st{};
    return st;
}
} // namespace b200

/// mrcpp::apply(prec, out, oper, inp, maxIter, absPrec): src/treebuilders/apply.cpp:68-93
template <int D, typename T>
void apply(double prec, FunctionTree<D, T> &out, ConvolutionOperator<D> &oper, FunctionTree<D, T> &inp, int maxIter = -1, bool absPrec = false) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_apply(prec, out.handle(), oper.handle(), inp.handle(), maxIter, absPrec ? 1 : 0, &b200::last_apply_stats());
}
namespace b200 {
/// One rank of a multi-GPU job (one process per GPU; the library creates the NCCL communicator). Rank 0 obtains the 128-byte
/// id with Comm::unique_id(), the host program ships it to every rank (MPI_Bcast in an MRCPP/MPI program), every rank
/// constructs its Comm. The reference has no equivalent inside one apply: it distributes whole trees over MPI ranks.
class Comm final {
public:
    static std::array<char, 128> unique_id() {
        std::array<char, 128> id{};
        mrx_comm_unique_id(id.data());
        return id;
    }
    Comm(int rank, int world, const std::array<char, 128> &id)
            : h(mrx_comm_create(rank, world, id.data())) {}
    ~Comm() { mrx_comm_destroy(h); }
    Comm(const Comm &) = delete;
    Comm &operator=(const Comm &) = delete;
    int rank() const { return mrx_comm_rank(h); }
    int size() const { return mrx_comm_size(h); }
    const mrx_comm *handle() const { return h; }

private:
    mrx_comm *h;
};
} // namespace b200

/// mrcpp::apply sharded over the GPUs of `comm` (collective: every rank calls it with identical arguments and ends with the
/// complete, bit-identical output tree)
template <int D, typename T>
void apply(double prec, FunctionTree<D, T> &out, ConvolutionOperator<D> &oper, FunctionTree<D, T> &inp, const b200::Comm &comm, int maxIter = -1,
           bool absPrec = false) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_apply_sharded(prec, out.handle(), oper.handle(), inp.handle(), maxIter, absPrec ? 1 : 0, comm.handle(), &b200::last_apply_stats());
}
/// mrcpp::apply(out, DerivativeOperator, inp, dir): src/treebuilders/apply.cpp:379-412
template <int D, typename T> void apply(FunctionTree<D, T> &out, DerivativeOperator<D> &oper, FunctionTree<D, T> &inp, int dir = -1) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_apply_derivative(out.handle(), oper.handle(), inp.handle(), dir, &b200::last_apply_stats());
}
/// mrcpp::add(prec, out, inp, maxIter, absPrec): src/treebuilders/add.cpp:41-70, from the grid `out` enters with; prec < 0 or
/// maxIter = 0: no refinement (the form mrcpp::divergence uses)
template <int D, typename T>
void add(double prec, FunctionTree<D, T> &out, FunctionTreeVector<D, T> &inp, int maxIter = -1, bool absPrec = false, bool conjugate = false) {
    (void)conjugate; // real trees
    std::vector<T> c;
    std::vector<mrx_tree *> h;
    for (auto &t : inp) {
        if (out.getMRA() != std::get<1>(t)->getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
        c.push_back(std::get<0>(t));
        h.push_back(std::get<1>(t)->handle());
    }
    mrx_tree_add_adaptive(prec, out.handle(), (int)h.size(), c.data(), h.data(), maxIter, absPrec ? 1 : 0);
}
template <int D, typename T>
void add(double prec, FunctionTree<D, T> &out, T a, FunctionTree<D, T> &inp_a, T b, FunctionTree<D, T> &inp_b, int maxIter = -1, bool absPrec = false,
         bool conjugate = false) {
    FunctionTreeVector<D, T> v;
    v.push_back(std::make_tuple(a, &inp_a));
    v.push_back(std::make_tuple(b, &inp_b));
    add(prec, out, v, maxIter, absPrec, conjugate);
}
/// copy_func(out, inp): src/treebuilders/grid.cpp:204-208 -- the function `inp` on the grid `out` enters with
template <int D, typename T> void copy_func(FunctionTree<D, T> &out, FunctionTree<D, T> &inp) {
    FunctionTreeVector<D, T> v;
    v.push_back(std::make_tuple(T(1.0), &inp));
    add(-1.0, out, v);
}
/// mrcpp::apply(prec, out, oper, inp, precTrees, maxIter, absPrec): src/treebuilders/apply.cpp:214-251 -- precision scaled per
/// output node by the largest norms of the precision trees (the coefficients of the vector entries are not used, as in the reference)
template <int D, typename T>
void apply(double prec, FunctionTree<D, T> &out, ConvolutionOperator<D> &oper, FunctionTree<D, T> &inp, FunctionTreeVector<D, T> &precTrees,
           int maxIter = -1, bool absPrec = false) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    std::vector<mrx_tree *> h;
    for (auto &t : precTrees) h.push_back(std::get<1>(t)->handle());
    mrx_apply_prec_trees(prec, out.handle(), oper.handle(), inp.handle(), (int)h.size(), h.data(), maxIter, absPrec ? 1 : 0, nullptr,
                         &b200::last_apply_stats());
}
/// mrcpp::apply_near_field / apply_far_field: src/treebuilders/apply.cpp:294-342 (periodic worlds: contributions from inside /
/// outside the unit cell only)
template <int D, typename T>
void apply_near_field(double prec, FunctionTree<D, T> &out, ConvolutionOperator<D> &oper, FunctionTree<D, T> &inp, int maxIter = -1,
                      bool absPrec = false) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_apply_unit_cell(1, prec, out.handle(), oper.handle(), inp.handle(), maxIter, absPrec ? 1 : 0, &b200::last_apply_stats());
}
template <int D, typename T>
void apply_far_field(double prec, FunctionTree<D, T> &out, ConvolutionOperator<D> &oper, FunctionTree<D, T> &inp, int maxIter = -1,
                     bool absPrec = false) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_apply_unit_cell(0, prec, out.handle(), oper.handle(), inp.handle(), maxIter, absPrec ? 1 : 0, &b200::last_apply_stats());
}
/// mrcpp::multiply(prec, out, inp, maxIter, absPrec, useMaxNorms, conjugate): src/treebuilders/multiply.cpp:104-136
template <int D, typename T>
void multiply(double prec, FunctionTree<D, T> &out, FunctionTreeVector<D, T> &inp, int maxIter = -1, bool absPrec = false, bool useMaxNorms = false,
              bool conjugate = false) {
    (void)conjugate; // real trees
    std::vector<T> c;
    std::vector<mrx_tree *> h;
    for (auto &t : inp) {
        if (out.getMRA() != std::get<1>(t)->getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
        c.push_back(std::get<0>(t));
        h.push_back(std::get<1>(t)->handle());
    }
    mrx_tree_multiply(prec, out.handle(), (int)h.size(), c.data(), h.data(), maxIter, absPrec ? 1 : 0, useMaxNorms ? 1 : 0);
}
template <int D, typename T>
void multiply(double prec, FunctionTree<D, T> &out, T c, FunctionTree<D, T> &inp_a, FunctionTree<D, T> &inp_b, int maxIter = -1, bool absPrec = false,
              bool useMaxNorms = false, bool conjugate = false) {
    FunctionTreeVector<D, T> v;
    v.push_back(std::make_tuple(c, &inp_a));
    v.push_back(std::make_tuple(T(1.0), &inp_b));
    multiply(prec, out, v, maxIter, absPrec, useMaxNorms, conjugate);
}
/// mrcpp::square(prec, out, inp): src/treebuilders/multiply.cpp (out = inp * inp)
template <int D, typename T> void square(double prec, FunctionTree<D, T> &out, FunctionTree<D, T> &inp, int maxIter = -1, bool absPrec = false) {
    FunctionTreeVector<D, T> v;
    v.push_back(std::make_tuple(T(1.0), &inp));
    v.push_back(std::make_tuple(T(1.0), &inp));
    multiply(prec, out, v, maxIter, absPrec);
}

/// mrcpp::power(prec, out, inp, p): src/treebuilders/multiply.cpp:211-234
template <int D, typename T> void power(double prec, FunctionTree<D, T> &out, FunctionTree<D, T> &inp, double p, int maxIter = -1, bool absPrec = false) {
    if (out.getMRA() != inp.getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    mrx_tree_power(prec, out.handle(), inp.handle(), p, maxIter, absPrec ? 1 : 0);
}
/// mrcpp::dot(prec, out, inp_a, inp_b, maxIter, absPrec): src/treebuilders/multiply.cpp:253-271 -- out = sum_d a_d b_d f_d g_d, every
/// product on the grid of `out` refined with the MultiplicationAdaptor (useMaxNorms), the sum on the union of the product grids
template <int D, typename T>
void dot(double prec, FunctionTree<D, T> &out, FunctionTreeVector<D, T> &inp_a, FunctionTreeVector<D, T> &inp_b, int maxIter = -1, bool absPrec = false) {
    if (inp_a.size() != inp_b.size()) MRCPP_B200_ABORT("Input length mismatch");
    FunctionTreeVector<D, T> tmp_vec;
    for (size_t d = 0; d < inp_a.size(); d++) {
        auto *out_d = new FunctionTree<D, T>(out.getMRA());
        build_grid(*out_d, out);
        multiply(prec, *out_d, T(1.0), get_func(inp_a, (int)d), get_func(inp_b, (int)d), maxIter, absPrec, true);
        tmp_vec.push_back(std::make_tuple(get_coef(inp_a, (int)d) * get_coef(inp_b, (int)d), out_d));
    }
    build_grid(out, tmp_vec);
    add(-1.0, out, tmp_vec, 0);
    clear(tmp_vec, true);
}

/// mrcpp::gradient(oper, inp): src/treebuilders/apply.cpp:444-452 (the caller owns the trees: clear(vec, true))
template <int D, typename T> FunctionTreeVector<D, T> gradient(DerivativeOperator<D> &oper, FunctionTree<D, T> &inp) {
    FunctionTreeVector<D, T> out;
    for (int d = 0; d < D; d++) {
        auto *grad_d = new FunctionTree<D, T>(inp.getMRA());
        apply(*grad_d, oper, inp, d);
        out.push_back(std::make_tuple(T(1.0), grad_d));
    }
    return out;
}
/// mrcpp::divergence(out, oper, inp): src/treebuilders/apply.cpp:514-530
template <int D, typename T> void divergence(FunctionTree<D, T> &out, DerivativeOperator<D> &oper, FunctionTreeVector<D, T> &inp) {
    if ((int)inp.size() != D) MRCPP_B200_ABORT("Dimension mismatch");
    for (auto &t : inp)
        if (out.getMRA() != std::get<1>(t)->getMRA()) MRCPP_B200_ABORT("Incompatible MRA");
    FunctionTreeVector<D, T> tmp_vec;
    for (int d = 0; d < D; d++) {
        auto *out_d = new FunctionTree<D, T>(get_func(inp, d).getMRA());
        apply(*out_d, oper, get_func(inp, d), d);
        tmp_vec.push_back(std::make_tuple(get_coef(inp, d), out_d));
    }
    build_grid(out, tmp_vec);
    add(-1.0, out, tmp_vec, 0); // addition on the union grid
    clear(tmp_vec, true);
}

/// mrcpp::dot(bra, ket): src/treebuilders/multiply.cpp:286-318
template <int D, typename T> T dot(FunctionTree<D, T> &bra, FunctionTree<D, T> &ket) {
    if (bra.getMRA() != ket.getMRA()) MRCPP_B200_ABORT("Trees not compatible");
    return mrx_dot(bra.handle(), ket.handle());
}

// ---- Printer (src/utils/Printer.h:61-133) ------------------------------------------------------------------------------
class Printer final {
public:
    static void init(int level = 0, int rank = 0, int size = 1, const char * /*file*/ = nullptr) {
        printLevel = level;
        printRank = rank;
        printSize = size;
        if (rank != 0) printLevel = -1; // only rank 0 prints to the screen (Printer.cpp: other ranks go to files)
        *out << std::scientific << std::setprecision(printPrec);
    }
    static void setScientific() { *out << std::scientific; }
    static void setFixed() { *out << std::fixed; }
    static int setWidth(int i) {
        int old = printWidth;
        printWidth = i;
        return old;
    }
    static int setPrecision(int i) {
        int old = printPrec;
        printPrec = i;
        *out << std::setprecision(i);
        return old;
    }
    static int setPrintLevel(int i) {
        int old = printLevel;
        printLevel = i;
        return old;
    }
    static int getWidth() { return printWidth; }
    static int getPrecision() { return printPrec; }
    static int getPrintLevel() { return printLevel; }
    static void setOutputStream(std::ostream &o) { out = &o; }
    static bool active(int level) { return level <= printLevel; }

    static inline std::ostream *out = &std::cout; // public like the reference's (the print macros below write to it)

private:
    static inline int printWidth = 60;
    static inline int printLevel = -1;
    static inline int printPrec = 12;
    static inline int printRank = 0;
    static inline int printSize = 1;
};

// print macros of src/utils/Printer.h:138-190
#define println(level, STR)                                                                                                    \
    {                                                                                                                          \
        if (level <= mrcpp::Printer::getPrintLevel()) *mrcpp::Printer::out << STR << std::endl;                                \
    }
#define printout(level, STR)                                                                                                   \
    {                                                                                                                          \
        if (level <= mrcpp::Printer::getPrintLevel()) *mrcpp::Printer::out << STR;                                             \
    }
#define MSG_INFO(STR)                                                                                                          \
    { *mrcpp::Printer::out << "Info: " << __FILE__ << ": " << __func__ << "(), line " << __LINE__ << ": " << STR << std::endl; }
#define MSG_WARN(STR)                                                                                                          \
    { *mrcpp::Printer::out << "Warning: " << __func__ << "(), line " << __LINE__ << ": " << STR << std::endl; }
#define MSG_ERROR(STR)                                                                                                         \
    { *mrcpp::Printer::out << "Error: " << __func__ << "(), line " << __LINE__ << ": " << STR << std::endl; }
#define MSG_ABORT(STR)                                                                                                         \
    {                                                                                                                          \
        *mrcpp::Printer::out << "Error: " << __FILE__ << ": " << __func__ << "(), line " << __LINE__ << ": " << STR << std::endl; \
        abort();                                                                                                               \
    }
#define NOT_IMPLEMENTED_ABORT                                                                                                  \
    {                                                                                                                          \
        *mrcpp::Printer::out << "Error: Not implemented, " << __FILE__ ", " << __func__ << "(), line " << __LINE__ << std::endl; \
        abort();                                                                                                               \
    }

namespace print {
inline void separator(int level, const char &c, int newlines = 0) {
    if (!Printer::active(level)) return;
    (*Printer::out) << std::string(Printer::getWidth(), c) << std::endl;
    for (int i = 0; i < newlines; i++) (*Printer::out) << std::endl;
}
inline void header(int level, const std::string &txt, int newlines = 0, const char &c = '=') {
    if (!Printer::active(level)) return;
    const int len = (int)txt.size();
    separator(level, c);
    (*Printer::out) << std::string(std::max(0, (Printer::getWidth() - len) / 2), ' ') << txt << std::endl;
    separator(level, '-', newlines);
}
inline void footer(int level, const Timer &timer, int newlines = 0, const char &c = '=') {
    if (!Printer::active(level)) return;
    std::ostringstream o;
    o << std::fixed << std::setprecision(5) << "Wall time: " << std::scientific << timer.elapsed() << " sec";
    const int len = (int)o.str().size();
    separator(level, '-');
    (*Printer::out) << std::string(std::max(0, (Printer::getWidth() - len) / 2), ' ') << o.str() << std::endl;
    separator(level, c, newlines);
}
inline void environment(int level) {
    if (!Printer::active(level)) return;
    b200::ensure_init();
    separator(level, '-', 1);
    (*Printer::out) << " MRCPP API on " << mrx_version() << std::endl;
    (*Printer::out) << " CUDA devices visible : " << mrx_device_count() << std::endl;
    (*Printer::out) << " device of this process: " << b200::device_ref() << (b200::device_ref() < 0 ? " (host only: hot-path calls abort)" : "") << std::endl;
    (*Printer::out) << std::endl;
    separator(level, '-', 1);
}
inline void memory(int level, const std::string &txt) {
    if (!Printer::active(level)) return;
    long pages = 0, rss = 0;
    if (FILE *f = std::fopen("/proc/self/statm", "r")) {
        if (std::fscanf(f, "%ld %ld", &pages, &rss) != 2) rss = 0;
        std::fclose(f);
    }
    const double mb = rss * 4096.0 / (1024.0 * 1024.0);
    std::ostringstream o;
    o << " " << txt;
    const int pad = std::max(1, Printer::getWidth() - (int)o.str().size() - 16);
    (*Printer::out) << o.str() << std::string(pad, ' ') << std::fixed << std::setprecision(2) << std::setw(10) << mb << " (MB)" << std::scientific
                   << std::setprecision(Printer::getPrecision()) << std::endl;
}
inline void value(int level, const std::string &txt, double v, const std::string &unit = "", int p = -1, bool sci = true) {
    if (!Printer::active(level)) return;
    if (p < 0) p = Printer::getPrecision();
    std::ostringstream o;
    o << " " << std::left << std::setw(30) << txt << std::right << std::setw(8) << unit << " ";
    if (sci) o << std::scientific;
    else o << std::fixed;
    o << std::setprecision(p) << std::setw(std::max(p + 8, Printer::getWidth() - 41)) << v;
    (*Printer::out) << o.str() << std::endl;
}
inline void time(int level, const std::string &txt, const Timer &timer) { value(level, txt, timer.elapsed(), "(sec)", 5); }
inline void tree(int level, const std::string &txt, int n, int m, double t) {
    if (!Printer::active(level)) return;
    std::ostringstream o;
    o << " " << std::left << std::setw(26) << txt << std::right << std::setw(8) << n << " nds " << std::setw(8) << m << " kB " << std::scientific
      << std::setprecision(2) << std::setw(9) << t << " sec";
    (*Printer::out) << o.str() << std::endl;
}
template <int D, typename T> void tree(int level, const std::string &txt, const MWTree<D, T> &tr, const Timer &timer) {
    tree(level, txt, tr.getNNodes(), tr.getSizeNodes(), timer.elapsed());
}
} // namespace print

} // namespace mrcpp
