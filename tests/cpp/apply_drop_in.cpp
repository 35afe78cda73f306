// The hot path through the C++ MRCPP mirror (include/MRCPP/ over the C ABI), written the way a program is written against
// MRCPP itself: Poisson apply (the case of the reference's examples/poisson.cpp and tests/operators/poisson_operator.cpp),
// the hydrogen 1s fixed point of the Helmholtz operator (tests/operators/helmholtz_operator.cpp) and an ABGV derivative of a
// Gaussian against the projection of its analytic derivative (tests/operators/derivative_operator.cpp pattern).
// Prints "key value" lines; tests/test_zz3_gpu_cpp_mirror.py checks them against the analytic answers and against the same
// cases run through the Python mirror. Needs a CUDA device: without one the first projection aborts (no CPU fallback).
#include "MRCPP/Gaussians"
#include "MRCPP/MWFunctions"
#include "MRCPP/MWOperators"
#include "MRCPP/Printer"
#include "MRCPP/Timer"

constexpr int D = 3;

static void poisson_case() {
    const int order = 7;
    const double prec = 1.0e-5;
    mrcpp::BoundingBox<D> world(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(order), 25);

    const double beta = 100.0;
    mrcpp::GaussFunc<D> rho(beta, std::pow(beta / mrcpp::pi, 1.5), mrcpp::Coord<D>{mrcpp::pi / 3.0, mrcpp::pi / 3.0, mrcpp::pi / 3.0});
    mrcpp::PoissonOperator P(MRA, prec);

    mrcpp::FunctionTree<D> f_tree(MRA);
    mrcpp::build_grid(f_tree, rho);
    mrcpp::project(prec, f_tree, rho);

    mrcpp::Timer t;
    mrcpp::FunctionTree<D> g_tree(MRA);
    mrcpp::apply(prec, g_tree, P, f_tree);
    t.stop();

    const auto &st = mrcpp::b200::last_apply_stats();
    std::printf("poisson_terms %d\n", P.size());
    std::printf("poisson_f_nodes %d\npoisson_g_nodes %d\n", f_tree.getNNodes(), g_tree.getNNodes());
    std::printf("poisson_f_integral %.17g\npoisson_f_sqnorm %.17g\n", f_tree.integrate(), f_tree.getSquareNorm());
    std::printf("poisson_g_integral %.17g\npoisson_g_sqnorm %.17g\n", g_tree.integrate(), g_tree.getSquareNorm());
    std::printf("poisson_energy %.17g\npoisson_analytic %.17g\n", mrcpp::dot(g_tree, f_tree), rho.calcCoulombEnergy(rho));
    std::printf("poisson_tuples %lld\npoisson_calc_nodes %lld\npoisson_launches %lld\n", st.f_applied, st.g_nodes, st.kernel_launches);
    mrcpp::print::tree(0, "g_tree", g_tree, t);

    // fixed-grid variant: the same grid, maxIter = 0 (no refinement)
    mrcpp::FunctionTree<D> h_tree(MRA);
    mrcpp::copy_grid(h_tree, g_tree);
    mrcpp::apply(prec, h_tree, P, f_tree, 0);
    std::printf("poisson_fixed_grid_nodes %d\npoisson_fixed_grid_energy %.17g\n", h_tree.getNNodes(), mrcpp::dot(h_tree, f_tree));

    // TopDown / BottomUp round trip leaves the norm where it was
    g_tree.mwTransform(mrcpp::TopDown);
    g_tree.mwTransform(mrcpp::BottomUp);
    g_tree.calcSquareNorm();
    std::printf("poisson_g_sqnorm_roundtrip %.17g\n", g_tree.getSquareNorm());
}

static void helmholtz_case() {
    const double proj_prec = 3.0e-3, apply_prec = 3.0e-2, build_prec = 3.0e-3;
    mrcpp::BoundingBox<D> world(-5, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(5), 25);
    const double c = 1.0 / std::sqrt(mrcpp::pi);
    auto psi = [c](const mrcpp::Coord<D> &r) -> double { return c * std::exp(-std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2])); };
    auto vpsi = [c](const mrcpp::Coord<D> &r) -> double {
        const double x = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        return -c * std::exp(-x) / x;
    };
    mrcpp::FunctionTree<D> psi_tree(MRA), vpsi_tree(MRA), out(MRA);
    mrcpp::project<D, double>(proj_prec, psi_tree, psi);
    mrcpp::project<D, double>(proj_prec, vpsi_tree, vpsi);
    mrcpp::HelmholtzOperator H(MRA, 1.0, build_prec); // mu = sqrt(-2 E), E = -1/2
    mrcpp::copy_grid(out, psi_tree);
    mrcpp::apply(apply_prec, out, H, vpsi_tree);
    out.rescale(-1.0 / (2.0 * mrcpp::pi));
    std::printf("helmholtz_terms %d\nhelmholtz_psi_sqnorm %.17g\n", H.size(), psi_tree.getSquareNorm());
    std::printf("helmholtz_out_norm %.17g\nhelmholtz_overlap %.17g\n", std::sqrt(out.getSquareNorm()), mrcpp::dot(out, psi_tree));
}

static void derivative_case() {
    const int order = 7;
    const double prec = 1.0e-5;
    mrcpp::BoundingBox<D> world(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(order), 25);
    const double beta = 20.0, alpha = std::pow(beta / mrcpp::pi, 1.5);
    const mrcpp::Coord<D> pos{0.3, -0.4, 0.5};
    mrcpp::GaussFunc<D> f(beta, alpha, pos);
    mrcpp::ABGVOperator<D> diff(MRA, 0.5, 0.5);
    mrcpp::FunctionTree<D> f_tree(MRA);
    mrcpp::build_grid(f_tree, f);
    mrcpp::project(prec, f_tree, f);
    for (int dir = 0; dir < D; dir++) {
        // d/dx_dir of alpha exp(-beta |r - pos|^2) = -2 beta alpha (x_dir - pos_dir) exp(...)
        std::array<int, D> power{0, 0, 0};
        power[dir] = 1;
        mrcpp::GaussFunc<D> df(beta, -2.0 * beta * alpha, pos, power);
        mrcpp::FunctionTree<D> df_tree(MRA), dg_tree(MRA);
        mrcpp::build_grid(df_tree, df);
        mrcpp::project(prec, df_tree, df);
        mrcpp::apply(dg_tree, diff, f_tree, dir);
        const double gg = mrcpp::dot(dg_tree, dg_tree), gf = mrcpp::dot(dg_tree, df_tree), ff = mrcpp::dot(df_tree, df_tree);
        std::printf("derivative_%d_nodes %d\nderivative_%d_sqnorm %.17g\nderivative_%d_rel_err %.17g\n", dir, dg_tree.getNNodes(), dir, gg, dir,
                    std::sqrt(std::abs(gg - 2.0 * gf + ff) / ff));
    }
}

static void divergence_case() {
    // gradient and divergence (apply.h:51,55): div grad f of a Gaussian against the projection of its analytic Laplacian
    // (4 beta^2 r^2 - 6 beta) f, and <div grad f | f> = -|grad f|^2 (the ABGV operator with a = b = 1/2 is antisymmetric)
    const int order = 7;
    const double prec = 1.0e-5;
    mrcpp::BoundingBox<D> world(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(order), 25);
    const double beta = 20.0, alpha = std::pow(beta / mrcpp::pi, 1.5);
    const mrcpp::Coord<D> pos{0.3, -0.4, 0.5};
    mrcpp::GaussFunc<D> f(beta, alpha, pos);
    mrcpp::ABGVOperator<D> diff(MRA, 0.5, 0.5);
    mrcpp::FunctionTree<D> f_tree(MRA), lap_tree(MRA), ana_tree(MRA), err_tree(MRA);
    mrcpp::build_grid(f_tree, f);
    mrcpp::project(prec, f_tree, f);
    auto grad = mrcpp::gradient(diff, f_tree);
    double grad_sq = 0.0;
    for (int d = 0; d < D; d++) grad_sq += mrcpp::get_func(grad, d).getSquareNorm();
    // point values of the gradient against the analytic ones (tests/operators/derivative_operator.cpp:417-453 pattern)
    const mrcpp::Coord<D> r{0.45, -0.3, 0.35};
    double worst = 0.0;
    for (int d = 0; d < D; d++) {
        const double ana = -2.0 * beta * (r[d] - pos[d]) * f.evalf(r);
        worst = std::max(worst, std::abs(mrcpp::get_func(grad, d).evalf_precise(r) - ana) / std::abs(ana));
    }
    std::printf("gradient_point_rel_err %.17g\nfunction_point_rel_err %.17g\n", worst, std::abs(f_tree.evalf_precise(r) - f.evalf(r)) / f.evalf(r));
    mrcpp::divergence(lap_tree, diff, grad);
    mrcpp::clear(grad, true);
    auto lap = [&](const mrcpp::Coord<D> &r) -> double {
        double r2 = 0.0;
        for (int d = 0; d < D; d++) r2 += (r[d] - pos[d]) * (r[d] - pos[d]);
        return (4.0 * beta * beta * r2 - 6.0 * beta) * alpha * std::exp(-beta * r2);
    };
    mrcpp::copy_grid(ana_tree, lap_tree);
    mrcpp::project<D, double>(-1.0, ana_tree, lap); // on the grid of the numerical result, no refinement
    mrcpp::build_grid(err_tree, lap_tree);
    mrcpp::add(-1.0, err_tree, 1.0, lap_tree, -1.0, ana_tree);
    std::printf("divergence_nodes %d\ndivergence_grad_sqnorm %.17g\ndivergence_overlap %.17g\n", lap_tree.getNNodes(), grad_sq,
                mrcpp::dot(lap_tree, f_tree));
    std::printf("divergence_integral %.17g\ndivergence_sqnorm %.17g\ndivergence_rel_err %.17g\n", lap_tree.integrate(), lap_tree.getSquareNorm(),
                std::sqrt(err_tree.getSquareNorm() / ana_tree.getSquareNorm()));
}

static void addition_case() {
    // adaptive sum of three projected Gaussians, g = f_1 - 2 f_2 + 3 f_3 (the case of the reference's examples/addition.cpp)
    const int order = 5;
    const double prec = 1.0e-4;
    mrcpp::BoundingBox<D> world(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(order), 25);
    const double beta = 20.0, alpha = std::pow(beta / mrcpp::pi, 1.5);
    mrcpp::GaussFunc<D> f1(beta, alpha, mrcpp::Coord<D>{0.0, 0.0, 0.1}), f2(beta, alpha, mrcpp::Coord<D>{0.0, 0.0, -0.1}),
        f3(beta, alpha, mrcpp::Coord<D>{0.0, 0.0, 0.3});
    mrcpp::FunctionTree<D> t1(MRA), t2(MRA), t3(MRA), g(MRA);
    mrcpp::project(prec, t1, f1);
    mrcpp::project(prec, t2, f2);
    mrcpp::project(prec, t3, f3);
    mrcpp::FunctionTreeVector<D> vec;
    vec.push_back(std::make_tuple(1.0, &t1));
    vec.push_back(std::make_tuple(-2.0, &t2));
    vec.push_back(std::make_tuple(3.0, &t3));
    mrcpp::add(prec, g, vec);
    const mrcpp::Coord<D> r{0.1, -0.05, 0.2};
    const double ana = f1.evalf(r) - 2.0 * f2.evalf(r) + 3.0 * f3.evalf(r);
    std::printf("addition_nodes %d\naddition_integral %.17g\naddition_sqnorm %.17g\naddition_point_rel_err %.17g\n", g.getNNodes(), g.integrate(),
                g.getSquareNorm(), std::abs(g.evalf_precise(r) - ana) / std::abs(ana));
    std::printf("addition_overlap_1 %.17g\naddition_overlap_expected %.17g\n", mrcpp::dot(g, t1),
                mrcpp::dot(t1, t1) - 2.0 * mrcpp::dot(t2, t1) + 3.0 * mrcpp::dot(t3, t1));
}

static void multiplication_case() {
    // h = f * g of two projected Gaussians against the analytic product Gaussian (the case of the reference's
    // examples/multiplication.cpp): exponent 2 beta at the midpoint, prefactor alpha^2 exp(-beta |d|^2 / 2)
    const int order = 5;
    const double prec = 1.0e-4;
    mrcpp::BoundingBox<D> world(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(order), 25);
    const double beta = 20.0, alpha = std::pow(beta / mrcpp::pi, 1.5);
    const mrcpp::Coord<D> f_pos{0.0, 0.0, 0.17}, g_pos{0.0, 0.0, -0.1};
    mrcpp::GaussFunc<D> f(beta, alpha, f_pos), g(beta, alpha, g_pos);
    double dist_2 = 0.0;
    mrcpp::Coord<D> p_pos{};
    for (int d = 0; d < D; d++) {
        dist_2 += (f_pos[d] - g_pos[d]) * (f_pos[d] - g_pos[d]);
        p_pos[d] = 0.5 * (f_pos[d] + g_pos[d]);
    }
    mrcpp::GaussFunc<D> prod(2.0 * beta, alpha * alpha * std::exp(-beta * 0.5 * dist_2), p_pos);
    mrcpp::FunctionTree<D> f_tree(MRA), g_tree(MRA), h_tree(MRA), sq_tree(MRA), prod_tree(MRA), err_tree(MRA);
    mrcpp::project(prec, f_tree, f);
    mrcpp::project(prec, g_tree, g);
    mrcpp::project(prec / 1000, prod_tree, prod);
    mrcpp::multiply(prec, h_tree, 1.0, f_tree, g_tree);
    mrcpp::square(prec, sq_tree, f_tree);
    mrcpp::build_grid(err_tree, h_tree);
    mrcpp::build_grid(err_tree, prod_tree);
    mrcpp::add(-1.0, err_tree, 1.0, h_tree, -1.0, prod_tree);
    const double ana_int = alpha * alpha * std::exp(-beta * 0.5 * dist_2) * std::pow(mrcpp::pi / (2.0 * beta), 1.5);
    const mrcpp::Coord<D> r{0.05, -0.1, 0.1};
    std::printf("multiplication_nodes %d\nmultiplication_integral %.17g\nmultiplication_integral_analytic %.17g\n", h_tree.getNNodes(),
                h_tree.integrate(), ana_int);
    std::printf("multiplication_rel_err %.17g\nmultiplication_point_rel_err %.17g\n", std::sqrt(err_tree.getSquareNorm() / prod_tree.getSquareNorm()),
                std::abs(h_tree.evalf_precise(r) - prod.evalf(r)) / prod.evalf(r));
    // <f | f> = integral of f^2
    std::printf("square_integral %.17g\nsquare_expected %.17g\n", sq_tree.integrate(), mrcpp::dot(f_tree, f_tree));
    // dot of two tree vectors (multiply.cpp:253-271): 2 f g - f f, integrated
    mrcpp::FunctionTreeVector<D> va, vb;
    va.push_back(std::make_tuple(2.0, &f_tree));
    vb.push_back(std::make_tuple(1.0, &g_tree));
    va.push_back(std::make_tuple(-1.0, &f_tree));
    vb.push_back(std::make_tuple(1.0, &f_tree));
    mrcpp::FunctionTree<D> dot_tree(MRA);
    mrcpp::dot(prec, dot_tree, va, vb);
    std::printf("dot_vectors_integral %.17g\ndot_vectors_expected %.17g\n", dot_tree.integrate(), 2.0 * mrcpp::dot(f_tree, g_tree) - mrcpp::dot(f_tree, f_tree));
}

int main(int argc, char **argv) {
    // argv[1]: print level (default -1); argv[2]: "core" (the apply path), "algebra" (its callers) or "all" (default)
    mrcpp::Printer::init(argc > 1 ? std::atoi(argv[1]) : -1);
    const std::string what = argc > 2 ? argv[2] : "all";
    mrcpp::print::environment(0);
    if (what != "algebra") {
        poisson_case();
        helmholtz_case();
        derivative_case();
    }
    if (what != "core") {
        divergence_case();
        addition_case();
        multiplication_case();
    }
    std::printf("done 1\n");
    return 0;
}
