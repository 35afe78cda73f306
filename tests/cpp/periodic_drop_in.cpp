// Periodic operator application and the apply with precision trees through the C++ MRCPP mirror (include/MRCPP/ over the C ABI),
// written like the reference's own tests: tests/operators/poisson_operator.cpp:156-199 ("Apply Periodic Poisson' operator": a
// cosine source whose periodic Poisson solution is known) and tests/operators/helmholtz_operator.cpp:200-241 (near field + far
// field = whole), here on the unit cell [-1, 1]^3 without a scaling factor (period 2), plus apply(prec, out, oper, inp, precTrees).
// Prints "key value" lines; tests/test_cpp_mirror.py (oracle backend) and tests/test_zz3_gpu_cpp_mirror.py (device) check them.
#include "MRCPP/MWFunctions"
#include "MRCPP/MWOperators"
#include "MRCPP/Printer"

constexpr int D = 3;
using mrcpp::pi;

int main() {
    mrcpp::Printer::init(-1);
    const double proj_prec = 1.0e-4, apply_prec = 1.0e-3, build_prec = 1.0e-3;
    auto world = mrcpp::BoundingBox<D>(0, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2}, std::array<double, D>{1.0, 1.0, 1.0}, true);
    mrcpp::MultiResolutionAnalysis<D> MRA(world, mrcpp::InterpolatingBasis(5), 25);
    std::printf("periodic %d\n", MRA.getWorldBox().isPeriodic() ? 1 : 0);

    const int oper_root = 0, oper_reach = 9;
    mrcpp::PoissonOperator P(MRA, build_prec, oper_root, oper_reach);
    mrcpp::HelmholtzOperator H(MRA, 4.3, build_prec, oper_root, oper_reach);
    std::printf("poisson_terms %d\nhelmholtz_terms %d\n", P.size(), H.size());

    // -lap u = 4 pi rho with u = cos(pi x) cos(pi y) cos(pi z) + a cos(2 pi x) cos(pi y) cos(3 pi z), a = 0.7 * 4 pi / (14 pi^2)
    auto source = [](const mrcpp::Coord<D> &r) {
        return 3.0 * pi * pi * std::cos(pi * r[0]) * std::cos(pi * r[1]) * std::cos(pi * r[2]) / (4.0 * pi) +
               0.7 * std::cos(2.0 * pi * r[0]) * std::cos(pi * r[1]) * std::cos(3.0 * pi * r[2]);
    };
    auto exact = [](double x, double y, double z) {
        return std::cos(pi * x) * std::cos(pi * y) * std::cos(pi * z) +
               0.7 * 4.0 * pi / (14.0 * pi * pi) * std::cos(2.0 * pi * x) * std::cos(pi * y) * std::cos(3.0 * pi * z);
    };
    mrcpp::FunctionTree<D> source_tree(MRA);
    mrcpp::project<D, double>(proj_prec, source_tree, source);
    std::printf("source_nodes %d\n", source_tree.getNNodes());

    mrcpp::FunctionTree<D> sol_tree(MRA), in_tree(MRA), out_tree(MRA);
    mrcpp::apply(apply_prec, sol_tree, P, source_tree);
    mrcpp::apply_near_field(apply_prec, in_tree, P, source_tree);
    mrcpp::apply_far_field(apply_prec, out_tree, P, source_tree);
    // the constant Fourier mode of the periodised kernel depends on the reach: compare differences of function values
    const double u00 = sol_tree.evalf({0.0, 0.0, 0.0}), u10 = sol_tree.evalf({1.0, 0.0, 0.0}), uh = sol_tree.evalf({0.5, 0.25, 0.0});
    std::printf("sol_nodes %d\n", sol_tree.getNNodes());
    std::printf("sol_diff_0_1 %.12g\nexact_diff_0_1 %.12g\n", u00 - u10, exact(0, 0, 0) - exact(1, 0, 0)); // (1,0,0) wraps to (-1,0,0)
    std::printf("sol_diff_0_h %.12g\nexact_diff_0_h %.12g\n", u00 - uh, exact(0, 0, 0) - exact(0.5, 0.25, 0));
    const double w = in_tree.evalf({0.3, -0.2, 0.6}) + out_tree.evalf({0.3, -0.2, 0.6});
    std::printf("near_plus_far %.12g\nwhole %.12g\n", w, sol_tree.evalf({0.3, -0.2, 0.6}));

    mrcpp::FunctionTree<D> hel_tree(MRA);
    mrcpp::apply(apply_prec, hel_tree, H, source_tree);
    std::printf("helmholtz_nodes %d\nhelmholtz_sqnorm %.12g\n", hel_tree.getNNodes(), hel_tree.getSquareNorm());

    // apply with precision trees: the precision scaled by the largest norms of the source changes the grid
    mrcpp::FunctionTreeVector<D> precTrees;
    precTrees.push_back(std::make_tuple(1.0, &source_tree));
    mrcpp::FunctionTree<D> scaled_tree(MRA);
    mrcpp::apply(apply_prec, scaled_tree, P, source_tree, precTrees);
    std::printf("scaled_nodes %d\nscaled_diff_0_1 %.12g\n", scaled_tree.getNNodes(),
                scaled_tree.evalf({0.0, 0.0, 0.0}) - scaled_tree.evalf({1.0, 0.0, 0.0}));
    std::printf("done 1\n");
    return 0;
}
