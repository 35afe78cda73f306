// Host-side part of the C++ MRCPP mirror (include/MRCPP/): objects, getters, analytic functions, operator construction and
// grid building need no CUDA device. Prints "key value" lines that tests/test_cpp_mirror.py compares with the Python mirror
// of the same C ABI. Written against the reference's public names (api/MWFunctions, api/MWOperators, api/Gaussians).
#include <algorithm>
#include <sstream>

#include "MRCPP/Gaussians"
#include "MRCPP/MWFunctions"
#include "MRCPP/MWOperators"
#include "MRCPP/Printer"
#include "MRCPP/Timer"

int main() {
    constexpr int D = 3;
    mrcpp::Timer timer;
    std::ostringstream log;
    mrcpp::Printer::init(0);
    mrcpp::Printer::setOutputStream(log);
    mrcpp::print::header(0, "host api");

    mrcpp::BoundingBox<D> world(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2});
    mrcpp::InterpolatingBasis basis(7);
    mrcpp::MultiResolutionAnalysis<D> MRA(world, basis, 25);
    std::printf("order %d\nroot_scale %d\nmax_scale %d\nlower %.17g\nupper %.17g\n", MRA.getOrder(), MRA.getRootScale(), MRA.getMaxScale(),
                MRA.getWorldBox().getLowerBound(0), MRA.getWorldBox().getUpperBound(2));

    const double beta = 100.0;
    mrcpp::GaussFunc<D> f(beta, std::pow(beta / mrcpp::pi, 1.5), mrcpp::Coord<D>{mrcpp::pi / 3.0, mrcpp::pi / 3.0, mrcpp::pi / 3.0});
    mrcpp::GaussFunc<D> g(50.0, std::pow(50.0 / mrcpp::pi, 1.5), mrcpp::Coord<D>{0.1, -0.2, 0.3});
    std::printf("self_energy %.17g\npair_energy %.17g\nevalf %.17g\n", f.calcCoulombEnergy(f), f.calcCoulombEnergy(g),
                f.evalf(mrcpp::Coord<D>{1.0, 1.1, 0.9}));
    mrcpp::GaussExp<D> expansion;
    expansion.append(f);
    expansion.append(g);
    std::printf("exp_size %d\nexp_evalf %.17g\n", expansion.size(), expansion.evalf(mrcpp::Coord<D>{0.2, -0.1, 0.4}));

    mrcpp::PoissonOperator P(MRA, 1.0e-5);
    mrcpp::HelmholtzOperator H(MRA, 1.0, 1.0e-5);
    mrcpp::ABGVOperator<D> Dx(MRA, 0.5, 0.5);
    mrcpp::PHOperator<D> ph(MRA, 2);
    mrcpp::BSOperator<D> bs(MRA, 3);
    std::printf("ph_order %d\nbs_order %d\nph_terms %d\n", ph.getOrder(), bs.getOrder(), ph.size());
    std::printf("poisson_terms %d\nhelmholtz_terms %d\nhelmholtz_mu %.17g\nabgv_order %d\nbuild_prec %.17g\n", P.size(), H.size(), H.getMu(),
                Dx.getOrder(), P.getBuildPrec());

    mrcpp::FunctionTree<D> tree(MRA), grid(MRA);
    std::printf("root_nodes %d\n", tree.getNNodes());
    mrcpp::build_grid(tree, f);
    std::printf("grid_nodes %d\ngrid_end_nodes %d\n", tree.getNNodes(), tree.getNEndNodes());
    mrcpp::build_grid(grid, expansion);
    std::printf("grid2_nodes %d\n", grid.getNNodes());
    mrcpp::FunctionTree<D> copy(MRA);
    mrcpp::copy_grid(copy, grid);
    std::printf("copy_nodes %d\nsquare_norm_empty %.17g\n", copy.getNNodes(), copy.getSquareNorm());
    mrcpp::clear_grid(copy);
    std::printf("clear_grid_nodes %d\n", copy.getNNodes());
    mrcpp::IdentityConvolution<D> I(MRA, 1.0e-4);
    std::printf("identity_terms %d\n", I.size());
    copy.clear();
    std::printf("cleared_nodes %d\n", copy.getNNodes());

    // the sharded form of apply and its communicator wrapper are part of the mirror (instantiated here, run on multi-GPU boxes)
    using ShardedApply = void (*)(double, mrcpp::FunctionTree<D> &, mrcpp::ConvolutionOperator<D> &, mrcpp::FunctionTree<D> &,
                                  const mrcpp::b200::Comm &, int, bool);
    ShardedApply sharded = &mrcpp::apply<D, double>;
    std::printf("sharded_apply_present %d\n", sharded != nullptr ? 1 : 0);

    timer.stop();
    mrcpp::print::value(0, "elapsed", timer.elapsed(), "(sec)");
    mrcpp::print::tree(0, "grid", tree, timer);
    mrcpp::print::footer(0, timer);
    const std::string text = log.str();
    std::printf("log_lines %d\n", (int)std::count(text.begin(), text.end(), '\n'));
    std::fputs(text.c_str(), stderr);
    return 0;
}
