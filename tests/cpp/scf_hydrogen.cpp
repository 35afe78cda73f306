// Ground state of the hydrogen atom by the integral-equation iteration  phi <- -2 G_mu [V phi],  G_mu = Helmholtz operator with
// mu = sqrt(-2 E), energy updated from the residual -- one complete self-consistency loop on resident trees: callback
// projection, multiply, Helmholtz apply, rescale, add, dot, normalize, operator construction per iteration (the workload shape
// of the reference's examples/scf.cpp). Written against the MRCPP API names of include/MRCPP/.
// Prints "key value" lines: energies per iteration, final energy (exact: -0.5 Hartree), norm of the last update.
#include <memory>

#include "MRCPP/MWFunctions"
#include "MRCPP/MWOperators"
#include "MRCPP/Printer"
#include "MRCPP/Timer"

using namespace mrcpp;
constexpr int D = 3;

int main(int argc, char **argv) {
    const double prec = argc > 1 ? std::atof(argv[1]) : 1.0e-4;
    const int order = 6;
    Printer::init(-1);
    MultiResolutionAnalysis<D> MRA(BoundingBox<D>(-4, std::array<int, D>{-1, -1, -1}, std::array<int, D>{2, 2, 2}), InterpolatingBasis(order), 25);

    // nuclear potential -1/r, smoothed below r = c (the error function form, regular at the origin)
    const double c = 0.00435 * prec;
    auto potential = [c](const Coord<D> &r) -> double {
        const double x = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) / c;
        if (x < 1.0e-12) return -23.0 / (3.0 * std::sqrt(pi)) / c; // limit x -> 0 of u(x) below: 2 / sqrt(pi) + 17 / (3 sqrt(pi))
        const double u = std::erf(x) / x + (std::exp(-x * x) + 16.0 * std::exp(-4.0 * x * x)) / (3.0 * std::sqrt(pi));
        return -u / c;
    };
    FunctionTree<D> V(MRA);
    project<D, double>(prec, V, potential);

    // starting orbital: a normalised Gaussian, deliberately not the exact exponential
    auto guess = [](const Coord<D> &r) -> double { return std::exp(-0.7 * (r[0] * r[0] + r[1] * r[1] + r[2] * r[2])); };
    auto phi = std::make_unique<FunctionTree<D>>(MRA);
    project<D, double>(prec, *phi, guess);
    phi->normalize();

    double energy = -0.4, update = 1.0;
    int iter = 0;
    Timer total;
    while (update > 10.0 * prec && iter < 12) {
        iter++;
        if (energy > 0.0) energy = -energy;
        HelmholtzOperator G(MRA, std::sqrt(-2.0 * energy), prec); // a new operator every cycle

        FunctionTree<D> Vphi(MRA);
        copy_grid(Vphi, *phi);
        multiply(prec, Vphi, 1.0, V, *phi, 1); // grid of the orbital, relaxed by at most one level

        auto next = std::make_unique<FunctionTree<D>>(MRA);
        apply(prec, *next, G, Vphi);
        next->rescale(-1.0 / (2.0 * pi));

        FunctionTree<D> delta(MRA);
        copy_grid(delta, *next);
        add(-1.0, delta, 1.0, *next, -1.0, *phi); // difference on the grid of the new orbital
        update = std::sqrt(delta.getSquareNorm());
        const double dE = dot(Vphi, delta) / next->getSquareNorm();
        energy += dE;
        std::printf("energy_%d %.17g\nupdate_%d %.17g\nnodes_%d %d\n", iter, energy, iter, update, iter, next->getNNodes());
        next->normalize();
        phi = std::move(next);
    }
    total.stop();
    std::printf("iterations %d\nfinal_energy %.17g\nfinal_update %.17g\norbital_norm %.17g\norbital_integral %.17g\n", iter, energy, update,
                std::sqrt(phi->getSquareNorm()), phi->integrate());
    std::printf("done 1\n");
    return 0;
}
