// Construction of separated-representation operators (one-off host cost) and the
// band-width bookkeeping `apply` needs. The result is a flat [term][depth][translation] table of
// (k+1)x(k+1) non-standard-form blocks that is packed once into HBM.
#include "mrx_host.hpp"

#include <algorithm>
#include <cstring>

namespace mrx {

double calc_min_distance(const MRA<3> &mra, double eps) { return std::sqrt(eps * std::pow(2.0, -mra.maxScale())); }

double calc_max_distance(const MRA<3> &mra) {
    // math_utils::calc_distance(lower, upper)
    double s = 0.0;
    for (int d = 0; d < 3; d++) {
        double x = mra.lower(d) - mra.upper(d);
        s += x * x;
    }
    return std::sqrt(s);
}

// PoissonKernel::PoissonKernel (src/operators/PoissonKernel.cpp:48-86). The reference mixes double
// variables with long-double literals; the same promotions are kept here.
GaussExp<1> poisson_kernel(double epsilon, double r_min, double r_max) {
    GaussExp<1> out;
    double r0 = r_min / r_max;
    double r1 = r_max;

    double t1 = 1.0L;
    while ((2.0 * t1 * std::exp(-t1)) > epsilon) t1 *= 1.1L;
    double t2 = 1.0L;
    while ((std::sqrt(t2) * std::exp(-t2) / r0) > epsilon) t2 *= 1.1L;

    double s1 = -std::log(2.0 * t1);
    double s2 = std::log(t2 / (r0 * r0)) / 2.0;

    double h = 1.0 / (0.2L - 0.47L * std::log10(epsilon));
    int n_exp = static_cast<int>(std::ceil((s2 - s1) / h) + 1);
    if (n_exp > MaxSepRank) MRX_ABORT("Maximum separation rank exceeded.");

    for (int i = 0; i < n_exp; i++) {
        double arg = s1 + h * i;
        double sinharg = std::sinh(arg);
        double cosharg = std::cosh(arg);
        double onepexp = 1.0 + std::exp(-sinharg);

        double expo = 4.0L * (sinharg + std::log(onepexp)) * (sinharg + std::log(onepexp));
        double coef = h * (4.0L / root_pi) * cosharg / onepexp;

        expo *= 1.0 / (r1 * r1);
        coef *= 1.0 / r1;
        if (i == 0 or i == (n_exp - 1)) coef *= 1.0 / 2.0;

        GaussFunc<1> g;
        g.alpha = expo;
        g.coef = coef;
        out.push_back(g);
    }
    return out;
}

// HelmholtzKernel::HelmholtzKernel (src/operators/HelmholtzKernel.cpp:47-80)
GaussExp<1> helmholtz_kernel(double mu, double epsilon, double r_min, double r_max) {
    GaussExp<1> out;
    double r0 = r_min / r_max;
    double r1 = r_max;
    double mu_tilde = mu * r1;

    double t = std::max((-2.5L * std::log(epsilon)), 5.0L);
    double s1 = -std::log(4 * t / (mu_tilde * mu_tilde)) / 2;
    double s2 = std::log(t / (r0 * r0)) / 2;

    double h = 1.0 / (0.20L - 0.47L * std::log10(epsilon));
    int n_exp = static_cast<int>(std::ceil((s2 - s1) / h) + 1);
    if (n_exp > MaxSepRank) MRX_ABORT("Maximum separation rank exceeded.");

    for (int i = 0; i < n_exp; i++) {
        double arg = s1 + h * i;
        double temp = -arg * 2.0;
        double temp2 = -mu_tilde * mu_tilde * std::exp(temp) / 4.0 + arg;
        double beta = (h * (2.0 / root_pi) * std::exp(temp2));
        double temp3 = 2.0L * arg;
        double alpha = std::exp(temp3);

        alpha *= 1.0 / (r1 * r1);
        beta *= 1.0 / r1;
        if (i == 0 or i == (n_exp - 1)) beta *= 1.0 / 2.0;

        GaussFunc<1> g;
        g.alpha = alpha;
        g.coef = beta;
        out.push_back(g);
    }
    return out;
}

namespace {

// OperatorTree::getMaxTranslations (OperatorTree.cpp:181-192)
std::vector<int> max_translations(const Tree<2> &t) {
    std::vector<int> mt(t.nDepths(), 0);
    for (const auto &nd : t.nodes) {
        int n = nd.scale - t.mra.rootScale;
        mt[n] = std::max(mt[n], std::abs(nd.l[0]));
        mt[n] = std::max(mt[n], std::abs(nd.l[1]));
    }
    return mt;
}

// OperatorTree::setupOperNodeCache (OperatorTree.cpp:200-238) flattened into an OperTerm
OperTerm flatten_oper_tree(Tree<2> &o_tree) {
    OperTerm term;
    const int K = o_tree.K;
    term.matStride = (size_t)4 * K * K;
    int nScales = o_tree.nDepths();
    std::vector<int> mt = max_translations(o_tree);
    term.nDepth = nScales;
    term.maxTransl = mt;
    term.offset.resize(nScales);
    size_t total = 0;
    for (int n = 0; n < nScales; n++) {
        term.offset[n] = total;
        total += 2 * (size_t)mt[n] + 1;
    }
    term.mats.assign(total * term.matStride, 0.0);
    term.norms.assign(total * 4, 0.0);
    for (int n = 0; n < nScales; n++) {
        int scale = o_tree.mra.rootScale + n;
        for (int t = -mt[n]; t <= mt[n]; t++) {
            std::array<int, 2> l = (t <= 0) ? std::array<int, 2>{0, -t} : std::array<int, 2>{t, 0};
            int node = o_tree.getNode(scale, l); // may generate (regular) operator nodes
            size_t slot = term.offset[n] + (size_t)(t + mt[n]);
            std::memcpy(term.mats.data() + slot * term.matStride, o_tree.coef(node), sizeof(double) * term.matStride);
            for (int c = 0; c < 4; c++) term.norms[slot * 4 + c] = o_tree.cnorm[(size_t)node * 4 + c];
        }
    }
    // generation cannot add depths or widen translations, but keep the invariant explicit
    if (o_tree.nDepths() != nScales) MRX_ABORT("operator cache generation changed tree depth");
    return term;
}

// CrossCorrelationCalculator::calcNode / applyCcc (CrossCorrelationCalculator.cpp:36-96)
void cross_correlation_node(Tree<2> &o_tree, int n, Tree<1> &kernel, const CrossCorr &cc) {
    o_tree.zeroCoefs(n);
    const int K = o_tree.K, Kd = o_tree.Kd; // Kd = K*K
    const int K2 = 2 * K;                   // kernel tree has order 2k+1 -> 2K scaling coefs
    int scale = o_tree.nodes[n].scale + 1;
    std::vector<double> vec_o((size_t)4 * Kd, 0.0);
    for (int i = 0; i < 4; i++) {
        int l0 = 2 * o_tree.nodes[n].l[0] + (i & 1);
        int l1 = 2 * o_tree.nodes[n].l[1] + ((i >> 1) & 1);
        int l_a = l1 - l0 - 1;
        int l_b = l1 - l0;
        int node_a = kernel.getNode(scale, {l_a});
        int node_b = kernel.getNode(scale, {l_b});
        const double *seg_a = kernel.coef(node_a);
        const double *seg_b = kernel.coef(node_b);
        for (int r = 0; r < Kd; r++) {
            double s = 0.0, s2 = 0.0;
            const double *Lr = cc.L.data() + (size_t)r * K2;
            const double *Rr = cc.R.data() + (size_t)r * K2;
            for (int j = 0; j < K2; j++) s += Lr[j] * seg_a[j];
            for (int j = 0; j < K2; j++) s2 += Rr[j] * seg_b[j];
            vec_o[(size_t)i * Kd + r] = s + s2;
        }
    }
    double *coefs = o_tree.coef(n);
    double two_n = std::pow(2.0, -scale / 2.0);
    for (int i = 0; i < 4 * Kd; i++) coefs[i] = std::sqrt(1.0) * two_n * vec_o[i];
    o_tree.mwTransformNode(n, Compression);
    o_tree.nodes[n].flags |= FlagHasCoefs;
    o_tree.calcNorms(n);
}

MRA<2> operator_mra(const MRA<3> &mra, int oper_root, int oper_reach) {
    // MWOperator::getOperatorMRA (MWOperator.cpp:110-129)
    int reach = oper_reach + 1;
    if (reach < 0)
        for (int i = 0; i < 3; i++)
            if (mra.nboxes[i] > reach) reach = mra.nboxes[i];
    MRA<2> o;
    o.order = mra.order;
    o.rootScale = oper_root;
    o.corner = {0, 0};
    o.nboxes = {reach, reach};
    o.maxDepth = MaxDepth;
    return o;
}

MRA<1> kernel_mra(const MRA<3> &mra, int oper_root, int oper_reach) {
    // ConvolutionOperator::getKernelMRA (ConvolutionOperator.cpp:110-142)
    int reach = oper_reach + 1;
    if (reach < 0)
        for (int i = 0; i < 3; i++)
            if (mra.nboxes[i] > reach) reach = mra.nboxes[i];
    MRA<1> km;
    km.order = 2 * mra.order + 1;
    km.rootScale = oper_root;
    km.corner = {-reach};
    km.nboxes = {2 * reach};
    km.maxDepth = MaxDepth;
    return km;
}

} // namespace

Operator build_convolution_operator(const MRA<3> &mra, const GaussExp<1> &kernel, double k_prec, double o_prec) {
    // ConvolutionOperator(mra) -> MWOperator(mra, mra.getRootScale(), -10); initialize(); initOperExp()
    return build_convolution_operator(mra, kernel, k_prec, o_prec, mra.rootScale, -10);
}

Operator build_convolution_operator(const MRA<3> &mra, const GaussExp<1> &kernel, double k_prec, double o_prec, int oper_root,
                                    int oper_reach) {
    Operator op;
    op.k = mra.order;
    op.K = mra.order + 1;
    op.operRoot = oper_root;
    op.operReach = oper_reach;
    op.buildPrec = o_prec;
    op.terms.resize(kernel.size());
    MRA<1> k_mra = kernel_mra(mra, oper_root, oper_reach);
    MRA<2> o_mra = operator_mra(mra, oper_root, oper_reach);
    const CrossCorr &cc = cross_corr(mra.order);
    filter_set(k_mra.order);
    filter_set(o_mra.order);
    quadrature(k_mra.order + 1);
    const int M = (int)kernel.size();
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < M; i++) {
        // Rescale Gaussian for D-dim application (ConvolutionOperator.cpp:86-88)
        GaussFunc<1> k_func = kernel[i];
        k_func.coef = std::copysign(std::pow(std::abs(k_func.coef), 1.0 / 3), k_func.coef);

        Tree<1> k_tree(k_mra);
        build_grid<1>(k_tree, k_func, -1);
        project<1>(
            k_prec, k_tree, [&k_func](const double *r) { return k_func.evalf(r); }, -1, false, true);

        Tree<2> o_tree(o_mra);
        o_tree.operNorms = true;
        o_tree.normPrec = o_prec;
        // TreeBuilder<2>::build with CrossCorrelationCalculator + OperatorAdaptor (OperatorAdaptor.h:38-48)
        build_tree<2>(
            o_tree, [&](Tree<2> &t, int n) { cross_correlation_node(t, n, k_tree, cc); },
            [](const Tree<2> &t, int n) {
                bool chkTransl = (t.nodes[n].l[0] == 0 or t.nodes[n].l[1] == 0);
                bool chkCompNorm = false;
                for (int c = 1; c < 4; c++)
                    if (t.cnorm[(size_t)n * 4 + c] > 0.0) chkCompNorm = true;
                return chkTransl and chkCompNorm;
            },
            -1, false, false);
        o_tree.mwTransformUpSerial();
        o_tree.calcSquareNorm();
        op.terms[i] = flatten_oper_tree(o_tree);
    }
    return op;
}

Operator build_poisson_operator(const MRA<3> &mra, double prec) {
    double o_prec = prec;
    double k_prec = prec / 10.0;
    double r_min = calc_min_distance(mra, k_prec);
    double r_max = calc_max_distance(mra);
    GaussExp<1> kernel = poisson_kernel(k_prec, r_min, r_max);
    return build_convolution_operator(mra, kernel, k_prec, o_prec);
}

Operator build_helmholtz_operator(const MRA<3> &mra, double mu, double prec) {
    double o_prec = prec;
    double k_prec = prec / 10.0;
    double r_min = calc_min_distance(mra, k_prec);
    double r_max = calc_max_distance(mra);
    GaussExp<1> kernel = helmholtz_kernel(mu, k_prec, r_min, r_max);
    return build_convolution_operator(mra, kernel, k_prec, o_prec);
}

Operator build_poisson_operator(const MRA<3> &mra, double prec, int oper_root, int oper_reach) {
    double o_prec = prec;
    double k_prec = prec / 100.0;
    double r_min = calc_min_distance(mra, k_prec);
    double r_max = calc_max_distance(mra);
    // Adjust r_max for periodic world (PoissonOperator.cpp:66-69)
    const int rel_root = oper_root - mra.rootScale;
    r_max *= std::pow(2.0, -rel_root);
    r_max *= (2.0 * oper_reach) + 1.0;
    GaussExp<1> kernel = poisson_kernel(k_prec, r_min, r_max);
    return build_convolution_operator(mra, kernel, k_prec, o_prec, oper_root, oper_reach);
}

Operator build_helmholtz_operator(const MRA<3> &mra, double mu, double prec, int oper_root, int oper_reach) {
    double o_prec = prec;
    double k_prec = prec / 100.0;
    double r_min = calc_min_distance(mra, k_prec);
    double r_max = calc_max_distance(mra);
    const int rel_root = oper_root - mra.rootScale;
    r_max *= std::pow(2.0, -rel_root);
    r_max *= (2.0 * oper_reach) + 1.0;
    GaussExp<1> kernel = helmholtz_kernel(mu, k_prec, r_min, r_max);
    return build_convolution_operator(mra, kernel, k_prec, o_prec, oper_root, oper_reach);
}

Operator build_abgv_operator(const MRA<3> &mra, double a, double b) {
    // ABGVOperator::initialize (ABGVOperator.cpp:52-74); DerivativeOperator(mra, root, reach=1)
    const int oper_root = mra.rootScale, oper_reach = 1;
    int bw = 0;
    if (std::abs(a) > MachineZero) bw = 1;
    if (std::abs(b) > MachineZero) bw = 1;
    MRA<2> o_mra = operator_mra(mra, oper_root, oper_reach);
    const int k = mra.order, K = k + 1;

    // ABGVCalculator ctor (ABGVCalculator.cpp:38-101), interpolating basis
    const Quadrature &q = quadrature(K);
    std::vector<double> Kmat((size_t)K * K), valueZero(K), valueOne(K);
    for (int i = 0; i < K; i++) {
        valueZero[i] = interp_scaling_eval(k, i, 0.0);
        valueOne[i] = interp_scaling_eval(k, i, 1.0);
        // ABGVCalculator::calcKMatrix: K(i,j) = 2 sqrt(w_j) * poly_i'(x_j) where Polynomial::calcDerivative
        // (Polynomial.cpp:229-237) differentiates w.r.t. the INTERNAL argument q = 2x-1, i.e. poly' = (1/2) dphi/dx
        for (int j = 0; j < K; j++)
            Kmat[i * K + j] = 2.0 * std::sqrt(q.weights[j]) * (0.5 * interp_scaling_deriv(k, i, q.roots[j]));
    }

    Tree<2> o_tree(o_mra);
    o_tree.operNorms = true;
    o_tree.normPrec = MachineZero;
    auto calc = [&](Tree<2> &t, int n) {
        // ABGVCalculator::calcNode (ABGVCalculator.cpp:104-168)
        t.zeroCoefs(n);
        int np1 = t.nodes[n].scale + 1;
        int kp1_d = K * K;
        double two_np1 = std::pow(2.0, np1);
        double *coefs = t.coef(n);
        switch (t.nodes[n].l[1] - t.nodes[n].l[0]) {
            case 1:
                if (b > MachineZero)
                    for (int i = 0; i < K; i++)
                        for (int j = 0; j < K; j++) coefs[1 * kp1_d + i * K + j] = -b * valueZero[i] * valueOne[j];
                break;
            case 0:
                for (int i = 0; i < K; i++) {
                    double zero_i = valueZero[i], one_i = valueOne[i];
                    for (int j = 0; j < K; j++) {
                        double K_ij = Kmat[i * K + j];
                        double one_j = valueOne[j], zero_j = valueZero[j];
                        double one_ij = one_i * one_j, zero_ij = zero_i * zero_j;
                        int idx = i * K + j;
                        coefs[0 * kp1_d + idx] = (1.0 - a) * one_ij - (1.0 - b) * zero_ij - K_ij;
                        coefs[1 * kp1_d + idx] = a * one_i * zero_j;
                        coefs[2 * kp1_d + idx] = -b * zero_i * one_j;
                        coefs[3 * kp1_d + idx] = (1.0 - a) * one_ij - (1.0 - b) * zero_ij - K_ij;
                    }
                }
                break;
            case -1:
                if (a > MachineZero)
                    for (int i = 0; i < K; i++)
                        for (int j = 0; j < K; j++) coefs[2 * kp1_d + i * K + j] = a * valueOne[i] * valueZero[j];
                break;
            default:
                break;
        }
        for (int i = 0; i < t.ncoef; i++) coefs[i] *= two_np1;
        t.mwTransformNode(n, Compression);
        t.nodes[n].flags |= FlagHasCoefs;
        t.calcNorms(n);
    };
    auto split = [bw](const Tree<2> &t, int n) {
        // BandWidthAdaptor::splitNode (BandWidthAdaptor.h:46-51)
        int dl = std::abs(t.nodes[n].l[0] - t.nodes[n].l[1]);
        return ((t.nodes[n].l[0] == 0) and (2 * dl <= bw));
    };
    build_tree<2>(o_tree, calc, split, -1, false, false);
    o_tree.calcSquareNorm();

    Operator op;
    op.k = k;
    op.K = K;
    op.operRoot = oper_root;
    op.derivative = true;
    op.order = 1;
    op.terms.push_back(flatten_oper_tree(o_tree));
    return op;
}

// PHOperator / BSOperator: the operator nodes are three tabulated K x K matrices scaled by 2^(order (n + 1)), laid out by
// PHCalculator::calcNode (PHCalculator.cpp:82-128) / BSCalculator::calcNode (BSCalculator.cpp:82-128, identical), compressed
// in the node; tree built by the BandWidthAdaptor with bandwidth 1.
static Operator build_tabulated_derivative(const MRA<3> &mra, int kind, int order, int oper_reach) {
    const int oper_root = mra.rootScale;
    const int bw = 1;
    MRA<2> o_mra = operator_mra(mra, oper_root, oper_reach);
    const int k = mra.order, K = k + 1;
    const std::vector<double> &tab = derivative_table(kind, k);
    const double *S_p1 = tab.data(), *S_0 = tab.data() + (size_t)K * K, *S_m1 = tab.data() + (size_t)2 * K * K;

    Tree<2> o_tree(o_mra);
    o_tree.operNorms = true;
    o_tree.normPrec = MachineZero;
    auto calc = [&](Tree<2> &t, int n) {
        t.zeroCoefs(n);
        const int np1 = t.nodes[n].scale + 1;
        const int kp1_d = K * K;
        const double two_np1 = std::pow(2.0, order * np1);
        double *coefs = t.coef(n);
        switch (t.nodes[n].l[1] - t.nodes[n].l[0]) {
            case 1:
                for (int idx = 0; idx < kp1_d; idx++) coefs[1 * kp1_d + idx] = two_np1 * S_p1[idx];
                break;
            case 0:
                for (int idx = 0; idx < kp1_d; idx++) {
                    coefs[0 * kp1_d + idx] = two_np1 * S_0[idx];
                    coefs[1 * kp1_d + idx] = two_np1 * S_m1[idx];
                    coefs[2 * kp1_d + idx] = two_np1 * S_p1[idx];
                    coefs[3 * kp1_d + idx] = two_np1 * S_0[idx];
                }
                break;
            case -1:
                for (int idx = 0; idx < kp1_d; idx++) coefs[2 * kp1_d + idx] = two_np1 * S_m1[idx];
                break;
            default:
                break;
        }
        t.mwTransformNode(n, Compression);
        t.nodes[n].flags |= FlagHasCoefs;
        t.calcNorms(n);
    };
    auto split = [bw](const Tree<2> &t, int n) {
        int dl = std::abs(t.nodes[n].l[0] - t.nodes[n].l[1]);
        return ((t.nodes[n].l[0] == 0) and (2 * dl <= bw)); // BandWidthAdaptor::splitNode (BandWidthAdaptor.h:46-51)
    };
    build_tree<2>(o_tree, calc, split, -1, false, false);
    o_tree.calcSquareNorm();

    Operator op;
    op.k = k;
    op.K = K;
    op.operRoot = oper_root;
    op.derivative = true;
    op.order = order;
    op.terms.push_back(flatten_oper_tree(o_tree));
    return op;
}

Operator build_ph_operator(const MRA<3> &mra, int order) {
    if (order < 1 || order > 2) MRX_ABORT("PHOperator: derivative order 1 or 2"); // PHCalculator.cpp:40-45
    return build_tabulated_derivative(mra, 3 + order, order, /*oper_reach=*/-10);  // PHOperator.cpp:41
}

Operator build_bs_operator(const MRA<3> &mra, int order) {
    if (order < 1 || order > 3) MRX_ABORT("BSOperator: derivative order 1, 2 or 3"); // BSCalculator.cpp:38-45
    return build_tabulated_derivative(mra, 5 + order, order, /*oper_reach=*/1);    // BSOperator.cpp:41
}

// OperatorTree::calcBandWidth (OperatorTree.cpp:109-134) per term, then MWOperator::calcBandWidths
// (MWOperator.cpp:78-108). prec < 0 means "use the operator's build precision".
void Operator::calcBandWidths(double prec) {
    int maxDepth = 0;
    for (auto &t : terms) {
        t.bw.assign(t.nDepth + 1, {-1, -1, -1, -1, -1});
        double p = (prec < 0.0) ? buildPrec : prec;
        for (int depth = 0; depth < t.nDepth; depth++) {
            int l = 0;
            bool done = false;
            while (not done) {
                done = true;
                const double *nrm = t.nodeNorms(depth, l);
                double thrs = std::max(MachinePrec, p / (8.0 * (1 << depth)));
                for (int c = 0; c < 4; c++) {
                    if (nrm[c] > thrs) {
                        t.bw[depth][c] = l;
                        if (l > t.bw[depth][4]) t.bw[depth][4] = l;
                        done = false;
                    }
                }
                if (++l > t.maxTransl[depth]) break;
            }
        }
        if (t.bwDepth() > maxDepth) maxDepth = t.bwDepth();
    }
    bandMax.assign(maxDepth + 1, -1);
    for (auto &t : terms)
        for (int n = 0; n <= t.bwDepth(); n++)
            for (int j = 0; j < 4; j++) bandMax[n] = std::max(bandMax[n], t.bw[n][j]);
}

void Operator::clearBandWidths() {
    for (auto &t : terms) t.bw.clear();
    bandMax.clear();
}

int Operator::getMaxBandWidth(int depth) const {
    int maxWidth = -1;
    if (depth < 0) {
        if (!bandMax.empty()) maxWidth = *std::max_element(bandMax.begin(), bandMax.end());
    } else if (depth < (int)bandMax.size()) {
        maxWidth = bandMax[depth];
    }
    return maxWidth;
}

} // namespace mrx
