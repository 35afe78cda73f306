"""GPU parity (-m gpu) at the configurations BASELINE.json states, AT THE STATED PRECISION (SURVEY.md §8(d) C1, C3, C4, C5 and the
bench generator), against the CPU oracle and — for the north-star target — against the real reference compiled in place
(oracle/_ref). Bars as everywhere: node set, slot order and screened-tuple count bit-exact, per-node coefficients within
1e-12 of the node norm (strict figure printed and recorded by tests/parity_util.py), energies against the analytic values
the reference's examples check (examples/poisson.cpp:7-63, examples/scf.cpp:101-111, examples/derivative.cpp).

Sizes are chosen so that the oracle (about 1 K output nodes per second on 16 host threads) finishes each case within a
minute; the full-size workloads are covered by the size-independent properties in tests/test_gpu_parity.py
(test_full_size_coulomb_energy) and by bench.py."""
import math

import numpy as np
import pytest

from test_gpu_parity import COEF_TOL, assert_same_tree, gaussians, world

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(libs):
    mw, orc = libs
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return mw, orc


def _poisson_case(mw, orc, k, prec, func, device_projection=True):
    mra = world(mw, k)
    P = mw.PoissonOperator(mra, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func, device=device_projection)
    orc.project(prec, fc, func)
    assert_same_tree(fg, fc)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, P, fg)
    sc = orc.apply(prec, gc, P, fc)
    assert sg.g_nodes == sc.gNodes and sg.iterations == sc.iters
    assert sg.f_applied == sc.fApplied, (sg.f_applied, sc.fApplied)
    # generated input nodes: the device creates only those its (conservative) early-out cannot rule out, the oracle every node of
    # the band like the reference; the nodes that carry surviving tuples are the same (tuple counts equal)
    assert 0 <= sg.gen_nodes <= sc.genUsed
    assert_same_tree(gg, gc)
    assert abs(gg.getSquareNorm() - gc.getSquareNorm()) <= 1e-12 * gc.getSquareNorm()
    return mra, P, fg, gg, fc, gc


def test_c1_target_k7_prec1e7_vs_oracle(gpu):
    """north-star target: examples/poisson.cpp at k = 7, prec 1e-7 (M = 105): adaptive apply, GPU vs oracle, and the
    reference's own acceptance check (Coulomb self-energy sqrt(2 beta / pi), examples/poisson.cpp:56-63)."""
    mw, orc = gpu
    beta = 100.0
    func = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    mra, P, fg, gg, fc, gc = _poisson_case(mw, orc, 7, 1e-7, func)
    assert P.size() == 105
    en = mw.dot(gg, fg)
    assert abs(en - math.sqrt(2.0 * beta / math.pi)) / math.sqrt(2.0 * beta / math.pi) < 1e-7
    assert abs(en - orc.dot(gc, fc)) <= 1e-12 * abs(en)


def test_c1_target_k7_prec1e7_vs_real_reference(gpu):
    """the same target case against the REAL reference (MRCPP's own sources compiled in place, oracle/_ref): node set
    identical, coefficients within 1e-12 of the node norm, energy to 1e-11."""
    import ref_api as ref
    from parity_util import coef_parity
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    mw, orc = gpu
    k, prec, beta = 7, 1e-7, 100.0
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    wargs = (k, -4, (-1, -1, -1), (2, 2, 2), 25)
    try:
        rm = ref.MRA(*wargs)
    except OSError as e:
        pytest.skip(f"oracle/_ref does not load here: {e}")
    rf, rg = ref.Tree(rm), ref.Tree(rm)
    ref.project(prec, rf, [f])
    RP = ref.poisson(rm, prec)
    ref.apply(prec, rg, RP, rf)
    mra = mw.MultiResolutionAnalysis(*wargs)
    P = mw.PoissonOperator(mra, prec)
    assert ref.lib().ref_oper_n_terms(RP) == P.size() == 105
    fg, gg = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, f, device=True)
    mw.apply(prec, gg, P, fg)
    for name, R, G in (("input", rf.export(), fg.to_arrays()), ("output", rg.export(), gg.to_arrays())):
        ri, gi = ref.by_index(R), ref.by_index(G)
        assert set(ri) == set(gi), name
        keys = list(ri)
        ia = np.array([ri[q] for q in keys])
        ja = np.array([gi[q] for q in keys])
        assert np.array_equal(R["branch"][ia] != 0, G["child0"][ja] >= 0)
        rep = coef_parity(G["coefs"][ja], R["coefs"][ia].reshape(len(keys), -1), tol=COEF_TOL, label=f"c1_k7_1e-7_vs_real_reference[{name}]")
        assert rep["floored"] < COEF_TOL, rep
    assert abs(mw.dot(gg, fg) - ref.dot(rg, rf)) < 1e-11 * abs(ref.dot(rg, rf))
    assert abs(gg.getSquareNorm() - rg.square_norm()) <= 1e-12 * rg.square_norm()


def test_bench_generator_k7_prec1e7_16_centres(gpu):
    """the bench workload's generator (seed 42, centres in [-8,8]^3, beta log-uniform in [10,1000]) at k = 7, prec 1e-7 on the
    16-centre sample the CPU arm of bench.py times: 4 224 output nodes, 8.0 M tuples, generated input nodes in play."""
    mw, orc = gpu
    func = gaussians(16, 42)
    mra, P, fg, gg, fc, gc = _poisson_case(mw, orc, 7, 1e-7, func)
    # pairwise analytic Coulomb energy (GaussFunc::calcCoulombEnergy, src/functions/GaussFunc.cpp:210-237)
    exact = sum(a.calc_coulomb_energy(b) for a in func for b in func)
    assert abs(mw.dot(gg, fg) - exact) < 1e-6 * abs(exact)


@pytest.mark.parametrize("a,b", [(0.5, 0.5), (0.0, 0.0)])
def test_c3_abgv_10_gaussians_prec1e7(gpu, a, b):
    """C3: k = 7, 10 normalised Gaussians (seed 1234, centres in [-4,4]^3, beta log-uniform in [1,100]) projected at prec 1e-7,
    ABGVOperator(a, b) in all three directions; GPU vs oracle, and the derivative against the projected analytic derivative
    (examples/derivative.cpp pattern) for the central-difference-free parameter set."""
    mw, orc = gpu
    mra = world(mw, 7)
    func = gaussians(10, 1234, box=4.0, lo=0.0, hi=2.0)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(1e-7, fg, func, device=True)
    orc.project(1e-7, fc, func)
    assert_same_tree(fg, fc)
    D = mw.ABGVOperator(mra, a, b)
    for d in range(3):
        og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        sg = mw.apply(None, og, D, fg, dir=d)
        sc = orc.apply_derivative(oc, D, fc, d)
        assert sg.f_applied == sc.fApplied
        assert_same_tree(og, oc)
        # analytic derivative of a Gaussian expansion: d/dx_d exp(-b r^2) = -2 b (x_d - p_d) exp(-b r^2)
        dfunc = mw.GaussExp()
        for g in func:
            pw = [0, 0, 0]
            pw[d] = 1
            dfunc.append(mw.GaussFunc(g.beta, -2.0 * g.beta * g.coef, g.pos, tuple(pw)))
        ref_t = mw.FunctionTree(mra)
        mw.project(1e-7, ref_t, dfunc, device=False)
        diff = mw.FunctionTree(mra)
        mw.build_grid(diff, og)
        mw.build_grid(diff, ref_t)
        mw.add(-1.0, diff, [(1.0, og), (-1.0, ref_t)])
        rel = math.sqrt(diff.getSquareNorm() / ref_t.getSquareNorm())
        assert rel < 1e-5, rel


def benzene_orbital(mw, j):
    """C4 input j: 12 benzene-like centres (6 at radius 2.64, 6 at 4.69 bohr, z = 0, 60 degrees apart), exponents 1.5 / 0.8,
    coefficients N(0,1) with seed 2024 + j"""
    rng = np.random.default_rng(2024 + j)
    ge = mw.GaussExp()
    for a in range(12):
        r = 2.64 if a < 6 else 4.69
        ang = math.pi / 3.0 * (a % 6)
        beta = 1.5 if a < 6 else 0.8
        ge.append(mw.GaussFunc(beta, float(rng.normal()), (r * math.cos(ang), r * math.sin(ang), 0.0)))
    return ge


def test_c4_helmholtz_k9_prec1e7_one_orbital(gpu):
    """C4: HelmholtzOperator(mu = 1, prec 1e-7) at k = 9 (M = 85) on one synthetic orbital tree of the scf.cpp pattern
    (examples/scf.cpp:101-111): GPU vs oracle."""
    mw, orc = gpu
    k, prec = 9, 1e-7
    mra = world(mw, k)
    func = benzene_orbital(mw, 0)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func, device=True)
    orc.project(prec, fc, func)
    assert_same_tree(fg, fc)
    H = mw.HelmholtzOperator(mra, 1.0, prec)
    assert H.size() == 85
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, H, fg)
    sc = orc.apply(prec, gc, H, fc)
    assert sg.g_nodes == sc.gNodes and sg.f_applied == sc.fApplied and sg.iterations == sc.iters
    assert_same_tree(gg, gc)


def test_c5_poisson_k11_prec1e9_10_centres(gpu):
    """C5: k = 11, prec 1e-9 (M = 143) on 10 centres of the C5 generator (seed 42): GPU vs oracle, and the pairwise analytic
    Coulomb energy to the requested precision."""
    mw, orc = gpu
    func = gaussians(10, 42)
    mra, P, fg, gg, fc, gc = _poisson_case(mw, orc, 11, 1e-9, func)
    assert P.size() == 143
    exact = sum(a.calc_coulomb_energy(b) for a in func for b in func)
    assert abs(mw.dot(gg, fg) - exact) < 1e-8 * abs(exact)
