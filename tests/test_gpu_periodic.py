"""GPU parity (-m gpu) of the PERIODIC operator application: mrcpp::apply on a periodic world with operators built for a reach
(src/treebuilders/apply.cpp:68-93; ConvolutionCalculator::makeOperBand periodic branch src/treebuilders/ConvolutionCalculator.cpp:166-172;
MWTree::getNode index wrap src/trees/MWTree.cpp:341, src/utils/periodic_utils.cpp:49-73) and apply_near_field / apply_far_field
(apply.cpp:294-342, fillOperBand :191-218). Device path against the oracle's restatement, which tests/test_reference_parity.py::
test_periodic_apply_matches_reference pins against the real reference on the same inputs; plus the reference's own acceptance
check (tests/operators/poisson_operator.cpp:156-199: the periodic Poisson solution of a cosine source) and near + far = whole."""
import math

import numpy as np
import pytest

from test_gpu_parity import assert_same_tree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(libs):
    mw, orc = libs
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return mw, orc


AMP = [3.0 * math.pi ** 2 / (4.0 * math.pi), 0.7]
KV = [[1, 1, 1], [2, 1, 3]]


def _inputs(mw, orc, k, proj_prec):
    mra = mw.MultiResolutionAnalysis(k, 0, (-1, -1, -1), (2, 2, 2), 25, periodic=True)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project_cosines(proj_prec, fg, AMP, KV)                      # host quadrature, device BottomUp
    mw.project_cosines(proj_prec, fc, AMP, KV, finalize=False)      # ... closed by the oracle
    orc.mw_transform_up(fc)
    orc.calc_square_norm(fc)
    assert_same_tree(fg, fc)
    return mra, fg, fc


@pytest.mark.parametrize("k,proj_prec,apply_prec,build_prec", [(5, 1e-4, 1e-3, 1e-3), (7, 1e-5, 1e-4, 1e-4)])
@pytest.mark.parametrize("kind", ["poisson", "helmholtz"])
def test_periodic_apply_and_near_far_field(gpu, kind, k, proj_prec, apply_prec, build_prec):
    mw, orc = gpu
    reach = 9
    mra, fg, fc = _inputs(mw, orc, k, proj_prec)
    P = mw.PoissonOperator(mra, build_prec, 0, reach) if kind == "poisson" else mw.HelmholtzOperator(mra, 4.3, build_prec, 0, reach)
    trees = {}
    for mode in ("plain", "near", "far"):
        gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        if mode == "plain":
            sg = mw.apply(apply_prec, gg, P, fg)
            sc = orc.apply(apply_prec, gc, P, fc)
        else:
            sg = (mw.apply_near_field if mode == "near" else mw.apply_far_field)(apply_prec, gg, P, fg)
            sc = orc.apply_unit_cell(mode == "near", apply_prec, gc, P, fc)
        assert sg.g_nodes == sc.gNodes and sg.iterations == sc.iters
        assert sg.f_applied == sc.fApplied, (mode, sg.f_applied, sc.fApplied)
        assert_same_tree(gg, gc)
        assert abs(gg.getSquareNorm() - gc.getSquareNorm()) <= 1e-12 * gc.getSquareNorm()
        trees[mode] = gg
    # near field + far field = the whole periodic apply (the band is partitioned by in_unit_cell): compare function values
    pts = np.random.default_rng(3).uniform(-1, 1, (40, 3))
    whole = trees["plain"].evalf(pts)
    parts = trees["near"].evalf(pts) + trees["far"].evalf(pts)
    assert np.abs(whole - parts).max() <= 20 * apply_prec * np.abs(whole).max()
    if kind == "poisson":
        # -lap u = 4 pi rho on the periodic cell: u = cos cos cos + 0.7 * 4 pi / (14 pi^2) cos(2 pi x) cos(pi y) cos(3 pi z)
        def exact(r):
            r = np.asarray(r)
            return (np.cos(math.pi * r[..., 0]) * np.cos(math.pi * r[..., 1]) * np.cos(math.pi * r[..., 2]) +
                    0.7 * 4.0 * math.pi / (14.0 * math.pi ** 2) * np.cos(2 * math.pi * r[..., 0]) * np.cos(math.pi * r[..., 1]) *
                    np.cos(3 * math.pi * r[..., 2]))
        # the constant (k = 0) Fourier mode of the periodised kernel is fixed by the finite reach: compare up to that offset
        off = float(np.mean(whole - exact(pts)))
        assert np.abs(whole - off - exact(pts)).max() < 3e-2, np.abs(whole - off - exact(pts)).max()


def test_periodic_apply_vs_real_reference(gpu):
    import ref_api as ref
    from parity_util import coef_parity
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    mw, orc = gpu
    k, proj_prec, apply_prec, build_prec, reach = 5, 1e-4, 1e-3, 1e-3, 9
    try:
        rm = ref.PeriodicMRA(k)
    except OSError as e:
        pytest.skip(f"oracle/_ref does not load here: {e}")
    rf = ref.Tree(rm)
    ref.project_cosines(proj_prec, rf, AMP, KV)
    RP = ref.poisson_reach(rm, build_prec, 0, reach)
    mra, fg, fc = _inputs(mw, orc, k, proj_prec)
    P = mw.PoissonOperator(mra, build_prec, 0, reach)
    assert ref.lib().ref_oper_n_terms(RP) == P.size()
    for mode in ("plain", "near", "far"):
        rg, gg = ref.Tree(rm), mw.FunctionTree(mra)
        if mode == "plain":
            ref.apply(apply_prec, rg, RP, rf)
            mw.apply(apply_prec, gg, P, fg)
        else:
            ref.apply_unit_cell(mode == "near", apply_prec, rg, RP, rf)
            (mw.apply_near_field if mode == "near" else mw.apply_far_field)(apply_prec, gg, P, fg)
        R, G = rg.export(), gg.to_arrays()
        ri, gi = ref.by_index(R), ref.by_index(G)
        assert set(ri) == set(gi), mode
        keys = list(ri)
        ia = np.array([ri[q] for q in keys])
        ja = np.array([gi[q] for q in keys])
        rep = coef_parity(G["coefs"][ja], R["coefs"][ia].reshape(len(keys), -1), label=f"periodic_{mode}_vs_real_reference")
        assert rep["floored"] < 1e-12, rep


def test_periodic_apply_needs_a_reach_operator(gpu):
    """a plain PoissonOperator(mra, prec) has no reach: the periodic apply refuses it (print + abort like MSG_ABORT), checked in a
    child process"""
    import subprocess
    import sys
    code = ("import mrcpp_b200 as mw\n"
            "m = mw.MultiResolutionAnalysis(5, 0, (-1,-1,-1), (2,2,2), 25, periodic=True)\n"
            "f = mw.FunctionTree(m); mw.project_cosines(1e-3, f, [1.0], [[1,1,1]])\n"
            "g = mw.FunctionTree(m); mw.apply(1e-3, g, mw.PoissonOperator(m, 1e-3), f)\n")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root, timeout=300)
    assert r.returncode != 0 and "reach" in r.stderr
