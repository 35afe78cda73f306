"""GPU parity (-m gpu) of apply(prec, out, oper, inp, precTrees, maxIter, absPrec) (src/treebuilders/apply.cpp:214-251): the device
path (per-node precision factor computed on the device, apply_prec.cu) against the oracle's restatement, which is pinned
against the REAL reference on the same three configurations (tests/test_reference_parity.py::
test_apply_with_prec_trees_matches_reference), and directly against the real reference where oracle/_ref is available."""
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


def _gauss(mw, n, seed, box, lo=1.0, hi=2.0):
    rng = np.random.default_rng(seed)
    g = mw.GaussExp()
    for _ in range(n):
        beta = 10.0 ** rng.uniform(lo, hi)
        g.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-box, box, 3))))
    return g


def _inputs(mw, orc, k, prec, box=(2.0, 2.0)):
    """the two trees of tests/test_reference_parity.py::_two_trees, on the device and on the host"""
    mra = world(mw, k)
    out = []
    for funcs in (_gauss(mw, 2, 71, box[0]), _gauss(mw, 3, 72, box[1])):
        a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
        mw.project(prec, a, funcs, device=True)
        orc.project(prec, b, funcs)
        assert_same_tree(a, b)
        out.append((a, b, funcs))
    return mra, out


@pytest.mark.parametrize("which", ["input", "other", "both", "none"])
def test_apply_with_prec_trees(gpu, which):
    mw, orc = gpu
    k, prec = 5, 1e-4
    mra, ((ga, ca, fa), (gb, cb, fb)) = _inputs(mw, orc, k, prec)
    P = mw.PoissonOperator(mra, prec)
    gpt = {"input": [ga], "other": [gb], "both": [ga, gb], "none": []}[which]
    cpt = {"input": [ca], "other": [cb], "both": [ca, cb], "none": []}[which]
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, og, P, ga, precTrees=gpt)
    sc = orc.apply_prec_trees(prec, oc, P, ca, cpt)
    assert sg.g_nodes == sc.gNodes and sg.iterations == sc.iters
    assert sg.f_applied == sc.fApplied, (sg.f_applied, sc.fApplied)
    assert_same_tree(og, oc)
    assert abs(og.getSquareNorm() - oc.getSquareNorm()) <= 1e-12 * oc.getSquareNorm()
    plain = mw.FunctionTree(mra)
    sp = mw.apply(prec, plain, P, ga)
    if which == "none":  # empty vector: factor 1, the plain apply
        assert plain.getNNodes() == og.getNNodes() and sp.f_applied == sg.f_applied
        assert np.array_equal(plain.to_arrays()["coefs"], og.to_arrays()["coefs"])
    else:                # the scaled precision changes the grid
        assert plain.getNNodes() != og.getNNodes()
    # the precision trees and the input are untouched (the reference deletes its generated nodes again)
    assert_same_tree(ga, ca)
    assert_same_tree(gb, cb)


def test_apply_with_prec_trees_k7_and_max_iter(gpu):
    """k = 7 (the DMMA contraction kernel), precision trees coarser and finer than the output grid, bounded iterations, absPrec"""
    mw, orc = gpu
    k, prec = 7, 1e-5
    mra, ((ga, ca, fa), (gb, cb, fb)) = _inputs(mw, orc, k, prec, box=(3.0, 1.0))
    P = mw.PoissonOperator(mra, prec)
    for max_iter, abs_prec in ((-1, False), (2, False), (-1, True)):
        og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        sg = mw.apply(prec, og, P, ga, maxIter=max_iter, absPrec=abs_prec, precTrees=[gb])
        sc = orc.apply_prec_trees(prec, oc, P, ca, [cb], maxIter=max_iter, absPrec=abs_prec)
        assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes
        assert_same_tree(og, oc)


def test_apply_with_prec_trees_vs_real_reference(gpu):
    import ref_api as ref
    from parity_util import coef_parity
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    mw, orc = gpu
    k, prec = 5, 1e-4
    wargs = (k, -4, (-1, -1, -1), (2, 2, 2), 25)
    try:
        rm = ref.MRA(*wargs)
    except OSError as e:
        pytest.skip(f"oracle/_ref does not load here: {e}")
    mra = mw.MultiResolutionAnalysis(*wargs)
    fa, fb = _gauss(mw, 2, 71, 2.0), _gauss(mw, 3, 72, 2.0)
    ra, rb = ref.Tree(rm), ref.Tree(rm)
    ref.project(prec, ra, list(fa))
    ref.project(prec, rb, list(fb))
    ga, gb = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, ga, fa, device=True)
    mw.project(prec, gb, fb, device=True)
    rg, og = ref.Tree(rm), mw.FunctionTree(mra)
    ref.apply_prec_trees(prec, rg, ref.poisson(rm, prec), ra, [ra, rb])
    mw.apply(prec, og, mw.PoissonOperator(mra, prec), ga, precTrees=[ga, gb])
    R, G = rg.export(), og.to_arrays()
    ri, gi = ref.by_index(R), ref.by_index(G)
    assert set(ri) == set(gi)
    keys = list(ri)
    ia = np.array([ri[q] for q in keys])
    ja = np.array([gi[q] for q in keys])
    rep = coef_parity(G["coefs"][ja], R["coefs"][ia].reshape(len(keys), -1), tol=COEF_TOL, label="prec_trees_vs_real_reference")
    assert rep["floored"] < COEF_TOL, rep
    assert abs(og.getSquareNorm() - rg.square_norm()) <= 1e-12 * rg.square_norm()
