"""GPU parity tests (-m gpu) of the callers next to the derivative apply (src/treebuilders/apply.h:51-55,
add.h, grid.h:37): add on a given grid, build_grid from trees, gradient, divergence, integrate -- the CUDA path through the
C ABI against the CPU oracle (which tests/test_reference_parity.py pins against the real reference for the same calls)."""
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


def two_trees(mw, orc, k, prec):
    mra = world(mw, k)
    out = []
    for n, seed, box in ((2, 71, 4.0), (3, 72, 2.0)):
        func = gaussians(n, seed, box=box, lo=1.0, hi=2.0)
        g, c = mw.FunctionTree(mra), mw.FunctionTree(mra)
        mw.project(prec, g, func)
        orc.project(prec, c, func)
        out.append((g, c))
    return mra, out


@pytest.mark.parametrize("family,order", [("ph", 1), ("ph", 2), ("bs", 1), ("bs", 2), ("bs", 3)])
def test_ph_bs_derivative(gpu, family, order):
    """PHOperator / BSOperator (PHOperator.cpp:40-69, BSOperator.cpp:40-66) through apply(out, D, inp, dir): the device path
    against the oracle (pinned against the real reference for the same operators in tests/test_reference_parity.py)"""
    mw, orc = gpu
    mra, ((ga, ca), _) = two_trees(mw, orc, 5, 1e-4)
    D = mw.PHOperator(mra, order) if family == "ph" else mw.BSOperator(mra, order)
    for d in range(3):
        og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        sg = mw.apply(None, og, D, ga, dir=d)
        sc = orc.apply_derivative(oc, D, ca, d)
        assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes
        assert_same_tree(og, oc, tol=1e-11)


@pytest.mark.parametrize("k,grid", [(5, "union"), (7, "union"), (5, "roots"), (5, "first"), (4, "union")])
def test_add_on_grid(gpu, k, grid):
    """add(-1.0, out, {(a, f), (b, g)}, 0) (add.cpp:41-70): union grid, bare roots (inputs truncated), grid of the first
    input (second input partly coarser -> generated nodes, partly finer -> truncated)"""
    mw, orc = gpu
    mra, ((ga, ca), (gb, cb)) = two_trees(mw, orc, k, 1e-4)
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    for o, a, b in ((og, ga, gb), (oc, ca, cb)):
        if grid in ("union", "first"):
            mw.build_grid(o, a)
        if grid == "union":
            mw.build_grid(o, b)
    mw.add(-1.0, og, [(0.5, ga), (-2.0, gb)], 0)
    orc.add(oc, [0.5, -2.0], [ca, cb])
    assert_same_tree(og, oc)
    assert abs(og.getSquareNorm() - oc.getSquareNorm()) <= 1e-12 * oc.getSquareNorm()
    assert abs(og.integrate() - oc.integrate()) <= 1e-12 * max(abs(oc.integrate()), 1.0)
    # linear functional: <sum | f> = a <f | f> + b <g | f> on the union grid (nothing truncated)
    if grid == "union":
        want = 0.5 * mw.dot(ga, ga) - 2.0 * mw.dot(gb, ga)
        assert abs(mw.dot(og, ga) - want) <= 1e-10 * abs(want)


def test_refine_grid_and_inplace_add(gpu):
    """refine_grid(out, prec) / refine_grid(out, scales) (grid.cpp:271-302) and FunctionTree::add(c, inp) in place
    (FunctionTree.cpp:687-706) on device-resident trees vs the oracle"""
    mw, orc = gpu
    mra, ((ga, ca), (gb, cb)) = two_trees(mw, orc, 5, 1e-3)
    n_g, n_c = mw.refine_grid(ga, 1e-5), orc.refine_grid(ca, prec=1e-5)
    assert n_g == n_c and n_g > 0
    assert_same_tree(ga, ca)
    n_g, n_c = mw.refine_grid(gb, scales=1), orc.refine_grid(cb, scales=1)
    assert n_g == n_c and n_g > 0
    assert_same_tree(gb, cb)
    assert abs(mw.dot(gb, gb) - orc.dot(cb, cb)) <= 1e-12 * orc.dot(cb, cb)
    ga.add(-0.5, gb)
    orc.add_inplace(ca, -0.5, cb)
    assert_same_tree(ga, ca)
    assert abs(ga.getSquareNorm() - ca.getSquareNorm()) <= 1e-12 * ca.getSquareNorm()
    grid = mw.FunctionTree(mra)            # a grid without coefficients is refined on the host
    assert mw.refine_grid(grid, scales=2) == 64 + 512 and grid.getNNodes() == 8 + 64 + 512


def test_gradient_and_divergence(gpu):
    """gradient(D, f) and divergence(out, D, {f, g, f}) (apply.cpp:444-452, :514-530) vs the oracle; integrate() of device
    resident trees (root blocks read back from HBM) vs the oracle's host trees"""
    mw, orc = gpu
    mra, ((ga, ca), (gb, cb)) = two_trees(mw, orc, 5, 1e-4)
    D = mw.ABGVOperator(mra, 0.5, 0.5)
    grad = mw.gradient(D, ga)
    for d, (c, g) in enumerate(grad):
        ref = mw.FunctionTree(mra)
        orc.apply_derivative(ref, D, ca, d)
        assert c == 1.0
        assert_same_tree(g, ref)
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.divergence(og, D, [(1.0, ga), (1.0, gb), (1.0, ga)])
    orc.divergence(oc, D, [ca, cb, ca])
    assert_same_tree(og, oc, tol=1e-11)
    assert abs(ga.integrate() - ca.integrate()) <= 1e-13 and abs(og.integrate() - oc.integrate()) <= 1e-10


@pytest.mark.parametrize("k,prec,max_iter,abs_prec,start", [(5, 1e-4, -1, False, "roots"), (7, 1e-5, -1, False, "roots"), (5, 1e-3, -1, True, "roots"),
                                                             (5, 1e-4, 2, False, "roots"), (5, 1e-5, -1, False, "first"), (4, 1e-4, -1, False, "roots")])
def test_adaptive_add(gpu, k, prec, max_iter, abs_prec, start):
    """add(prec, out, {(a, f), (b, g)}, maxIter, absPrec) (add.cpp:41-70), the adaptive form: node set, coefficients, norm"""
    mw, orc = gpu
    mra, ((ga, ca), (gb, cb)) = two_trees(mw, orc, k, 1e-5)
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    if start == "first":
        mw.build_grid(og, ga)
        mw.build_grid(oc, ca)
    mw.add(prec, og, [(1.0, ga), (-2.0, gb)], max_iter, abs_prec)
    orc.add(oc, [1.0, -2.0], [ca, cb], prec=prec, maxIter=max_iter, absPrec=abs_prec)
    assert og.getNNodes() > 8
    assert_same_tree(og, oc)
    assert abs(og.getSquareNorm() - oc.getSquareNorm()) <= 1e-12 * oc.getSquareNorm()
    # the sum is usable as an apply input like any other tree
    P = mw.PoissonOperator(mra, 1e-3)
    vg, vc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(1e-3, vg, P, og)
    sc = orc.apply(1e-3, vc, P, oc)
    assert sg.f_applied == sc.fApplied
    assert_same_tree(vg, vc)


@pytest.mark.parametrize("k,prec,max_iter,abs_prec,start", [(5, 1e-4, -1, False, "roots"), (7, 1e-5, -1, False, "roots"), (5, 1e-3, -1, True, "roots"),
                                                             (5, 1e-4, 2, False, "roots"), (5, -1.0, -1, False, "union"), (4, 1e-4, -1, False, "roots")])
def test_multiply(gpu, k, prec, max_iter, abs_prec, start):
    """multiply(prec, out, {(c, f), (1, g)}, maxIter, absPrec) (multiply.cpp:104-136): node set identical; coefficients against
    the largest node norm (a product carries the rounding of its larger factor into regions where it is tiny itself, see
    tests/test_reference_parity.py::test_multiply_matches_reference)"""
    mw, orc = gpu
    mra = world(mw, k)
    trees = []
    for n, seed in ((2, 71), (3, 72)):
        func = gaussians(n, seed, box=1.0, lo=1.0, hi=2.0)   # overlapping functions: a product of O(1)
        g, c = mw.FunctionTree(mra), mw.FunctionTree(mra)
        mw.project(1e-5, g, func)
        orc.project(1e-5, c, func)
        trees.append((g, c))
    (ga, ca), (gb, cb) = trees
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    if start == "union":
        for o, a, b in ((og, ga, gb), (oc, ca, cb)):
            mw.build_grid(o, a)
            mw.build_grid(o, b)
    mw.multiply(prec, og, [(0.7, ga), (1.0, gb)], max_iter, abs_prec)
    orc.multiply(oc, [0.7, 1.0], [ca, cb], prec=prec, maxIter=max_iter, absPrec=abs_prec)
    A, B = og.to_arrays(), oc.to_arrays()
    assert A["scale"].shape == B["scale"].shape and np.array_equal(A["transl"], B["transl"]) and np.array_equal(A["child0"], B["child0"])
    nrm = np.sqrt((B["coefs"] ** 2).sum(axis=1))
    assert og.getNNodes() > 8 and nrm.max() > 1e-2
    assert (np.abs(A["coefs"] - B["coefs"]).max(axis=1) / nrm.max()).max() < 1e-10
    assert abs(og.getSquareNorm() - oc.getSquareNorm()) <= 1e-10 * oc.getSquareNorm()
    assert abs(og.integrate() - oc.integrate()) <= 1e-10 * max(abs(oc.integrate()), 1e-3)


@pytest.mark.parametrize("prec,start", [(1e-4, "roots"), (1e-5, "first")])
def test_multiply_max_norms(gpu, prec, start):
    """multiply(..., useMaxNorms = true) (multiply.cpp:112-115, MultiplicationAdaptor.h:46-66): the grid follows the largest scaling /
    wavelet norms of the inputs; node set and coefficients vs the oracle"""
    mw, orc = gpu
    mra = world(mw, 5)
    trees = []
    for n, seed in ((2, 71), (3, 72)):
        func = gaussians(n, seed, box=1.0, lo=1.0, hi=2.0)
        g, c = mw.FunctionTree(mra), mw.FunctionTree(mra)
        mw.project(1e-5, g, func)
        orc.project(1e-5, c, func)
        trees.append((g, c))
    (ga, ca), (gb, cb) = trees
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    if start == "first":
        mw.build_grid(og, ga)
        mw.build_grid(oc, ca)
    mw.multiply(prec, og, [(1.0, ga), (1.0, gb)], -1, True, True)
    orc.multiply(oc, [1.0, 1.0], [ca, cb], prec=prec, absPrec=True, useMaxNorms=True)
    A, B = og.to_arrays(), oc.to_arrays()
    assert A["scale"].shape == B["scale"].shape and np.array_equal(A["transl"], B["transl"]) and np.array_equal(A["child0"], B["child0"])
    nrm = np.sqrt((B["coefs"] ** 2).sum(axis=1))
    assert og.getNNodes() > 8 and (np.abs(A["coefs"] - B["coefs"]).max(axis=1) / nrm.max()).max() < 1e-10


def test_tree_algebra_vs_real_reference(gpu):
    """add (adaptive), multiply and divergence on the device against the REAL reference (oracle/_ref): node sets identical;
    coefficients within 1e-12 of the node norm for the linear operations, 1e-10 of the largest node norm for the product"""
    import ref_api as ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    mw, orc = gpu
    k, prec = 5, 1e-5
    wd = (k, -4, (-1, -1, -1), (2, 2, 2), 25)
    try:
        rm = ref.MRA(*wd)
    except OSError as e:
        pytest.skip(f"oracle/_ref does not load here: {e}")
    mra = mw.MultiResolutionAnalysis(*wd)
    rng = np.random.default_rng(3)
    trees = []
    for n in (2, 3):
        funcs = [mw.GaussFunc(b, (b / math.pi) ** 1.5 / n, tuple(rng.uniform(-1, 1, 3))) for b in 10.0 ** rng.uniform(1, 2, n)]
        rt, gt = ref.Tree(rm), mw.FunctionTree(mra)
        ref.project(prec, rt, funcs)
        e = mw.GaussExp()
        for f in funcs:
            e.append(f)
        mw.project(prec, gt, e, device=True)
        trees.append((rt, gt))
    (ra, ga), (rb, gb) = trees

    def compare(R, G, tol, floor):
        ri, gi = ref.by_index(R), ref.by_index(G)
        assert set(ri) == set(gi)
        nmax = max(np.linalg.norm(R["coefs"][i]) for i in ri.values())
        worst = max(np.abs(R["coefs"][i] - G["coefs"][gi[key]]).max() / max(np.linalg.norm(R["coefs"][i]), floor * nmax) for key, i in ri.items())
        assert worst < tol, worst

    ro, go = ref.Tree(rm), mw.FunctionTree(mra)
    ref.add(ro, [1.0, -2.0], [ra, rb], prec=1e-4)
    mw.add(1e-4, go, [(1.0, ga), (-2.0, gb)])
    compare(ro.export(), go.to_arrays(), COEF_TOL, 1e-3)
    ro, go = ref.Tree(rm), mw.FunctionTree(mra)
    ref.multiply(ro, [0.7, 1.0], [ra, rb], prec=1e-4)
    mw.multiply(1e-4, go, [(0.7, ga), (1.0, gb)])
    compare(ro.export(), go.to_arrays(), 1e-10, 1.0)
    RD, GD = ref.abgv(rm, 0.5, 0.5), mw.ABGVOperator(mra, 0.5, 0.5)
    ro, go = ref.Tree(rm), mw.FunctionTree(mra)
    ref.divergence(ro, RD, [ra, rb, ra])
    mw.divergence(go, GD, [(1.0, ga), (1.0, gb), (1.0, ga)])
    compare(ro.export(), go.to_arrays(), 1e-11, 1e-3)


@pytest.mark.parametrize("p,prec", [(2.0, 1e-4), (3.0, 1e-3)])
def test_power(gpu, p, prec):
    """power(prec, out, inp, p) (multiply.cpp:211-234): node set and coefficients (against the largest node norm) vs the oracle"""
    mw, orc = gpu
    mra = world(mw, 5)
    func = gaussians(2, 71, box=1.0, lo=1.0, hi=2.0)
    ga, ca = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(1e-5, ga, func)
    orc.project(1e-5, ca, func)
    og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.power(prec, og, ga, p)
    orc.power(oc, ca, p, prec=prec)
    A, B = og.to_arrays(), oc.to_arrays()
    assert A["scale"].shape == B["scale"].shape and np.array_equal(A["transl"], B["transl"]) and np.array_equal(A["child0"], B["child0"])
    nrm = np.sqrt((B["coefs"] ** 2).sum(axis=1))
    assert og.getNNodes() > 8 and (np.abs(A["coefs"] - B["coefs"]).max(axis=1) / nrm.max()).max() < 1e-10
