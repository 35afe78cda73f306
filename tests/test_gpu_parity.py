"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the
same seeded inputs. Bars: node set and indexing bit-exact; per-node coefficients within 1e-12 relative
to the node norm (BASELINE.json north_star); screened tuple count identical."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COEF_TOL = 1e-12  # relative to the node norm


@pytest.fixture(scope="module")
def gpu(libs):
    mw, orc = libs
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return mw, orc


def assert_same_tree(a, b, tol=COEF_TOL, floor=1e-3):
    """node set, indexing and slot order bit-exact; coefficients within tol of the node norm. The assertion is on the floored
    figure (nodes below `floor` x the largest node norm are measured against that floor); the strict per-node figure and the
    number of nodes that needed the floor are printed and recorded (tests/parity_util.py)."""
    from parity_util import coef_parity
    A, B = a.to_arrays(), b.to_arrays()
    assert A["scale"].shape == B["scale"].shape
    assert np.array_equal(A["scale"], B["scale"])
    assert np.array_equal(A["transl"], B["transl"])
    assert np.array_equal(A["parent"], B["parent"])
    assert np.array_equal(A["child0"], B["child0"])
    rep = coef_parity(A["coefs"], B["coefs"], tol=tol, floor=floor)
    assert rep["floored"] < tol, rep
    nmax = float(np.sqrt((B["coefs"] ** 2).sum(axis=1)).max()) if len(B["coefs"]) else 0.0
    assert np.allclose(A["norms"], B["norms"], rtol=1e-10, atol=1e-12 * nmax)
    return rep["floored"]


def gaussians(n, seed, box=8.0, lo=1.0, hi=3.0):
    import mrcpp_b200 as mw
    rng = np.random.default_rng(seed)
    g = mw.GaussExp()
    for _ in range(n):
        beta = 10.0 ** rng.uniform(lo, hi)
        g.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-box, box, 3))))
    return g


def world(mw, k):
    return mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)


@pytest.mark.parametrize("k", [3, 5, 7, 9])
def test_bottom_up_and_norms(gpu, k):
    """project: host quadrature + DEVICE mwTransform(BottomUp)/calcSquareNorm vs oracle transform."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(3, 7)
    a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(1e-4, a, func)
    orc.project(1e-4, b, func)
    assert_same_tree(a, b)
    assert abs(a.getSquareNorm() - b.getSquareNorm()) <= 1e-13 * b.getSquareNorm()


@pytest.mark.parametrize("k,n,prec", [(5, 3, 1e-4), (7, 6, 1e-5), (9, 2, 1e-4), (6, 2, 1e-4)])
def test_device_projection(gpu, k, n, prec):
    """project with the quadrature, cvTransform, in-node compression and norms on the DEVICE vs the oracle
    projection: same node set; coefficients within 1e-12 of the node norm (device exp() and glibc exp() differ by <= 1 ulp);
    the projected density integrates to ~1 through its Coulomb/self-overlap invariants (norm equal to the oracle's)."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(n, 23)
    a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, a, func, device=True)
    orc.project(prec, b, func)
    assert_same_tree(a, b)
    assert abs(a.getSquareNorm() - b.getSquareNorm()) <= 1e-13 * b.getSquareNorm()
    # the device-born tree feeds the apply like any other
    P = mw.PoissonOperator(mra, prec)
    ga, gb = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sa = mw.apply(prec, ga, P, a)
    sb = orc.apply(prec, gb, P, b)
    assert sa.f_applied == sb.fApplied
    assert_same_tree(ga, gb)


@pytest.mark.parametrize("k,prec", [(7, 1e-5), (5, 1e-4)])
def test_apply_input_in_host_memory(gpu, k, prec):
    """Input tree in (pinned) host memory, as a reference-side binding hands it over: the apply gathers only the nodes it
    reads (lazy residency) and must give bit-identical results to the apply on a fully resident input; a second apply on
    the partially resident tree and a whole-tree operation afterwards (which completes the upload) stay correct."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(5, 31)
    P = mw.PoissonOperator(mra, prec)
    f = mw.FunctionTree(mra)
    mw.project(prec, f, func)
    ref = mw.FunctionTree(mra)
    s0 = mw.apply(prec, ref, P, f)
    assert s0.h2d_bytes <= 16 * f.getNNodes()  # resident input: at most the band-walk topology (16 B per node) is uploaded
    R = ref.to_arrays()
    f.drop_device()                       # host copy is the only copy now
    g = mw.FunctionTree(mra)
    s1 = mw.apply(prec, g, P, f)
    G = g.to_arrays()
    assert s1.f_applied == s0.f_applied
    assert np.array_equal(G["transl"], R["transl"]) and np.array_equal(G["coefs"], R["coefs"])
    assert 0 < s1.h2d_bytes < f.nbytes()   # only part of the input crossed PCIe
    g2 = mw.FunctionTree(mra)
    s2 = mw.apply(prec, g2, P, f)          # partially resident input: nothing (or little) left to fetch
    assert np.array_equal(g2.to_arrays()["coefs"], R["coefs"]) and s2.h2d_bytes <= s1.h2d_bytes
    D = mw.ABGVOperator(mra, 0.5, 0.5)
    d1, d2 = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.apply(None, d1, D, f, dir=1)        # derivative apply on the partially resident tree
    before = mw.dot(f, f)                  # whole-tree consumer: completes the upload
    mw.apply(None, d2, D, f, dir=1)
    assert np.array_equal(d1.to_arrays()["coefs"], d2.to_arrays()["coefs"])
    assert abs(before - f.getSquareNorm()) <= 1e-12 * f.getSquareNorm()


@pytest.mark.parametrize("k", [5, 7])
def test_top_down_roundtrip(gpu, k):
    """mwTransform(TopDown) then (BottomUp) on the device: TopDown(overwrite) parity vs oracle and the
    size-independent property compress(reconstruct(x)) == x."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(2, 11)
    a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(1e-4, a, func)
    orc.project(1e-4, b, func)
    before = a.to_arrays()
    a.mwTransform(mw.TopDown, overwrite=False)  # children.s += reconstruct(parent)
    orc.mw_transform_down(b, overwrite=False)
    assert_same_tree(a, b)
    a.mwTransform(mw.BottomUp)
    orc.mw_transform_up(b)
    assert_same_tree(a, b)
    after = a.to_arrays()
    # leaves were doubled at most levels; roots re-compressed consistently: norms finite, grid unchanged
    assert np.array_equal(before["transl"], after["transl"])


@pytest.mark.parametrize("k,prec", [(5, 1e-3), (7, 1e-5), (9, 1e-4), (3, 1e-2), (11, 1e-4), (4, 1e-3), (6, 1e-4), (8, 1e-4), (10, 1e-4)])
def test_poisson_apply_adaptive(gpu, k, prec):
    """examples/poisson.cpp (C1): adaptive apply; node set, tuple count, coefficients, energy."""
    mw, orc = gpu
    mra = world(mw, k)
    beta = 100.0
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
    P = mw.PoissonOperator(mra, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, f)
    orc.project(prec, fc, f)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, P, fg)
    sc = orc.apply(prec, gc, P, fc)
    assert sg.f_applied == sc.fApplied
    assert sg.g_nodes == sc.gNodes and sg.iterations == sc.iters
    assert_same_tree(gg, gc)
    assert abs(gg.getSquareNorm() - gc.getSquareNorm()) <= 1e-12 * gc.getSquareNorm()
    en = mw.dot(gg, fg)
    assert abs(en - orc.dot(gc, fc)) <= 1e-12 * abs(en)
    assert abs(en - 7.978845608) / 7.978845608 < (prec if k >= 5 else 5 * prec)
    assert fg.getNNodes() == fc.getNNodes()  # generated nodes were cleaned up (apply.cpp:86)


@pytest.mark.parametrize("k,prec", [(5, 1e-3), (7, 1e-4)])
def test_poisson_apply_fixed_grid(gpu, k, prec):
    """parity mode A: fixed output grid, maxIter=0, no norm screening -> 1e-12 coefficients."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(3, 5)
    P = mw.PoissonOperator(mra, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func)
    orc.project(prec, fc, func)
    ref = mw.FunctionTree(mra)
    orc.apply(prec, ref, P, fc)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.copy_grid(gg, ref)
    mw.copy_grid(gc, ref)
    sg = mw.apply(prec, gg, P, fg, maxIter=0)
    sc = orc.apply(prec, gc, P, fc, maxIter=0)
    assert sg.f_applied == sc.fApplied and sg.g_nodes == ref.getNNodes()
    assert_same_tree(gg, gc)


@pytest.mark.parametrize("k,prec,max_iter,abs_prec", [(7, 1e-4, 2, False), (7, 1e-4, -1, True), (5, 1e-3, 1, True), (4, 1e-3, -1, False)])
def test_apply_variants(gpu, k, prec, max_iter, abs_prec):
    """maxIter-limited refinement (TreeBuilder.cpp:73), absolute precision (tree_utils.cpp:55-57) and an even order
    (k = 4 -> K = 5: odd K, element-wise partial blocks and the FMA transform kernels)."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(3, 21, box=4.0, lo=1.0, hi=2.0)
    P = mw.PoissonOperator(mra, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func)
    orc.project(prec, fc, func)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, P, fg, maxIter=max_iter, absPrec=abs_prec)
    sc = orc.apply(prec, gc, P, fc, maxIter=max_iter, absPrec=abs_prec)
    assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes and sg.iterations == sc.iters
    assert_same_tree(gg, gc)


def test_apply_refines_prebuilt_grid(gpu):
    """`out` enters with a pre-built grid (apply.cpp:63-66 keeps it as the starting grid) and is refined further: the first
    work vector holds branch and leaf nodes of mixed depths, only the leaves may split, children of several depths share the
    later work vectors. Node set, tuple count and coefficients vs the oracle."""
    mw, orc = gpu
    prec, k = 1e-5, 7
    mra = world(mw, k)
    func = gaussians(4, 77, box=4.0, lo=1.0, hi=2.0)
    P = mw.PoissonOperator(mra, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func)
    orc.project(prec, fc, func)
    coarse = mw.FunctionTree(mra)
    orc.apply(prec, coarse, P, fc, maxIter=2)  # a shallow adaptive grid to start from (roots + two refinement levels)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.copy_grid(gg, coarse)
    mw.copy_grid(gc, coarse)
    assert gg.getNNodes() > 8
    sg = mw.apply(prec, gg, P, fg)
    sc = orc.apply(prec, gc, P, fc)
    assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes and sg.iterations == sc.iters
    assert gg.getNNodes() > coarse.getNNodes()
    assert_same_tree(gg, gc)


def test_identity_convolution_noncubic_world(gpu):
    """The reference's identity-convolution case (tests/operators/identity_convolution.cpp, 3D): generic single-term Gaussian
    kernel through ConvolutionOperator, k = 5, a 1 x 2 x 3 world at root scale 1 with corner (-1, 0, 1). GPU vs oracle."""
    mw, orc = gpu
    proj_prec, apply_prec, build_prec = 1e-3, 1e-3, 1e-4
    mra = mw.MultiResolutionAnalysis(5, 1, (-1, 0, 1), (1, 2, 3), 25)
    beta = 1.0e4
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (-0.2, 0.5, 1.0))
    expo = math.sqrt(1.0 / (build_prec / 10.0))
    I = mw.ConvolutionOperator(mra, [(expo / math.pi) ** 1.5], [expo], build_prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(proj_prec, fg, f)
    orc.project(proj_prec, fc, f)
    assert_same_tree(fg, fc)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(apply_prec, gg, I, fg)
    sc = orc.apply(apply_prec, gc, I, fc)
    assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes
    assert_same_tree(gg, gc)
    assert gg.getNNodes() <= fg.getNNodes()
    # device projection on the same world
    fd = mw.FunctionTree(mra)
    mw.project(proj_prec, fd, f, device=True)
    assert_same_tree(fd, fc)


def test_hydrogen_fixed_point_helmholtz(gpu):
    """The reference's Helmholtz apply case (tests/operators/helmholtz_operator.cpp): psi_1s = -1/(2 pi) H_1 [V psi_1s] on
    [-32, 32]^3, k = 5, inputs projected through the callback projection, output refined from psi's grid. GPU vs oracle,
    and the norm of the right-hand side is 1 within apply_prec."""
    mw, orc = gpu
    proj_prec, apply_prec, build_prec = 3.0e-3, 3.0e-2, 3.0e-3
    mra = mw.MultiResolutionAnalysis(5, -5, (-1, -1, -1), (2, 2, 2), 25)
    c = 1.0 / math.sqrt(math.pi)

    def psi(x, y, z):
        return c * math.exp(-math.sqrt(x * x + y * y + z * z))

    def vpsi(x, y, z):
        r = math.sqrt(x * x + y * y + z * z)
        return -c * math.exp(-r) / r

    pg, vg = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project_function(proj_prec, pg, psi)      # host quadrature + device BottomUp
    mw.project_function(proj_prec, vg, vpsi)
    pc, vc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    for t, f in ((pc, psi), (vc, vpsi)):
        mw.project_function(proj_prec, t, f, finalize=False)
        orc.mw_transform_up(t)
        orc.calc_square_norm(t)
    assert_same_tree(vg, vc)
    H = mw.HelmholtzOperator(mra, 1.0, build_prec)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.copy_grid(gg, pg)
    mw.copy_grid(gc, pc)
    sg = mw.apply(apply_prec, gg, H, vg)
    sc = orc.apply(apply_prec, gc, H, vc)
    assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes
    assert_same_tree(gg, gc)
    gg.rescale(-1.0 / (2.0 * math.pi))
    assert abs(math.sqrt(gg.getSquareNorm()) - 1.0) < apply_prec
    assert abs(mw.dot(gg, pg) - 1.0) < apply_prec


def test_multi_center_density(gpu):
    """10 seeded Gaussians, k=7: adaptive parity + pairwise analytic Coulomb energy (SURVEY §8c KAT 3)."""
    mw, orc = gpu
    prec = 1e-5
    mra = world(mw, 7)
    func = gaussians(10, 42, box=4.0, lo=1.0, hi=2.0)
    P = mw.PoissonOperator(mra, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func)
    orc.project(prec, fc, func)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, P, fg)
    sc = orc.apply(prec, gc, P, fc)
    assert sg.f_applied == sc.fApplied
    assert_same_tree(gg, gc)
    ana = sum(a.calc_coulomb_energy(b) for a in func for b in func)
    en = mw.dot(gg, fg)
    assert abs(en - ana) / ana < 10 * prec


def test_helmholtz_apply(gpu):
    """HelmholtzOperator (mu=1) apply, k=5: parity vs oracle."""
    mw, orc = gpu
    prec = 1e-4
    mra = world(mw, 5)
    func = gaussians(4, 3, box=3.0, lo=0.5, hi=1.5)
    H = mw.HelmholtzOperator(mra, 1.0, prec)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func)
    orc.project(prec, fc, func)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, H, fg)
    sc = orc.apply(prec, gc, H, fc)
    assert sg.f_applied == sc.fApplied
    assert_same_tree(gg, gc)


def test_helmholtz_orbital_k9(gpu):
    """C4 shape: HelmholtzOperator mu=1 at k=9 on a benzene-like 12-centre orbital (6 centres at radius
    2.64, 6 at 4.69 bohr, z=0), two independent trees (the 50-tree batch shards by tree)."""
    mw, orc = gpu
    prec = 1e-4
    mra = world(mw, 9)
    H = mw.HelmholtzOperator(mra, 1.0, prec)
    for j in range(2):
        rng = np.random.default_rng(2024 + j)
        func = mw.GaussExp()
        for ring, (rad, beta) in enumerate(((2.64, 1.5), (4.69, 0.8))):
            for a in range(6):
                ang = math.pi / 3 * a
                func.append(mw.GaussFunc(beta, float(rng.normal()), (rad * math.cos(ang), rad * math.sin(ang), 0.0)))
        fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        mw.project(prec, fg, func)
        orc.project(prec, fc, func)
        gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        sg = mw.apply(prec, gg, H, fg)
        sc = orc.apply(prec, gc, H, fc)
        assert sg.f_applied == sc.fApplied
        assert_same_tree(gg, gc)
        gg.rescale(-1.0 / (2.0 * math.pi))  # scf.cpp:108 pattern
        assert abs(gg.getSquareNorm() - gc.getSquareNorm() / (2.0 * math.pi) ** 2) <= 1e-12 * gg.getSquareNorm()


@pytest.mark.parametrize("a,b", [(0.5, 0.5), (0.0, 0.0)])
def test_abgv_derivative(gpu, a, b):
    """C3: apply(out, ABGVOperator, inp, dir) on the device vs oracle for dir 0,1,2 (fixed, widened grid)."""
    mw, orc = gpu
    prec = 1e-5
    mra = world(mw, 7)
    func = gaussians(4, 1234, box=4.0, lo=0.0, hi=2.0)
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, func)
    orc.project(prec, fc, func)
    D = mw.ABGVOperator(mra, a, b)
    for d in range(3):
        og, oc = mw.FunctionTree(mra), mw.FunctionTree(mra)
        sg = mw.apply(None, og, D, fg, dir=d)
        sc = orc.apply_derivative(oc, D, fc, d)
        assert sg.f_applied == sc.fApplied and sg.g_nodes == sc.gNodes
        assert_same_tree(og, oc)


def test_empty_and_edge_inputs(gpu):
    """edge cases: zero function (every f component norm < MachineZero -> nothing applied), function
    hugging the world border (band clipping), maxIter=0 on bare roots."""
    mw, orc = gpu
    prec = 1e-3
    mra = world(mw, 5)
    P = mw.PoissonOperator(mra, prec)
    zero = mw.FunctionTree(mra)
    zero.mwTransform(mw.BottomUp)  # roots only, all-zero coefficients
    out = mw.FunctionTree(mra)
    st = mw.apply(prec, out, P, zero)
    assert st.f_applied == 0 and out.getNNodes() == 8
    assert out.getSquareNorm() == 0.0
    beta = 50.0
    edge = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (15.5, -15.5, 15.0))
    fg, fc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.project(prec, fg, edge)
    orc.project(prec, fc, edge)
    gg, gc = mw.FunctionTree(mra), mw.FunctionTree(mra)
    sg = mw.apply(prec, gg, P, fg)
    sc = orc.apply(prec, gc, P, fc)
    assert sg.f_applied == sc.fApplied
    assert_same_tree(gg, gc)
    g0, c0 = mw.FunctionTree(mra), mw.FunctionTree(mra)
    s0 = mw.apply(prec, g0, P, fg, maxIter=0)
    t0 = orc.apply(prec, c0, P, fc, maxIter=0)
    assert s0.g_nodes == 8 and s0.f_applied == t0.fApplied
    assert_same_tree(g0, c0)


def test_linearity_property(gpu):
    """size-independent property: apply is linear -> P(2f) == 2 P(f) on a fixed grid. Scaling by a power of two is
    exact in FP64, so the only differences come from the reference's ABSOLUTE |f_ft| < MachineZero skip
    (ConvolutionCalculator.cpp:254), i.e. contributions below 1e-14 * |O|."""
    mw, orc = gpu
    prec = 1e-4
    mra = world(mw, 7)
    func = gaussians(5, 9, box=5.0)
    P = mw.PoissonOperator(mra, prec)
    f1 = mw.FunctionTree(mra)
    mw.project(prec, f1, func)
    g1 = mw.FunctionTree(mra)
    mw.apply(prec, g1, P, f1)
    f2 = mw.FunctionTree(mra)
    mw.project(prec, f2, func)
    f2.rescale(2.0)
    a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.copy_grid(a, g1)
    mw.copy_grid(b, g1)
    mw.apply(prec, a, P, f1, maxIter=0)
    mw.apply(prec, b, P, f2, maxIter=0)
    A, B = a.to_arrays(), b.to_arrays()
    nrm = np.sqrt((B["coefs"] ** 2).sum(axis=1))
    err = np.abs(2.0 * A["coefs"] - B["coefs"]).max(axis=1)
    assert (err / np.maximum(nrm, 1e-3 * nrm.max())).max() < 1e-11


def test_golden_fixture(gpu):
    """committed golden vectors (tests/golden/make_golden.py, generated with the oracle in the build container)."""
    mw, orc = gpu
    path = os.path.join(os.path.dirname(__file__), "golden", "poisson_small.npz")
    gold = np.load(path)
    k, prec = int(gold["k"]), float(gold["prec"])
    mra = world(mw, k)
    beta = float(gold["beta"])
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, tuple(gold["pos"]))
    P = mw.PoissonOperator(mra, prec)
    assert P.size() == int(gold["n_terms"])
    ft = mw.FunctionTree(mra)
    mw.project(prec, ft, f)
    F = ft.to_arrays()
    assert np.array_equal(F["transl"], gold["f_transl"]) and np.array_equal(F["scale"], gold["f_scale"])
    assert np.abs(F["coefs"][:8] - gold["f_coefs_head"]).max() <= 1e-13 * np.abs(gold["f_coefs_head"]).max()
    assert np.allclose(F["norms"], gold["f_norms"], rtol=1e-12, atol=1e-16)
    gt = mw.FunctionTree(mra)
    st = mw.apply(prec, gt, P, ft)
    G = gt.to_arrays()
    assert st.f_applied == int(gold["f_applied"])
    assert np.array_equal(G["transl"], gold["g_transl"]) and np.array_equal(G["scale"], gold["g_scale"])
    nrm = np.sqrt((gold["g_coefs"] ** 2).sum(axis=1))
    err = np.abs(G["coefs"] - gold["g_coefs"]).max(axis=1)
    assert (err / np.maximum(nrm, 1e-300)).max() < COEF_TOL


def test_full_size_coulomb_energy(gpu):
    """BASELINE.json's full-size workload (k=7, prec 1e-7, 100-centre density of bench.py) through a size-independent
    property: <f|P f> must equal the analytic pairwise Coulomb energy of the Gaussian expansion
    (GaussFunc::calcCoulombEnergy, src/functions/GaussFunc.cpp:210-237), and a second apply must reproduce the first
    bit for bit (fixed summation order)."""
    mw, orc = gpu
    prec = 1e-7
    mra = world(mw, 7)
    func = gaussians(100, 42, box=8.0, lo=1.0, hi=3.0)
    P = mw.PoissonOperator(mra, prec)
    f = mw.FunctionTree(mra)
    mw.project(prec, f, func)
    g1, g2 = mw.FunctionTree(mra), mw.FunctionTree(mra)
    s1 = mw.apply(prec, g1, P, f)
    s2 = mw.apply(prec, g2, P, f)
    assert s1.f_applied == s2.f_applied and s1.g_nodes == s2.g_nodes
    assert g1.getSquareNorm() == g2.getSquareNorm()
    A, B = g1.to_arrays(), g2.to_arrays()
    assert np.array_equal(A["transl"], B["transl"]) and np.array_equal(A["coefs"], B["coefs"])
    ana = sum(a.calc_coulomb_energy(b) for a in func for b in func)
    en = mw.dot(g1, f)
    assert abs(en - ana) / ana < 10 * prec
    assert f.getNNodes() > 40000 and s1.f_applied > 2e7  # really the full-size case


def test_end_to_end_path_sub_ranges_and_gather_beside_the_contraction(gpu, monkeypatch):
    """The two overlaps of the end-to-end path at k = 7: a large iteration runs in node sub-ranges; (1) with a
    host-resident input the fill pass and the gather of its blocks are per sub-range, the gather of sub-range s + 1 on its own stream
    beside the contraction of sub-range s; (2) with a mirrored output contraction / reduce / TopDown step / download are per
    sub-range. Both must leave every bit of the result as the plain apply on a resident input gives it, in every combination,
    and on the partially resident tree afterwards."""
    mw, orc = gpu
    k, prec = 7, 1e-6
    mra = world(mw, k)
    func = gaussians(40, 2026)
    P = mw.PoissonOperator(mra, prec)
    f = mw.FunctionTree(mra)
    mw.project(prec, f, func, device=True)
    ref = mw.FunctionTree(mra)
    s0 = mw.apply(prec, ref, P, f)
    R = ref.to_arrays()
    f.to_arrays()  # host copy of the input
    monkeypatch.setenv("MRX_SUB_MIN_TILES", "1")  # iterations of >= 512 nodes run in sub-ranges of >= 256 nodes
    monkeypatch.setenv("MRX_SUB_RANGES", "8")
    for overlap, subs in ((True, True), (False, True), (True, False)):
        if overlap: monkeypatch.delenv("MRX_NO_FETCH_OVERLAP", raising=False)
        else: monkeypatch.setenv("MRX_NO_FETCH_OVERLAP", "1")
        monkeypatch.setenv("MRX_SUB_RANGES", "8" if subs else "1")
        f.drop_device()
        g = mw.FunctionTree(mra)
        g.set_host_mirror(True)
        s1 = mw.apply(prec, g, P, f)
        assert s1.f_applied == s0.f_applied and 0 < s1.h2d_bytes < f.nbytes()
        G = g.to_arrays()
        assert np.array_equal(G["transl"], R["transl"]) and np.array_equal(G["coefs"], R["coefs"]) and np.array_equal(G["norms"], R["norms"])
        g2 = mw.FunctionTree(mra)  # partially resident input, plain output
        s2 = mw.apply(prec, g2, P, f)
        assert s2.h2d_bytes <= s1.h2d_bytes and np.array_equal(g2.to_arrays()["coefs"], R["coefs"])
    assert max(int(x) for x in np.bincount(R["scale"] - R["scale"].min())) >= 512  # the sub-range path did run


@pytest.mark.parametrize("k,n,prec", [(7, 6, 1e-5), (5, 3, 1e-4), (9, 2, 1e-4)])
def test_apply_with_host_mirror(gpu, k, n, prec):
    """mrx_tree_set_host_mirror: the apply streams its result into the output tree's pinned host chunks while it runs (wavelet
    blocks per refinement iteration on the copy engines, scaling blocks and branch nodes behind the closing passes). The host
    copy must be bit-identical to a download after a plain apply -- on bare roots, on a pre-built grid (fixed-grid mode) and when
    the same tree object is applied onto again -- and the block-granular lazy gather of a host-resident input must not change it."""
    mw, orc = gpu
    mra = world(mw, k)
    func = gaussians(n, 77)
    P = mw.PoissonOperator(mra, prec)
    f = mw.FunctionTree(mra)
    mw.project(prec, f, func, device=True)
    ref = mw.FunctionTree(mra)
    s0 = mw.apply(prec, ref, P, f)
    R = ref.to_arrays()
    g = mw.FunctionTree(mra)
    g.set_host_mirror(True)
    s1 = mw.apply(prec, g, P, f)
    assert s1.d2h_bytes == g.nbytes() > 0 and s0.d2h_bytes == 0
    G = g.to_arrays()  # no download left to do: the arrays come from the mirrored host chunks
    assert np.array_equal(G["transl"], R["transl"]) and np.array_equal(G["coefs"], R["coefs"]) and np.array_equal(G["norms"], R["norms"])
    # fixed-grid mode on a pre-built grid (arbitrary slots in the first work vector)
    h0, h1 = mw.FunctionTree(mra), mw.FunctionTree(mra)
    mw.copy_grid(h0, ref)
    mw.copy_grid(h1, ref)
    h1.set_host_mirror(True)
    mw.apply(prec, h0, P, f, maxIter=0)
    mw.apply(prec, h1, P, f, maxIter=0)
    assert np.array_equal(h0.to_arrays()["coefs"], h1.to_arrays()["coefs"])
    # host-resident input (block-granular gather) + mirrored output: the end-to-end path of bench.py
    f.drop_device()
    e = mw.FunctionTree(mra)
    e.set_host_mirror(True)
    s2 = mw.apply(prec, e, P, f)
    assert 0 < s2.h2d_bytes < f.nbytes()
    assert np.array_equal(e.to_arrays()["coefs"], R["coefs"])
    # the mirrored tree is a valid input afterwards (host and device copies agree)
    assert abs(mw.dot(e, e) - mw.dot(ref, ref)) <= 1e-13 * mw.dot(ref, ref)
