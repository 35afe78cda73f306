"""GPU tests (-m gpu) of the standalone node-level transforms behind MWNode::mwTransform(Compression | Reconstruction)
(src/trees/MWNode.cpp:557-594) and MWNode::cvTransform(Forward | Backward) (MWNode.cpp:448-490): mrx_node_mw_transform and
mrx_node_cv_transform on the resident node store, against
  - the oracle's tree-level transform (reconstruction of a branch node = the scaling blocks TopDown gives its children),
  - a numpy restatement of the interpolating coefficient-value map (diagonal: sqrt(1 / w_j) forward, InterpolatingBasis.cpp:115-124),
  - the analytic function values at the children's quadrature points (what cvTransform(Forward) means),
  - the round trips Compression(Reconstruction(x)) = x and Backward(Forward(x)) = x."""
import ctypes as C
import math

import numpy as np
import pytest

from test_gpu_parity import gaussians, world

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(libs):
    mw, orc = libs
    from mrcpp_b200 import _lib
    if _lib.device() is None or _lib.device() < 0:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return mw, orc


def quadrature(K):
    from mrcpp_b200 import _lib
    r, w = np.zeros(K), np.zeros(K)
    _lib.load().mrx_quadrature(K, r.ctypes.data_as(C.POINTER(C.c_double)), w.ctypes.data_as(C.POINTER(C.c_double)))
    return r, w


@pytest.mark.parametrize("k", [5, 7, 9, 4])
def test_node_mw_and_cv_transform(gpu, k):
    mw, orc = gpu
    K = k + 1
    Kd = K ** 3
    mra = world(mw, k)
    func = gaussians(3, 5)
    f = mw.FunctionTree(mra)
    mw.project(1e-4, f, func, device=True)
    A0 = f.to_arrays()
    n = len(A0["scale"])
    branch = np.nonzero(A0["child0"] >= 0)[0]
    leaves = np.nonzero(A0["child0"] < 0)[0]
    assert len(branch) > 0 and len(leaves) > 0
    nmax = np.sqrt((A0["coefs"] ** 2).sum(axis=1)).max()

    # ---- Reconstruction in the node: a branch node's 8 blocks become the scaling blocks of its children
    f.nodeMwTransform(mw.Reconstruction, branch)
    A1 = f.to_arrays()
    for p in branch[:200]:
        c0 = A0["child0"][p]
        want = np.stack([A0["coefs"][c0 + t][:Kd] for t in range(8)]).reshape(-1)
        assert np.abs(A1["coefs"][p] - want).max() <= 1e-12 * nmax
    assert np.array_equal(A1["coefs"][leaves], A0["coefs"][leaves])  # unlisted nodes untouched
    # ---- and back
    f.nodeMwTransform(mw.Compression, branch)
    A2 = f.to_arrays()
    assert np.abs(A2["coefs"] - A0["coefs"]).max() <= 1e-13 * nmax
    assert np.allclose(A2["norms"], A0["norms"], rtol=1e-10, atol=1e-13 * nmax)

    # ---- cvTransform(Forward) of every node after Reconstruction: function values at the children's quadrature points
    f.nodeMwTransform(mw.Reconstruction)
    R = f.to_arrays()
    f.nodeCvTransform(mw.Forward)
    V = f.to_arrays()
    roots, w = quadrature(K)
    m = np.sqrt(1.0 / w)
    # numpy restatement, same operation order: two_fac * (((c * m[x]) * m[y]) * m[z])
    cube = ((np.ones((K, K, K)) * m[None, None, :]) * m[None, :, None]) * m[:, None, None]  # [z][y][x], x fastest
    for i in list(leaves[:50]) + list(branch[:50]):
        two_fac = math.sqrt(2.0 ** (3 * (int(R["scale"][i]) + 1)))
        c = R["coefs"][i].reshape(8, K, K, K)
        want = two_fac * (((c * m[None, None, None, :]) * m[None, None, :, None]) * m[None, :, None, None])
        got = V["coefs"][i].reshape(8, K, K, K)
        assert np.allclose(got, want, rtol=4e-16, atol=0.0), np.abs(got - want).max()
    # meaning: values of the projected function at the quadrature points of the children of a leaf (to the projection precision)
    big = leaves[np.argsort(-np.sqrt((A0["coefs"][leaves] ** 2).sum(axis=1)))[:5]]
    for i in big:
        s, l = int(V["scale"][i]), V["transl"][i]
        got = V["coefs"][i].reshape(8, K, K, K)
        h = 2.0 ** (-(s + 1))
        for t in range(8):
            pts = np.stack(np.meshgrid(*[h * (roots + 2 * l[d] + ((t >> d) & 1)) for d in (2, 1, 0)], indexing="ij"), axis=-1)[..., ::-1]
            exact = func.evalf(pts.reshape(-1, 3)).reshape(K, K, K)
            assert np.abs(got[t] - exact).max() <= 2e-3 * max(np.abs(exact).max(), 1e-3), (i, t)
    # ---- Backward(Forward(x)) = x, then Compression restores the tree
    f.nodeCvTransform(mw.Backward)
    B = f.to_arrays()
    assert np.allclose(B["coefs"], R["coefs"], rtol=1e-14, atol=1e-15 * nmax)
    f.nodeMwTransform(mw.Compression)
    Z = f.to_arrays()
    assert np.abs(Z["coefs"] - A0["coefs"]).max() <= 1e-13 * nmax
    assert abs(f.getSquareNorm() - sum(0 for _ in ()) - f.getSquareNorm()) == 0.0
