"""Text interchange of function trees in the reference's `saveTreeTXT` format (src/trees/FunctionTree.cpp:306-372).

SURVEY.md §8(c) names this format as the portable bridge to a real MRCPP build: the file holds, for every end node, the
function VALUES at the expanded quadrature points of its 8 children (MWNode::mwTransform(Reconstruction) followed by
MWNode::cvTransform(Forward), MWNode.cpp:448-594), written child by child with MADNESS conventions for level, translation and
index order. Host-side numpy only (nothing here is on the hot path): it works from `FunctionTree.to_arrays()`.

    save_tree_txt(tree, "f.txt")            # a file MRCPP's loadTreeTXT / MADNESS can read
    blocks = load_tree_txt("f.txt")         # {(scale, lx, ly, lz): values[K, K, K] (x index fastest, MRCPP order)}
    k, rscale, arrays = tree_arrays_from_txt("f.txt")   # -> FunctionTree.from_arrays(mra, *arrays), then mwTransform(BottomUp)

Pinned against the reference's own saveTreeTXT / loadTreeTXT in tests/test_reference_parity.py.
"""
import os
import struct

import numpy as np

_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mwtables.bin")


def _filters(k):
    """H0, G0 (row-major K x K) of order k from the packed tables; H1, G1 by the interpolating symmetry (MWFilter.cpp:229-235)"""
    with open(_TABLES, "rb") as f:
        assert f.read(4) == b"MRXT"
        (n,) = struct.unpack("<i", f.read(4))
        tabs = {}
        for _ in range(n):
            kind, kk, cnt = struct.unpack("<iii", f.read(12))
            buf = f.read(8 * cnt)
            if kk == k and kind in (0, 1):
                tabs[kind] = np.frombuffer(buf, dtype="<f8").reshape(k + 1, k + 1).copy()
    K = k + 1
    H0, G0 = tabs[0], tabs[1]
    i = np.arange(K)
    H1 = H0[::-1, ::-1]
    G1 = ((-1.0) ** (i + K))[:, None] * G0[:, ::-1]
    return H0, G0, H1, G1


def _quadrature(K):
    x, w = np.polynomial.legendre.leggauss(K)
    return 0.5 * (x + 1.0), 0.5 * w


def child_values(node_coefs, k, scale):
    """values[c, iz, iy, ix] at the quadrature points of the 8 children of a node at `scale`, from its 8 (s, d) blocks
    (block t bit d = wavelet along d, x index fastest): reconstruction s_b = H_b^T s + G_b^T d per dimension
    (math_utils::apply_filter), then coefficient -> value: divide by sqrt(w) per dimension, times 2^{3 (scale + 1) / 2}."""
    K = k + 1
    H0, G0, H1, G1 = _filters(k)
    X = np.empty((2, 2, K, K))  # X[t_bit, b_bit][i, j]: out_j = sum_i in_i X(i, j)
    X[0, 0], X[1, 0], X[0, 1], X[1, 1] = H0, G0, H1, G1
    c = np.asarray(node_coefs, dtype=np.float64).reshape(2, 2, 2, K, K, K)  # [tz, ty, tx, iz, iy, ix]
    out = np.einsum("abcijk,cgkn->abgijn", c, X)   # x:  sum over tx, ix -> (bx, jx)
    out = np.einsum("abgijn,bhjm->ahgimn", out, X)  # y
    out = np.einsum("ahgimn,afil->fhglmn", out, X)  # z  -> [bz, by, bx, jz, jy, jx]
    _, w = _quadrature(K)
    sw = np.sqrt(w)
    out = out / (sw[:, None, None] * sw[None, :, None] * sw[None, None, :])
    out = out * 2.0 ** (1.5 * (scale + 1))
    return out.reshape(8, K, K, K)


def _madness_order(K):
    """mapMRC of saveTreeTXT: MADNESS writes z fastest ... with descending indices (FunctionTree.cpp:335-342)"""
    m = []
    for x in range(K - 1, -1, -1):
        for y in range(K - 1, -1, -1):
            for z in range(K - 1, -1, -1):
                m.append(z * K * K + y * K + x)
    return np.array(m)


def save_tree_txt(tree, fname):
    """FunctionTree::saveTreeTXT for a 3-D tree on a cubic world of 2 x 2 x 2 root boxes centred at the origin"""
    mra = tree.mra
    A = tree.to_arrays()
    K, k = mra.kp1, mra.kp1 - 1
    rscale = int(A["scale"].min())
    L = 2.0 ** (-rscale)
    ends = np.nonzero(A["child0"] < 0)[0]
    order = _madness_order(K)
    with open(fname, "w") as out:
        out.write("3\n")
        for _ in range(3):
            out.write(f"{-L:.14g} {L:.14g}\n")
        out.write(f"{K}\n{8 * len(ends)}\n")
        for n in ends:
            scale = int(A["scale"][n])
            vals = child_values(A["coefs"][n], k, scale)
            half = int(round(2.0 ** scale * L))  # 2^scale L is integral for every scale >= root scale
            l0 = [2 * (int(A["transl"][n, d]) + half) for d in range(3)]  # MADNESS translations start at 0
            for c in range(8):
                out.write(f"{scale - rscale + 2} " + " ".join(str(l0[d] + ((c >> d) & 1)) for d in range(3)) + " \n")
                flat = vals[c].reshape(-1)
                out.write(" ".join(f"{v:.14g}" for v in flat[order]) + " \n")


def load_tree_txt(fname):
    """-> (K, {(scale, lx, ly, lz): values[iz, iy, ix]}) with MRCPP's scale / translation / index conventions restored"""
    with open(fname) as f:
        tok = f.read().split()
    pos = 0
    D = int(tok[pos]); pos += 1
    assert D == 3
    lo = float(tok[pos]); pos += 2 * D
    K = int(tok[pos]); pos += 1
    nblk = int(tok[pos]); pos += 1
    rscale = -int(round(np.log2(-lo)))
    L = 2.0 ** (-rscale)
    inv = np.argsort(_madness_order(K))
    blocks = {}
    for _ in range(nblk):
        lev = int(tok[pos]); lm = [int(t) for t in tok[pos + 1:pos + 4]]; pos += 4
        vals = np.array(tok[pos:pos + K ** 3], dtype=np.float64); pos += K ** 3
        scale = lev + rscale - 1          # child scale
        shift = int(round(2.0 ** scale * L))
        blocks[(scale, lm[0] - shift, lm[1] - shift, lm[2] - shift)] = vals[inv].reshape(K, K, K)
    return K, blocks


def node_coefs_from_child_values(vals, k, scale):
    """inverse of child_values: the 8 (s, d) blocks of a node at `scale` from the values at the quadrature points of its 8 children
    (MWNode::cvTransform(Backward) then MWNode::mwTransform(Compression), MWNode.cpp:448-594)"""
    K = k + 1
    H0, G0, H1, G1 = _filters(k)
    X = np.empty((2, 2, K, K))
    X[0, 0], X[1, 0], X[0, 1], X[1, 1] = H0, G0, H1, G1
    _, w = _quadrature(K)
    sw = np.sqrt(w)
    s = np.asarray(vals, dtype=np.float64).reshape(2, 2, 2, K, K, K)  # [bz, by, bx, jz, jy, jx]
    s = s * (sw[:, None, None] * sw[None, :, None] * sw[None, None, :]) * 2.0 ** (-1.5 * (scale + 1))
    # the two-scale filter matrix is orthogonal: compression is the transpose of the reconstruction of child_values
    out = np.einsum("fhglmn,afil->ahgimn", s, X)    # z
    out = np.einsum("ahgimn,bhjm->abgijn", out, X)  # y
    out = np.einsum("abgijn,cgkn->abcijk", out, X)  # x -> [tz, ty, tx, iz, iy, ix]
    return out.reshape(8 * K ** 3)


def tree_arrays_from_txt(fname, corner=(-1, -1, -1), boxes=(2, 2, 2)):
    """Arrays for FunctionTree.from_arrays (scale, transl, parent, child0, coefs in slot order) of the tree a saveTreeTXT file
    describes: the end nodes are the parents of the eight-block groups of the file; branch nodes carry zeros and are filled by the
    mwTransform(BottomUp) the caller runs afterwards, like FunctionTree::loadTreeTXT does (FunctionTree.cpp:240-305). Files
    written by MRCPP itself (complete sibling groups)."""
    K, blocks = load_tree_txt(fname)
    k = K - 1
    ends = {}
    for (scale, lx, ly, lz), v in blocks.items():
        key = (scale - 1, lx >> 1, ly >> 1, lz >> 1)
        ends.setdefault(key, np.zeros((8, K, K, K)))[(lx & 1) | ((ly & 1) << 1) | ((lz & 1) << 2)] = v
    rscale = min(key[0] for key in ends)
    while any(not (corner[d] <= (key[1 + d] >> (key[0] - rscale)) < corner[d] + boxes[d]) for key in ends for d in range(3)):
        rscale -= 1  # (only if the file has no end node at the root scale and the guess was too fine)
    branches = set()
    for (s, lx, ly, lz) in ends:
        while s > rscale:
            s, lx, ly, lz = s - 1, lx >> 1, ly >> 1, lz >> 1
            branches.add((s, lx, ly, lz))
    scale, transl, parent, child0 = [], [], [], []
    for r in range(boxes[0] * boxes[1] * boxes[2]):  # roots in box order, x fastest
        scale.append(rscale)
        transl.append((corner[0] + r % boxes[0], corner[1] + (r // boxes[0]) % boxes[1], corner[2] + r // (boxes[0] * boxes[1])))
        parent.append(-1)
        child0.append(-1)
    n = 0
    while n < len(scale):  # children of every branch node, appended in creation order
        key = (scale[n], *transl[n])
        if key in branches:
            child0[n] = len(scale)
            for c in range(8):
                scale.append(scale[n] + 1)
                transl.append(tuple(2 * transl[n][d] + ((c >> d) & 1) for d in range(3)))
                parent.append(n)
                child0.append(-1)
        n += 1
    coefs = np.zeros((len(scale), 8 * K ** 3))
    for i in range(len(scale)):
        key = (scale[i], *transl[i])
        if key in ends:
            coefs[i] = node_coefs_from_child_values(ends[key], k, scale[i])
    return k, rscale, (np.array(scale, np.int32), np.array(transl, np.int32), np.array(parent, np.int32), np.array(child0, np.int32), coefs)
