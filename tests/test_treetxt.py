"""Text interchange (mrcpp_b200/treetxt.py) in the reference's saveTreeTXT format (src/trees/FunctionTree.cpp:306-372):
header, MADNESS level / translation / index conventions, and the values themselves: the end nodes of a projected tree hold the
function sampled at the expanded child quadrature points, so what the file contains must equal the analytic Gaussian there."""
import math
import os

import numpy as np


def test_save_tree_txt_matches_analytic_values(libs, tmp_path):
    mw, orc = libs
    from mrcpp_b200 import treetxt
    k, K, prec = 5, 6, 1e-4
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    beta, pos = 30.0, (0.4, -0.7, 1.1)
    f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, pos)
    t = mw.FunctionTree(mra)
    orc.project(prec, t, f)
    fname = os.path.join(tmp_path, "f.txt")
    treetxt.save_tree_txt(t, fname)
    # header as FunctionTree::saveTreeTXT writes it
    with open(fname) as fh:
        head = [fh.readline().split() for _ in range(6)]
    assert head[0] == ["3"] and head[1] == ["-16", "16"] and head[4] == [str(K)]
    assert int(head[5][0]) == 8 * t.getNEndNodes()
    Kr, blocks = treetxt.load_tree_txt(fname)
    assert Kr == K and len(blocks) == 8 * t.getNEndNodes()
    x01, _ = treetxt._quadrature(K)
    worst, peak = 0.0, (beta / math.pi) ** 1.5
    for (scale, lx, ly, lz), vals in blocks.items():
        h = 2.0 ** (-scale)
        X, Y, Z = h * (lx + x01), h * (ly + x01), h * (lz + x01)
        ana = peak * np.exp(-beta * ((Z[:, None, None] - pos[2]) ** 2 + (Y[None, :, None] - pos[1]) ** 2 + (X[None, None, :] - pos[0]) ** 2))
        worst = max(worst, float(np.abs(vals - ana).max()))
    # 14 significant digits in the file
    assert worst < 1e-12 * peak
    # every child block of the file is a child of an end node of the tree
    A = t.to_arrays(coefs=False)
    ends = {(int(A["scale"][n]), *map(int, A["transl"][n])) for n in np.nonzero(A["child0"] < 0)[0]}
    for (scale, lx, ly, lz) in blocks:
        assert (scale - 1, lx >> 1, ly >> 1, lz >> 1) in ends
