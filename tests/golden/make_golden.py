"""Generates tests/golden/*.npz with the CPU oracle (run in the build container: python tests/golden/make_golden.py).
The reference itself cannot be built here (no Eigen), so these vectors pin the oracle's output across
machines/compilers; the oracle in turn is pinned by the reference's KATs (tests/test_oracle_kats.py)."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mrcpp_b200 as mw  # noqa: E402
from mrcpp_b200 import _lib  # noqa: E402

_lib.init()
import oracle_api as orc  # noqa: E402

k, prec, beta, pos = 5, 1e-2, 10.0, (0.7, -0.4, 0.2)
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
f = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, pos)
P = mw.PoissonOperator(mra, prec)
ft = mw.FunctionTree(mra)
orc.project(prec, ft, f)
gt = mw.FunctionTree(mra)
st = orc.apply(prec, gt, P, ft)
F, G = ft.to_arrays(), gt.to_arrays()
np.savez_compressed(os.path.join(HERE, "poisson_small.npz"), k=k, prec=prec, beta=beta, pos=np.array(pos), n_terms=P.size(),
                    f_scale=F["scale"], f_transl=F["transl"], f_norms=F["norms"], f_coefs_head=F["coefs"][:8], g_scale=G["scale"], g_transl=G["transl"],
                    g_coefs=G["coefs"], f_applied=st.fApplied, energy=orc.dot(gt, ft))
print("f nodes", len(F["scale"]), "g nodes", len(G["scale"]), "tuples", st.fApplied, "energy", orc.dot(gt, ft),
      "analytic", f.calc_coulomb_energy(f))
