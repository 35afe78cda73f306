"""Generates tests/golden/poisson_ref.npz with the REAL reference (MRCPP's own sources compiled in place, oracle/_ref, see
oracle/build_ref.sh): project + PoissonOperator + apply of one Gaussian, every node keyed by (scale, translation).
Run in the build container (where /root/reference exists): python tests/golden/make_golden_ref.py"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_api as ref  # noqa: E402


class G:  # the fields ref_api.project reads
    def __init__(self, beta, coef, pos):
        self.beta, self.coef, self.pos, self.power = beta, coef, pos, (0, 0, 0)


k, prec, beta, pos = 5, 1e-2, 10.0, (0.7, -0.4, 0.2)
rm = ref.MRA(k, -4, (-1, -1, -1), (2, 2, 2), 25)
f = G(beta, (beta / math.pi) ** 1.5, pos)
rf, rg = ref.Tree(rm), ref.Tree(rm)
ref.project(prec, rf, [f])
P = ref.poisson(rm, prec)
ref.apply(prec, rg, P, rf)
F, Gt = rf.export(), rg.export()
np.savez_compressed(os.path.join(HERE, "poisson_ref.npz"), k=k, prec=prec, beta=beta, pos=np.array(pos),
                    n_terms=ref.lib().ref_oper_n_terms(P), f_scale=F["scale"], f_transl=F["transl"], f_coefs=F["coefs"][::12],  # every 12th input node keeps the fixture small
                    f_coefs_rows=np.arange(len(F["scale"]))[::12],
                    f_square_norm=rf.square_norm(), g_scale=Gt["scale"], g_transl=Gt["transl"], g_branch=Gt["branch"],
                    g_coefs=Gt["coefs"], g_norms=Gt["norms"], g_square_norm=rg.square_norm(), energy=ref.dot(rg, rf))
print("reference: f nodes", len(F["scale"]), "g nodes", len(Gt["scale"]), "terms", ref.lib().ref_oper_n_terms(P), "energy", ref.dot(rg, rf))
