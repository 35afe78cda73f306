"""End-to-end profile (one GPU): input in pinned host memory, output with a host mirror; MRX_PROFILE=1 prints the phases."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
k, prec = 7, 1e-7
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
P = mw.PoissonOperator(mra, prec)
rng = np.random.default_rng(42)
func = mw.GaussExp()
for i in range(n):
    beta = 10.0 ** rng.uniform(1, 3)
    func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
f = mw.FunctionTree(mra); mw.project(prec, f, func, device=True); f.sync_device()
for mirror in ((True, True) if os.environ.get('MRX_E2E_MIRROR_ONLY') else (True, False, True, False)):
    for rep in range(3):
        f.drop_device()
        g = mw.FunctionTree(mra)
        g.set_host_mirror(mirror)
        t = time.perf_counter()
        st = mw.apply(prec, g, P, f)
        t1 = time.perf_counter()
        g.sync_host()
        t2 = time.perf_counter()
        print(f"mirror {mirror} rep {rep}: apply {1e3*(t1-t):.1f} ms + sync_host {1e3*(t2-t1):.1f} ms = {1e3*(t2-t):.1f} ms; h2d {st.h2d_bytes/1e6:.0f} MB", flush=True)
        del g
if os.environ.get('MRX_E2E_KEEP'):  # partially resident input (nothing left to gather): the contraction kernel that checks arrival flags, alone
    for rep in range(2):
        g = mw.FunctionTree(mra); g.set_host_mirror(True)
        t = time.perf_counter(); st = mw.apply(prec, g, P, f); t1 = time.perf_counter()
        print(f"input partially resident rep {rep}: apply {1e3*(t1-t):.1f} ms; h2d {st.h2d_bytes/1e6:.0f} MB; contract {st.ms_contract:.2f} ms", flush=True)
        del g
