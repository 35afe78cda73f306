import math, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
k = int(sys.argv[1]); prec = float(sys.argv[2]); n = int(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
rng = np.random.default_rng(42)
func = mw.GaussExp()
for i in range(n):
    beta = 10.0 ** rng.uniform(1, 3)
    func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
P = mw.PoissonOperator(mra, prec)
f = mw.FunctionTree(mra); t = time.time(); mw.project(prec, f, func, device=True); print("project s", time.time() - t, "nodes", f.getNNodes(), flush=True)
for r in range(reps):
    g = mw.FunctionTree(mra); t = time.time(); st = mw.apply(prec, g, P, f); dt = time.time() - t
    K = k + 1
    print(f"apply {dt*1e3:.1f} ms  kernel {st.ms_kernel:.1f} ms contract {st.ms_contract:.1f} ms  build {st.ms_build:.1f} post {st.ms_post:.1f} nodes {st.g_nodes} tuples {st.f_applied} gen {st.gen_nodes} "
          f"TF/s contract {st.f_applied*6*K**4/st.ms_contract/1e9:.2f} total {st.f_applied*6*K**4/dt/1e12:.2f}", flush=True)
A = f.to_arrays(coefs=False); B = g.to_arrays(coefs=False)
print("input nodes by depth ", np.bincount(A["scale"] + 4).tolist())
print("output nodes by depth", np.bincount(B["scale"] + 4).tolist())
