"""Quick GPU-vs-oracle check used during development (the real tests are tests/test_gpu_*.py)."""
import math, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
import oracle_api as orc

def compare(a, b, name):
    A = a.to_arrays(); B = b.to_arrays()
    same = (A["scale"].shape == B["scale"].shape and np.array_equal(A["scale"], B["scale"]) and np.array_equal(A["transl"], B["transl"])
            and np.array_equal(A["child0"], B["child0"]))
    print(f"[{name}] nodes gpu={len(A['scale'])} oracle={len(B['scale'])} node-set identical={same}")
    if same:
        nrm = np.sqrt((B["coefs"] ** 2).sum(axis=1))
        err = np.abs(A["coefs"] - B["coefs"]).max(axis=1)
        rel = err / np.maximum(nrm, 1e-300)
        print(f"[{name}] max |dcoef|/||node|| = {rel.max():.3e}  (max abs {err.max():.3e})")
    return same

k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
prec = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
ncent = int(sys.argv[3]) if len(sys.argv) > 3 else 1
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
if ncent == 1:
    beta = 100.0
    func = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
else:
    rng = np.random.default_rng(42)
    func = mw.GaussExp()
    for i in range(ncent):
        beta = 10.0 ** rng.uniform(1, 3)
        func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / ncent, tuple(rng.uniform(-8, 8, 3))))
P = mw.PoissonOperator(mra, prec)
print("terms", P.size())
f_gpu = mw.FunctionTree(mra); t = time.time(); mw.project(prec, f_gpu, func); print("gpu project s", time.time() - t, "nodes", f_gpu.getNNodes())
f_cpu = mw.FunctionTree(mra); t = time.time(); orc.project(prec, f_cpu, func); print("cpu project s", time.time() - t)
compare(f_gpu, f_cpu, "project/BottomUp")
print("sqnorm gpu", f_gpu.getSquareNorm(), "cpu", f_cpu.getSquareNorm())
g_cpu = mw.FunctionTree(mra); t = time.time(); so = orc.apply(prec, g_cpu, P, f_cpu); print("cpu apply s", time.time() - t, so.as_dict())
for rep in range(3):
    g_gpu = mw.FunctionTree(mra); t = time.time(); sg = mw.apply(prec, g_gpu, P, f_gpu); print("gpu apply s", time.time() - t, sg.as_dict())
K = k + 1
print("GPU algorithmic TFLOP/s (kernel time):", sg.f_applied * 6 * K ** 4 / (sg.ms_kernel * 1e-3) / 1e12)
compare(g_gpu, g_cpu, "apply adaptive")
print("tuples gpu", sg.f_applied, "oracle", so.fApplied)
print("energy gpu", mw.dot(g_gpu, f_gpu), "cpu", orc.dot(g_cpu, f_cpu))
# fixed grid mode A
gA = mw.FunctionTree(mra); mw.copy_grid(gA, g_cpu); sA = mw.apply(prec, gA, P, f_gpu, maxIter=0)
gB = mw.FunctionTree(mra); mw.copy_grid(gB, g_cpu); sB = orc.apply(prec, gB, P, f_cpu, maxIter=0)
print("fixed grid tuples gpu", sA.f_applied, "oracle", sB.fApplied, "gpu ms_kernel", sA.ms_kernel,
      "TFLOP/s", sA.f_applied * 6 * K ** 4 / (sA.ms_kernel * 1e-3) / 1e12)
compare(gA, gB, "apply fixed grid")
# derivative
D = mw.ABGVOperator(mra, 0.5, 0.5)
for d in range(3):
    og = mw.FunctionTree(mra); sd = mw.apply(None, og, D, f_gpu, dir=d)
    oc = mw.FunctionTree(mra); sc = orc.apply_derivative(oc, D, f_cpu, d)
    print("deriv dir", d, "tuples", sd.f_applied, sc.fApplied)
    compare(og, oc, f"ABGV dir {d}")
