"""C2: filter kernels of mwTransform(TopDown, overwrite) / mwTransform(BottomUp) over a projected multi-centre
Gaussian tree, k = 3..11: level launches only, CUDA events (mrx_bench_mw_transform)."""
import ctypes as C, math, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
L = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
orders = [int(a) for a in sys.argv[3:]] or [3, 5, 7, 9, 11]
for k in orders:
    prec = 1e-6
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    rng = np.random.default_rng(42)
    func = mw.GaussExp()
    for i in range(n):
        beta = 10.0 ** rng.uniform(1, 3)
        func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
    f = mw.FunctionTree(mra)
    mw.project(prec, f, func, device=True)
    K = k + 1
    nb = C.c_int(0)
    for kind, name in ((0, "TopDown(overwrite)"), (1, "BottomUp")):
        L.mrx_bench_mw_transform(f._h, kind, 2, C.byref(nb))
        ms = L.mrx_bench_mw_transform(f._h, kind, reps, C.byref(nb))
        gb = nb.value * 128 * K ** 3 / 1e9
        real = gb * (4.5 if kind == 0 else 1.0)  # TopDown(overwrite) also zeroes 7 wavelet blocks of every child
        print(f"k={k} {name}: nodes {f.getNNodes()} parents {nb.value}  {ms:.3f} ms/pass  {nb.value/ms/1e3:.2f} Mnodes/s  "
              f"{gb/ms*1e3:.0f} GB/s algorithmic ({real/ms*1e3:.0f} GB/s moved)  {nb.value*96*K**4/ms/1e9:.2f} TFLOP/s", flush=True)
    del f
