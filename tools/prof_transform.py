"""C2: mwTransform(TopDown) + mwTransform(BottomUp) sweep over a projected Gaussian tree, k = 5/7/9."""
import math, sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
L = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for k in (5, 7, 9):
    prec = 1e-6
    mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)
    rng = np.random.default_rng(42)
    func = mw.GaussExp()
    for i in range(n):
        beta = 10.0 ** rng.uniform(1, 3)
        func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-8, 8, 3))))
    f = mw.FunctionTree(mra)
    mw.project(prec, f, func)
    A = f.to_arrays()
    nb = int((A["child0"] >= 0).sum())
    K = k + 1
    f.sync_device()
    for kind, name in ((mw.TopDown, "TopDown"), (mw.BottomUp, "BottomUp")):
        f.mwTransform(kind)
        L.mrx_timer_start()
        for r in range(reps):
            f.mwTransform(kind)
        ms = L.mrx_timer_stop_ms() / reps
        gb = nb * 128 * K ** 3 / 1e9
        print(f"k={k} {name}: nodes {len(A['scale'])} branch {nb}  {ms:.3f} ms/pass  {nb/ms/1e3:.2f} Mnodes/s  {gb/ms*1e3:.0f} GB/s algorithmic  "
              f"{nb*96*K**4/ms/1e9:.2f} TFLOP/s", flush=True)
