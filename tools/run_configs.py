"""BASELINE.json configs C1..C5 on one GPU: sizes, times, algorithmic TFLOP/s, analytic checks.
usage: python tools/run_configs.py [c1 c3 c4 c5]   (C2 = tools/prof_transform.py)"""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
which = [a.lower() for a in sys.argv[1:]] or ["c1", "c3", "c4", "c5"]


def world(k):
    return mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)


def timed_apply(prec, mra, oper, f, reps=3, **kw):
    best, st = 1e30, None
    for _ in range(reps):
        g = mw.FunctionTree(mra)
        t = time.perf_counter()
        st = mw.apply(prec, g, oper, f, **kw)
        best = min(best, time.perf_counter() - t)
    return g, st, best


def report(name, k, st, dt, extra=""):
    K = k + 1
    print(f"{name}: output nodes {st.g_nodes} final {st.n_nodes_out} tuples {st.f_applied} iterations {st.iterations} "
          f"apply {dt*1e3:.1f} ms ({st.g_nodes/dt/1e3:.1f} K nodes/s) contraction {st.ms_contract:.1f} ms = "
          f"{st.f_applied*6*K**4/max(st.ms_contract,1e-9)/1e9:.2f} TFLOP/s, whole apply {st.f_applied*6*K**4/dt/1e12:.2f} TFLOP/s {extra}", flush=True)


if "c1" in which:  # examples/poisson.cpp: single Gaussian, k=7, prec 1e-5 (and the 1e-7 target variant)
    for prec in (1e-5, 1e-7):
        mra = world(7)
        beta = 100.0
        func = mw.GaussFunc(beta, (beta / math.pi) ** 1.5, (math.pi / 3,) * 3)
        P = mw.PoissonOperator(mra, prec)
        f = mw.FunctionTree(mra); mw.project(prec, f, func, device=True)
        g, st, dt = timed_apply(prec, mra, P, f)
        en = mw.dot(g, f)
        report(f"C1 poisson.cpp k=7 prec={prec:g} M={P.size()}", 7, st, dt, f"| energy {en:.10f} vs sqrt(2 beta/pi) = 7.9788456080 (rel {abs(en-7.978845608)/7.978845608:.1e})")

if "c3" in which:  # ABGV derivative, k=7, 10 normalised Gaussians, seed 1234, prec 1e-7
    mra = world(7)
    rng = np.random.default_rng(1234)
    func = mw.GaussExp()
    for _ in range(10):
        beta = 10.0 ** rng.uniform(0, 2)
        func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / 10, tuple(rng.uniform(-4, 4, 3))))
    f = mw.FunctionTree(mra); mw.project(1e-7, f, func, device=True)
    for (a, b) in ((0.5, 0.5), (0.0, 0.0)):
        D = mw.ABGVOperator(mra, a, b)
        for d in range(3):
            best = 1e30
            for _ in range(3):
                g = mw.FunctionTree(mra)
                t = time.perf_counter(); st = mw.apply(None, g, D, f, dir=d); best = min(best, time.perf_counter() - t)
            print(f"C3 ABGV({a},{b}) dir {d}: input nodes {f.getNNodes()} output nodes {st.n_nodes_out} tuples {st.f_applied} "
                  f"apply {best*1e3:.1f} ms ({st.n_nodes_out/best/1e6:.2f} M nodes/s)", flush=True)

if "c4" in which:  # Helmholtz k=9 prec 1e-7 mu=1, orbital-like inputs on 12 benzene-like centres (first 4 of the 50 trees)
    k, prec = 9, 1e-7
    mra = world(k)
    H = mw.HelmholtzOperator(mra, 1.0, prec)
    centres = [(2.64 * math.cos(i * math.pi / 3), 2.64 * math.sin(i * math.pi / 3), 0.0, 1.5) for i in range(6)] + \
              [(4.69 * math.cos(i * math.pi / 3), 4.69 * math.sin(i * math.pi / 3), 0.0, 0.8) for i in range(6)]
    tot_t, tot_nodes, tot_tuples, tot_c = 0.0, 0, 0, 0.0
    for j in range(4):
        rng = np.random.default_rng(2024 + j)
        func = mw.GaussExp()
        for (x, y, z, beta) in centres:
            func.append(mw.GaussFunc(beta, float(rng.normal()), (x, y, z)))
        f = mw.FunctionTree(mra); mw.project(prec, f, func, device=True)
        g, st, dt = timed_apply(prec, mra, H, f, reps=2)
        g.rescale(-1.0 / (2.0 * math.pi))
        tot_t += dt; tot_nodes += st.g_nodes; tot_tuples += st.f_applied; tot_c += st.ms_contract
        report(f"C4 Helmholtz k=9 prec=1e-7 M={H.size()} orbital {j} (input nodes {f.getNNodes()})", k, st, dt)
    K = k + 1
    print(f"C4 total (4 of 50 trees): {tot_nodes/tot_t/1e3:.1f} K nodes/s, contraction {tot_tuples*6*K**4/tot_c/1e9:.2f} TFLOP/s, whole {tot_tuples*6*K**4/tot_t/1e12:.2f} TFLOP/s")

if "c5" in which:  # Poisson k=11 prec 1e-9, 100 normalised Gaussians / 100, seed 42
    k, prec = 11, 1e-9
    mra = world(k)
    rng = np.random.default_rng(42)
    func = mw.GaussExp()
    for _ in range(100):
        beta = 10.0 ** rng.uniform(1, 3)
        func.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / 100, tuple(rng.uniform(-8, 8, 3))))
    t = time.perf_counter()
    P = mw.PoissonOperator(mra, prec)
    t_op = time.perf_counter() - t
    t = time.perf_counter()
    f = mw.FunctionTree(mra); mw.project(prec, f, func, device=True)
    t_pr = time.perf_counter() - t
    print(f"C5 setup: operator M={P.size()} {t_op:.2f} s, projection {t_pr:.2f} s, input nodes {f.getNNodes()} ({f.nbytes()/1e9:.1f} GB)", flush=True)
    g, st, dt = timed_apply(prec, mra, P, f, reps=2)
    ana = sum(a.calc_coulomb_energy(b) for a in func for b in func)
    en = mw.dot(g, f)
    report("C5 Poisson k=11 prec=1e-9", k, st, dt, f"| Coulomb energy {en:.12f} vs analytic {ana:.12f} (rel {abs(en-ana)/ana:.1e})")
