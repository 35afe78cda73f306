"""Tree algebra around the apply on bench-like trees: add (fixed grid, adaptive), multiply, in-place add,
refine_grid, gradient / divergence, timed on the library stream (mrx_timer_*), with the algorithmic HBM traffic of the dominant
kernels:  add on a grid      : per shared node read 7 K^3 + write 7 K^3 doubles (axpy_nodes_kernel) + one TopDown(+=) pass
                                (128 K^3 B per parent node)
          multiply           : per end node and input 9 slots x 8 K^3 doubles of scratch traffic around product_values_kernel
                                (8 K^3 values read + written)
usage: python tools/prof_algebra.py [centres=100] [k=7] [prec=1e-6]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mrcpp_b200 as mw
from mrcpp_b200 import _lib
_lib.init()
L = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
k = int(sys.argv[2]) if len(sys.argv) > 2 else 7
prec = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
K = k + 1
mra = mw.MultiResolutionAnalysis(k, -4, (-1, -1, -1), (2, 2, 2), 25)


def density(seed, box):
    rng = np.random.default_rng(seed)
    g = mw.GaussExp()
    for _ in range(n):
        beta = 10.0 ** rng.uniform(1, 3)
        g.append(mw.GaussFunc(beta, (beta / math.pi) ** 1.5 / n, tuple(rng.uniform(-box, box, 3))))
    return g


def timed(name, fn, nodes_fn, bytes_per_node):
    fn()  # warm-up (allocator, caches)
    L.mrx_timer_start()
    out = fn()
    ms = L.mrx_timer_stop_ms()
    nodes = nodes_fn(out)
    print(f"{name:34s} {ms:9.3f} ms  nodes {nodes:8d}  {nodes / ms / 1e3:8.2f} Mnodes/s  {nodes * bytes_per_node / ms / 1e6:8.0f} GB/s algorithmic",
          flush=True)
    return out


a, b = mw.FunctionTree(mra), mw.FunctionTree(mra)
mw.project(prec, a, density(42, 8.0), device=True)
mw.project(prec, b, density(43, 8.0), device=True)
print(f"k={k} prec={prec:g} centres={n}: trees of {a.getNNodes()} and {b.getNNodes()} nodes ({a.nbytes() / 1e9:.2f} / {b.nbytes() / 1e9:.2f} GB)")
node_bytes = 8 * K ** 3 * 8


def add_fixed():
    o = mw.FunctionTree(mra)
    mw.build_grid(o, a)
    mw.build_grid(o, b)
    mw.add(-1.0, o, [(1.0, a), (-2.0, b)], 0)
    return o


def add_adaptive():
    o = mw.FunctionTree(mra)
    mw.add(prec, o, [(1.0, a), (-2.0, b)])
    return o


def multiply():
    o = mw.FunctionTree(mra)
    mw.multiply(prec, o, [(1.0, a), (1.0, b)])
    return o


def square():
    o = mw.FunctionTree(mra)
    mw.multiply(prec, o, [(1.0, a), (1.0, a)])
    return o


def inplace():
    o = mw.FunctionTree(mra)
    mw.copy_grid(o, a)
    mw.add(-1.0, o, [(1.0, a)], 0)
    o.add(0.5, b)
    return o


def divergence():
    D = mw.ABGVOperator(mra, 0.5, 0.5)
    o = mw.FunctionTree(mra)
    mw.divergence(o, D, [(1.0, a), (1.0, b), (1.0, a)])
    return o


nn = lambda t: t.getNNodes()
timed("add on the union grid", add_fixed, nn, 2 * node_bytes * 2 + 2 * node_bytes)   # two inputs read + out r/w, + TopDown r/w
timed("add, adaptive", add_adaptive, nn, 2 * node_bytes * 2 + 2 * node_bytes)
timed("multiply, adaptive", multiply, nn, 2 * 9 * node_bytes * 2)                      # two inputs x 9 scratch slots r/w
timed("square, adaptive", square, nn, 2 * 9 * node_bytes * 2)
timed("copy + in-place add", inplace, nn, 4 * node_bytes)
timed("divergence (3 derivative applies)", divergence, nn, 6 * node_bytes)
