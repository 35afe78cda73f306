"""ctypes wrapper of the REAL reference compiled in place (oracle/_ref/libref_driver.so, built by oracle/build_ref.sh from
/root/reference with the Eigen stand-in of oracle/eigen_shim). TEST INFRASTRUCTURE: imported only by tests/ and by bench.py's
reference arm. Absent library -> available() is False and the tests that need it skip."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_driver.so")
_l = None


FILTERS = os.path.join(ROOT, "oracle", "_ref", "mwfilters")  # unpacked from mrcpp_b200/data/mwtables.bin by build_ref.sh


def available():
    return os.path.exists(LIB) and os.path.isdir(FILTERS)


def lib():
    global _l
    if _l is None:
        os.environ["MWFILTERS_DIR"] = FILTERS  # details::find_filters (src/utils/details.cpp:53-69) looks here first
        l = C.CDLL(LIB)
        P, D, I = C.c_void_p, C.c_double, C.c_int
        PD, PI = C.POINTER(C.c_double), C.POINTER(C.c_int)
        sig = {
            "ref_mra_create": (P, [I, I, PI, PI, I]), "ref_mra_destroy": (None, [P]),
            "ref_tree_create": (P, [P]), "ref_tree_destroy": (None, [P]), "ref_tree_n_nodes": (I, [P]),
            "ref_tree_square_norm": (D, [P]), "ref_project_gaussians": (None, [P, D, I, PD, PD, PD, PI, I]),
            "ref_poisson_create": (P, [P, D]), "ref_helmholtz_create": (P, [P, D, D]), "ref_abgv_create": (P, [P, D, D]),
            "ref_oper_n_terms": (I, [P]), "ref_conv_destroy": (None, [P]), "ref_deriv_destroy": (None, [P]),
            "ref_apply": (D, [D, P, P, P, I, I]), "ref_apply_derivative": (None, [P, P, P, I]), "ref_dot": (D, [P, P]),
            "ref_copy_grid": (None, [P, P]), "ref_mw_transform": (None, [P, I, I]),
            "ref_tree_export": (I, [P, PI, PI, PI, PD, PD]), "ref_num_threads": (I, []),
            "ref_ph_create": (P, [P, I]), "ref_bs_create": (P, [P, I]), "ref_tree_integrate": (D, [P]), "ref_tree_save_txt": (None, [P, C.c_char_p]), "ref_tree_load_txt": (None, [P, C.c_char_p]), "ref_tree_evalf": (D, [P, PD, I]), "ref_build_grid_tree": (None, [P, P]), "ref_add": (None, [P, I, PD, C.POINTER(P)]),
            "ref_divergence": (None, [P, P, C.POINTER(P)]), "ref_add_adaptive": (None, [D, P, I, PD, C.POINTER(P), I, I]), "ref_multiply": (None, [D, P, I, PD, C.POINTER(P), I, I, I]), "ref_refine_grid": (I, [P, D, I, I]), "ref_power": (None, [D, P, P, D, I, I]), "ref_apply_prec_trees": (D, [D, P, P, P, I, C.POINTER(P), I, I]),
            "ref_mra_create_periodic": (P, [I, I]), "ref_project_cosines": (None, [P, D, I, PD, PD]),
            "ref_poisson_create_reach": (P, [P, D, I, I]), "ref_helmholtz_create_reach": (P, [P, D, D, I, I]),
            "ref_apply_unit_cell": (None, [I, D, P, P, P, I, I]),
            "ref_add_inplace": (None, [P, D, P]), "ref_clear_grid": (None, [P]), "ref_build_grid_gaussians": (None, [P, I, PD, PD, PD, PI]),
        }
        for name, (res, args) in sig.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _l = l
    return _l


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class MRA:
    def __init__(self, order, root_scale, corner, nboxes, max_depth):
        c = np.ascontiguousarray(corner, dtype=np.int32)
        b = np.ascontiguousarray(nboxes, dtype=np.int32)
        self.k = order
        self._h = lib().ref_mra_create(order, root_scale, _ip(c), _ip(b), max_depth)


class PeriodicMRA(MRA):
    """unit cell [-1, 1]^3 with periodic boundary conditions (BoundingBox(0, -1, 2, sf = 1, pbc = true))"""

    def __init__(self, order, max_depth=25):
        self.k = order
        self._h = lib().ref_mra_create_periodic(order, max_depth)


class Tree:
    def __init__(self, mra):
        self.mra = mra
        self._h = lib().ref_tree_create(mra._h)

    def n_nodes(self):
        return lib().ref_tree_n_nodes(self._h)

    def square_norm(self):
        return lib().ref_tree_square_norm(self._h)

    def integrate(self):
        return lib().ref_tree_integrate(self._h)

    def save_txt(self, path):
        lib().ref_tree_save_txt(self._h, str(path).encode())

    def load_txt(self, path):
        lib().ref_tree_load_txt(self._h, str(path).encode())

    def evalf(self, r, precise=False):
        x = np.ascontiguousarray(r, dtype=np.float64)
        return lib().ref_tree_evalf(self._h, _dp(x), 1 if precise else 0)

    def export(self):
        """dict keyed like FunctionTree.to_arrays(), nodes in the reference's node-table order"""
        n = self.n_nodes()
        K = self.mra.k + 1
        out = {"scale": np.zeros(n, np.int32), "transl": np.zeros((n, 3), np.int32), "branch": np.zeros(n, np.int32),
               "coefs": np.zeros((n, 8 * K ** 3)), "norms": np.zeros((n, 8))}
        m = lib().ref_tree_export(self._h, _ip(out["scale"]), _ip(out["transl"]), _ip(out["branch"]), _dp(out["coefs"]), _dp(out["norms"]))
        assert m == n
        return out


def project(prec, tree, gauss_list, build_grid=True):
    coef = np.array([g.coef for g in gauss_list], dtype=np.float64)
    expo = np.array([g.beta for g in gauss_list], dtype=np.float64)
    pos = np.ascontiguousarray([g.pos for g in gauss_list], dtype=np.float64)
    power = np.ascontiguousarray([g.power for g in gauss_list], dtype=np.int32)
    lib().ref_project_gaussians(tree._h, float(prec), len(coef), _dp(coef), _dp(expo), _dp(pos), _ip(power), 1 if build_grid else 0)


def build_grid(tree, gauss_list):
    coef = np.array([g.coef for g in gauss_list], dtype=np.float64)
    expo = np.array([g.beta for g in gauss_list], dtype=np.float64)
    pos = np.ascontiguousarray([g.pos for g in gauss_list], dtype=np.float64)
    power = np.ascontiguousarray([g.power for g in gauss_list], dtype=np.int32)
    lib().ref_build_grid_gaussians(tree._h, len(coef), _dp(coef), _dp(expo), _dp(pos), _ip(power))


def poisson(mra, prec):
    return lib().ref_poisson_create(mra._h, float(prec))


def helmholtz(mra, mu, prec):
    return lib().ref_helmholtz_create(mra._h, float(mu), float(prec))


def abgv(mra, a, b):
    return lib().ref_abgv_create(mra._h, float(a), float(b))


def ph(mra, order):
    return lib().ref_ph_create(mra._h, int(order))


def bs(mra, order):
    return lib().ref_bs_create(mra._h, int(order))


def apply(prec, out, oper, inp, maxIter=-1, absPrec=False):
    return lib().ref_apply(float(prec), out._h, oper, inp._h, int(maxIter), 1 if absPrec else 0)


def project_cosines(prec, tree, amp, kvec):
    a = np.ascontiguousarray(amp, dtype=np.float64)
    k = np.ascontiguousarray(kvec, dtype=np.float64).reshape(len(a), 3)
    lib().ref_project_cosines(tree._h, float(prec), len(a), _dp(a), _dp(k))


def poisson_reach(mra, prec, root, reach):
    return lib().ref_poisson_create_reach(mra._h, float(prec), int(root), int(reach))


def helmholtz_reach(mra, mu, prec, root, reach):
    return lib().ref_helmholtz_create_reach(mra._h, float(mu), float(prec), int(root), int(reach))


def apply_unit_cell(inside, prec, out, oper, inp, maxIter=-1, absPrec=False):
    lib().ref_apply_unit_cell(1 if inside else 0, float(prec), out._h, oper, inp._h, int(maxIter), 1 if absPrec else 0)


def apply_prec_trees(prec, out, oper, inp, prec_trees, maxIter=-1, absPrec=False):
    h = (C.c_void_p * len(prec_trees))(*[t._h for t in prec_trees])
    return lib().ref_apply_prec_trees(float(prec), out._h, oper, inp._h, len(prec_trees), h, int(maxIter), 1 if absPrec else 0)


def apply_derivative(out, oper, inp, d):
    lib().ref_apply_derivative(out._h, oper, inp._h, int(d))


def dot(a, b):
    return lib().ref_dot(a._h, b._h)


def build_grid_tree(out, inp):
    lib().ref_build_grid_tree(out._h, inp._h)


def add(out, coefs, trees, prec=None, maxIter=-1, absPrec=False):
    c = np.ascontiguousarray(coefs, dtype=np.float64)
    h = (C.c_void_p * len(trees))(*[t._h for t in trees])
    if prec is None:
        lib().ref_add(out._h, len(trees), _dp(c), h)
    else:
        lib().ref_add_adaptive(float(prec), out._h, len(trees), _dp(c), h, int(maxIter), 1 if absPrec else 0)


def multiply(out, coefs, trees, prec=-1.0, maxIter=-1, absPrec=False, useMaxNorms=False):
    c = np.ascontiguousarray(coefs, dtype=np.float64)
    h = (C.c_void_p * len(trees))(*[t._h for t in trees])
    lib().ref_multiply(float(prec), out._h, len(trees), _dp(c), h, int(maxIter), 1 if absPrec else 0, 1 if useMaxNorms else 0)


def refine_grid(tree, prec=-1.0, absPrec=False, scales=0):
    return lib().ref_refine_grid(tree._h, float(prec), 1 if absPrec else 0, int(scales))


def add_inplace(out, c, inp):
    lib().ref_add_inplace(out._h, float(c), inp._h)


def power(out, inp, p, prec=-1.0, maxIter=-1, absPrec=False):
    lib().ref_power(float(prec), out._h, inp._h, float(p), int(maxIter), 1 if absPrec else 0)


def divergence(out, oper, trees):
    h = (C.c_void_p * 3)(*[t._h for t in trees])
    lib().ref_divergence(out._h, oper, h)


def by_index(arrays):
    """{(scale, lx, ly, lz): row} of an exported / to_arrays() tree"""
    return {(int(s), int(l[0]), int(l[1]), int(l[2])): i for i, (s, l) in enumerate(zip(arrays["scale"], arrays["transl"]))}
