"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp). TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os

from mrcpp_b200 import _lib as _plib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")


class OrcStats(C.Structure):
    _fields_ = [("gNodes", C.c_longlong), ("fApplied", C.c_longlong), ("genUsed", C.c_longlong), ("iters", C.c_int),
                ("nNodesOut", C.c_int), ("t_band", C.c_double), ("t_calc", C.c_double), ("t_post", C.c_double),
                ("t_total", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_o = None


def lib():
    global _o
    if _o is None:
        if not os.path.exists(ORACLE_LIB):
            raise ImportError(f"{ORACLE_LIB} missing: run `make -C oracle`")
        o = C.CDLL(ORACLE_LIB)
        o.orc_set_table_path.argtypes = [C.c_char_p]
        o.orc_num_threads.restype = C.c_int
        o.orc_set_num_threads.argtypes = [C.c_int]
        o.orc_apply.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(OrcStats)]
        o.orc_apply_prec_trees.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                           C.POINTER(OrcStats)]
        o.orc_apply_unit_cell.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(OrcStats)]
        o.orc_apply_derivative.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OrcStats)]
        o.orc_mw_transform_down.argtypes = [C.c_void_p, C.c_int]
        o.orc_mw_transform_up.argtypes = [C.c_void_p]
        o.orc_calc_square_norm.argtypes = [C.c_void_p]
        o.orc_dot.argtypes = [C.c_void_p, C.c_void_p]
        o.orc_dot.restype = C.c_double
        o.orc_refine_grid.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        o.orc_refine_grid.restype = C.c_int
        o.orc_add_inplace.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
        o.orc_power.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int]
        o.orc_multiply.argtypes = [C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        o.orc_add.argtypes = [C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_void_p), C.c_int, C.c_int]
        o.orc_set_table_path(_plib.TABLES.encode())
        _o = o
    return _o


def _th(tree):
    return _plib.load().mrx_tree_host_handle(tree._h)


def _oh(oper):
    return _plib.load().mrx_oper_host_handle(oper._h)


def _modified(tree):
    _plib.load().mrx_tree_host_modified(tree._h)


def num_threads():
    return lib().orc_num_threads()


def apply(prec, out, oper, inp, maxIter=-1, absPrec=False):
    st = OrcStats()
    inp.sync_host()
    lib().orc_apply(prec, _th(out), _oh(oper), _th(inp), maxIter, 1 if absPrec else 0, C.byref(st))
    _modified(out)
    return st


def apply_unit_cell(inside, prec, out, oper, inp, maxIter=-1, absPrec=False):
    """apply_near_field (inside=True) / apply_far_field (inside=False) on a periodic world (src/treebuilders/apply.cpp:294-342)"""
    st = OrcStats()
    inp.sync_host()
    lib().orc_apply_unit_cell(1 if inside else 0, prec, _th(out), _oh(oper), _th(inp), maxIter, 1 if absPrec else 0, C.byref(st))
    _modified(out)
    return st


def apply_prec_trees(prec, out, oper, inp, prec_trees, maxIter=-1, absPrec=False):
    """apply(prec, out, oper, inp, precTrees, maxIter, absPrec) (src/treebuilders/apply.cpp:214-251): precision scaled per node by
    the largest norms of the precision trees. Oracle only: the device path does not build this variant yet."""
    st = OrcStats()
    inp.sync_host()
    for t in prec_trees:
        t.sync_host()
    h = (C.c_void_p * len(prec_trees))(*[_th(t) for t in prec_trees])
    lib().orc_apply_prec_trees(prec, _th(out), _oh(oper), _th(inp), len(prec_trees), h, maxIter, 1 if absPrec else 0, C.byref(st))
    _modified(out)
    return st


def apply_derivative(out, oper, inp, dir):
    st = OrcStats()
    inp.sync_host()
    lib().orc_apply_derivative(_th(out), _oh(oper), _th(inp), dir, C.byref(st))
    _modified(out)
    return st


def mw_transform_up(tree):
    tree.sync_host()
    lib().orc_mw_transform_up(_th(tree))
    _modified(tree)


def mw_transform_down(tree, overwrite=True):
    tree.sync_host()
    lib().orc_mw_transform_down(_th(tree), 1 if overwrite else 0)
    _modified(tree)


def calc_square_norm(tree):
    lib().orc_calc_square_norm(_th(tree))


def dot(bra, ket):
    bra.sync_host()
    ket.sync_host()
    return lib().orc_dot(_th(bra), _th(ket))


def add(out, coefs, trees, prec=-1.0, maxIter=-1, absPrec=False):
    """add(prec, out, {(c_i, tree_i)}, maxIter, absPrec) from the grid `out` enters with (src/treebuilders/add.cpp:41-70);
    prec < 0: no refinement"""
    for t in trees:
        t.sync_host()
    c = (C.c_double * len(trees))(*[float(x) for x in coefs])
    h = (C.c_void_p * len(trees))(*[_th(t) for t in trees])
    lib().orc_add(float(prec), _th(out), len(trees), c, h, int(maxIter), 1 if absPrec else 0)
    _modified(out)


def multiply(out, coefs, trees, prec=-1.0, maxIter=-1, absPrec=False, useMaxNorms=False):
    """multiply(prec, out, {(c_i, tree_i)}, maxIter, absPrec, useMaxNorms) (src/treebuilders/multiply.cpp:104-136)"""
    for t in trees:
        t.sync_host()
    c = (C.c_double * len(trees))(*[float(x) for x in coefs])
    h = (C.c_void_p * len(trees))(*[_th(t) for t in trees])
    lib().orc_multiply(float(prec), _th(out), len(trees), c, h, int(maxIter), 1 if absPrec else 0, 1 if useMaxNorms else 0)
    _modified(out)


def refine_grid(tree, prec=-1.0, absPrec=False, scales=0):
    """refine_grid(out, prec, absPrec) / refine_grid(out, scales) (src/treebuilders/grid.cpp:271-302); returns the new nodes"""
    tree.sync_host()
    n = lib().orc_refine_grid(_th(tree), float(prec), 1 if absPrec else 0, int(scales))
    _modified(tree)
    return n


def add_inplace(out, c, inp):
    """FunctionTree::add(c, inp) (src/trees/FunctionTree.cpp:687-706)"""
    out.sync_host()
    inp.sync_host()
    lib().orc_add_inplace(_th(out), float(c), _th(inp))
    _modified(out)


def power(out, inp, p, prec=-1.0, maxIter=-1, absPrec=False):
    """power(prec, out, inp, p) (src/treebuilders/multiply.cpp:211-234)"""
    inp.sync_host()
    lib().orc_power(float(prec), _th(out), _th(inp), float(p), int(maxIter), 1 if absPrec else 0)
    _modified(out)


def divergence(out, oper, trees):
    """divergence(out, D, {f_x, f_y, f_z}) (src/treebuilders/apply.cpp:514-530): derivative apply per direction, union grid, sum"""
    import mrcpp_b200 as mw
    parts = []
    for d, t in enumerate(trees):
        p = mw.FunctionTree(out.mra)
        apply_derivative(p, oper, t, d)
        parts.append(p)
    for p in parts:
        mw.build_grid(out, p)
    add(out, [1.0] * len(parts), parts)


def project(prec, out, func, build_grid=True):
    """product host projection (quadrature + in-node compression) closed by the ORACLE's BottomUp"""
    import mrcpp_b200 as mw
    mw.project(prec, out, func, build_grid=build_grid, finalize=False)
    mw_transform_up(out)
    calc_square_norm(out)
