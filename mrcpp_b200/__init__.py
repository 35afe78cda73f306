"""mrcpp_b200 — B200-native operator application for adaptive multiwavelet function trees.

Python host mirror of the part of the MRCPP C++ API that sits on the hot path (names and argument
meaning follow the reference: api/MRCPP/MWFunctions, MWOperators). Everything here is a thin wrapper
over the C-ABI in include/mrcpp_b200.h; the numerics run in hand-written sm_100a kernels.

    mra  = MultiResolutionAnalysis(order=7, root_scale=-4, corner=(-1,-1,-1), boxes=(2,2,2), max_depth=25)
    f    = FunctionTree(mra); project(prec, f, GaussFunc(beta, alpha, pos))
    P    = PoissonOperator(mra, prec)
    g    = FunctionTree(mra); apply(prec, g, P, f)
    e    = dot(g, f)
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import ApplyStats

TopDown = 0   # api/constants.h
BottomUp = 1
Forward, Backward = 0, 1             # CV_Transform
Compression, Reconstruction = 0, 1   # MW_Transform


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class MultiResolutionAnalysis:
    """BoundingBox<3> + InterpolatingBasis + max depth (src/trees/MultiResolutionAnalysis.h:49)."""

    def __init__(self, order, root_scale=0, corner=(0, 0, 0), boxes=(1, 1, 1), max_depth=30, periodic=False):
        _lib.init()
        self.order, self.root_scale, self.corner, self.boxes, self.max_depth = order, root_scale, tuple(corner), tuple(boxes), max_depth
        self.periodic = bool(periodic)
        c = np.asarray(corner, dtype=np.int32)
        b = np.asarray(boxes, dtype=np.int32)
        self._h = _lib.load().mrx_mra_create(order, root_scale, _ip(c), _ip(b), max_depth)
        if periodic:  # BoundingBox(..., pbc=True): unit cell [-1, 1]^3 (root scale 0, corner -1, 2 boxes per dimension)
            _lib.load().mrx_mra_set_periodic(self._h, 1)

    @property
    def kp1(self):
        return self.order + 1

    def lower_bounds(self):
        return tuple(2.0 ** (-self.root_scale) * c for c in self.corner)

    def upper_bounds(self):
        return tuple(2.0 ** (-self.root_scale) * (c + b) for c, b in zip(self.corner, self.boxes))

    def __del__(self):
        try:
            _lib.load().mrx_mra_destroy(self._h)
        except Exception:
            pass


class GaussFunc:
    """coef * prod (x-pos)^pow * exp(-beta |x-pos|^2) (src/functions/GaussFunc.h)."""

    def __init__(self, beta, alpha=1.0, pos=(0.0, 0.0, 0.0), power=(0, 0, 0)):
        self.beta, self.coef, self.pos, self.power = float(beta), float(alpha), tuple(pos), tuple(power)

    def calc_coulomb_energy(self, other):
        """GaussFunc::calcCoulombEnergy (src/functions/GaussFunc.cpp:210-237) for s-type Gaussians."""
        import math
        p, q = self.beta, other.beta
        alpha = p * q / (p + q)
        R2 = sum((a - b) ** 2 for a, b in zip(self.pos, other.pos))
        x = alpha * R2
        boys = 1.0 if x < 1e-14 else 0.5 * math.sqrt(math.pi / x) * math.erf(math.sqrt(x))
        return math.sqrt(4.0 * alpha / math.pi) * boys * self.coef * other.coef * (math.pi / p) ** 1.5 * (math.pi / q) ** 1.5

    def evalf(self, r):
        r = np.asarray(r, dtype=float)
        q = r - np.asarray(self.pos)
        poly = np.prod(q ** np.asarray(self.power), axis=-1)
        return self.coef * poly * np.exp(-self.beta * np.sum(q * q, axis=-1))


class GaussExp(list):
    """Sum of GaussFunc (src/functions/GaussExp.h)."""

    def evalf(self, r):
        return sum(g.evalf(r) for g in self)


class FunctionTree:
    """FunctionTree<3,double> with an HBM-resident node store (src/trees/FunctionTree.h)."""

    def __init__(self, mra, _handle=None):
        self.mra = mra
        self._h = _handle if _handle is not None else _lib.load().mrx_tree_create(mra._h)

    def __del__(self):
        try:
            _lib.load().mrx_tree_destroy(self._h)
        except Exception:
            pass

    # -- reference-named accessors
    def getNNodes(self):
        return _lib.load().mrx_tree_n_nodes(self._h)

    def getNEndNodes(self):
        return _lib.load().mrx_tree_n_end_nodes(self._h)

    def getSquareNorm(self):
        return _lib.load().mrx_tree_square_norm(self._h)

    def evalf(self, r, precise=False):
        """FunctionTree::evalf / evalf_precise (src/trees/FunctionTree.cpp:374-436) at one point or an (n, 3) array of points"""
        pts = np.ascontiguousarray(np.atleast_2d(np.asarray(r, dtype=np.float64)))
        out = np.zeros(len(pts))
        _lib.load().mrx_tree_evalf(self._h, len(pts), _dp(pts), _dp(out), 1 if precise else 0)
        return float(out[0]) if np.ndim(r) == 1 else out

    def add(self, c, inp):
        """FunctionTree::add(c, inp): in place, on this tree's grid (src/trees/FunctionTree.cpp:687-706)"""
        _lib.load().mrx_tree_add_inplace(self._h, float(c), inp._h)

    def saveTreeTXT(self, path):
        """FunctionTree::saveTreeTXT (src/trees/FunctionTree.cpp:306-372)"""
        _lib.load().mrx_tree_save_txt(self._h, str(path).encode())

    def loadTreeTXT(self, path):
        """FunctionTree::loadTreeTXT (src/trees/FunctionTree.cpp:240-305), files written by saveTreeTXT"""
        _lib.load().mrx_tree_load_txt(self._h, str(path).encode())

    def integrate(self):
        """FunctionTree::integrate (src/trees/FunctionTree.cpp:438-454)"""
        return _lib.load().mrx_tree_integrate(self._h)

    def clear(self):
        _lib.load().mrx_tree_clear(self._h)

    def mwTransform(self, kind, overwrite=True):
        _lib.load().mrx_mw_transform(self._h, kind, 1 if overwrite else 0)

    def nodeMwTransform(self, kind, slots=None):
        """MWNode::mwTransform(kind) (src/trees/MWNode.cpp:557-594) of the listed nodes (None: every node), in place"""
        a = None if slots is None else np.ascontiguousarray(slots, dtype=np.int32)
        _lib.load().mrx_node_mw_transform(self._h, int(kind), -1 if a is None else len(a), None if a is None else _ip(a))

    def nodeCvTransform(self, kind, slots=None):
        """MWNode::cvTransform(kind) (src/trees/MWNode.cpp:448-490) of the listed nodes (None: every node), in place"""
        a = None if slots is None else np.ascontiguousarray(slots, dtype=np.int32)
        _lib.load().mrx_node_cv_transform(self._h, int(kind), -1 if a is None else len(a), None if a is None else _ip(a))

    def calcSquareNorm(self):
        return _lib.load().mrx_calc_square_norm(self._h)

    def rescale(self, c):
        _lib.load().mrx_tree_rescale(self._h, float(c))

    # -- array interface
    @classmethod
    def from_arrays(cls, mra, scale, transl, parent, child0, coefs):
        scale = np.ascontiguousarray(scale, dtype=np.int32)
        transl = np.ascontiguousarray(transl, dtype=np.int32)
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        child0 = np.ascontiguousarray(child0, dtype=np.int32)
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        h = _lib.load().mrx_tree_from_arrays(mra._h, len(scale), _ip(scale), _ip(transl), _ip(parent), _ip(child0), _dp(coefs))
        return cls(mra, _handle=h)

    def to_arrays(self, coefs=True):
        n = self.getNNodes()
        K = self.mra.kp1
        out = {
            "scale": np.zeros(n, dtype=np.int32),
            "transl": np.zeros((n, 3), dtype=np.int32),
            "parent": np.zeros(n, dtype=np.int32),
            "child0": np.zeros(n, dtype=np.int32),
            "norms": np.zeros((n, 8), dtype=np.float64),
        }
        cp = None
        if coefs:
            out["coefs"] = np.zeros((n, 8 * K ** 3), dtype=np.float64)
            cp = _dp(out["coefs"])
        _lib.load().mrx_tree_to_arrays(self._h, _ip(out["scale"]), _ip(out["transl"]), _ip(out["parent"]), _ip(out["child0"]),
                                       cp, _dp(out["norms"]))
        return out

    def sync_device(self):
        _lib.load().mrx_tree_sync_device(self._h)

    def sync_host(self):
        _lib.load().mrx_tree_sync_host(self._h)

    def set_host_mirror(self, on=True, comm=None):
        """keep the host copy of this tree current when an apply writes it (result streamed down while the apply runs).
        With `comm` (every rank calls it; `comm.host_arena(nbytes)` first): the host chunks move into the ranks' shared host
        arena and every rank downloads its share over its own PCIe link. Returns True when the shared path is in use."""
        if comm is not None and on:
            self._mirror_comm = comm  # the arena must outlive the tree
            return _lib.load().mrx_tree_set_shared_host_mirror(self._h, comm._h) == 0
        _lib.load().mrx_tree_set_host_mirror(self._h, 1 if on else 0)
        return False

    def drop_device(self):
        _lib.load().mrx_tree_drop_device(self._h)

    def nbytes(self):
        return _lib.load().mrx_tree_bytes(self._h)


class _Operator:
    def __init__(self, mra, handle):
        self.mra = mra
        self._h = handle

    def __del__(self):
        try:
            _lib.load().mrx_oper_destroy(self._h)
        except Exception:
            pass

    def size(self):
        return _lib.load().mrx_oper_n_terms(self._h)

    def band_widths(self, prec):
        """getMaxBandWidth(depth) for every depth after calcBandWidths(prec) (MWOperator.cpp:63-108)."""
        buf = np.zeros(64, dtype=np.int32)
        n = _lib.load().mrx_oper_band_widths(self._h, float(prec), _ip(buf), 64)
        return buf[:n].copy()

    def node(self, term, depth, transl):
        K = self.mra.kp1
        mats = np.zeros((4, K * K))
        norms = np.zeros(4)
        rc = _lib.load().mrx_oper_node(self._h, term, depth, transl, _dp(mats), _dp(norms))
        if rc != 0:
            raise IndexError((term, depth, transl))
        return mats, norms


def apply_near_field(prec, out, oper, inp, maxIter=-1, absPrec=False):
    """mrcpp::apply_near_field (src/treebuilders/apply.cpp:318-342): periodic world, contributions from inside the unit cell only"""
    st = ApplyStats()
    _lib.load().mrx_apply_unit_cell(1, float(prec), out._h, oper._h, inp._h, int(maxIter), 1 if absPrec else 0, C.byref(st))
    return st


def apply_far_field(prec, out, oper, inp, maxIter=-1, absPrec=False):
    """mrcpp::apply_far_field (src/treebuilders/apply.cpp:272-316): periodic world, contributions from outside the unit cell only"""
    st = ApplyStats()
    _lib.load().mrx_apply_unit_cell(0, float(prec), out._h, oper._h, inp._h, int(maxIter), 1 if absPrec else 0, C.byref(st))
    return st


def project_cosines(prec, out, amp, kvec, finalize=True):
    """project(prec, out, f) of f(r) = sum_i amp[i] prod_d cos(pi kvec[i][d] r_d) (host quadrature, native callback)"""
    a = np.ascontiguousarray(amp, dtype=np.float64)
    k = np.ascontiguousarray(kvec, dtype=np.float64).reshape(len(a), 3)
    _lib.load().mrx_project_cosines(out._h, float(prec), len(a), _dp(a), _dp(k), 1 if finalize else 0)


class PoissonOperator(_Operator):
    """src/operators/PoissonOperator.cpp:40-55"""

    def __init__(self, mra, prec, root=None, reach=None):
        """PoissonOperator(mra, prec) or, for periodic worlds, PoissonOperator(mra, prec, root, reach) (PoissonOperator.cpp:56-77)"""
        _lib.init()
        if root is None:
            super().__init__(mra, _lib.load().mrx_poisson_create(mra._h, float(prec)))
        else:
            super().__init__(mra, _lib.load().mrx_poisson_create_reach(mra._h, float(prec), int(root), int(reach)))


class HelmholtzOperator(_Operator):
    """src/operators/HelmholtzOperator.cpp:44-59"""

    def __init__(self, mra, mu, prec, root=None, reach=None):
        """HelmholtzOperator(mra, mu, prec) or, for periodic worlds, (mra, mu, prec, root, reach) (HelmholtzOperator.cpp:60-81)"""
        _lib.init()
        if root is None:
            super().__init__(mra, _lib.load().mrx_helmholtz_create(mra._h, float(mu), float(prec)))
        else:
            super().__init__(mra, _lib.load().mrx_helmholtz_create_reach(mra._h, float(mu), float(prec), int(root), int(reach)))


class ConvolutionOperator(_Operator):
    """ConvolutionOperator<3>(mra, GaussExp<1>, prec): src/operators/ConvolutionOperator.cpp:50-62"""

    def __init__(self, mra, coefs, expos, prec):
        _lib.init()
        c = np.ascontiguousarray(coefs, dtype=np.float64)
        e = np.ascontiguousarray(expos, dtype=np.float64)
        super().__init__(mra, _lib.load().mrx_convolution_create(mra._h, len(c), _dp(c), _dp(e), float(prec)))


class ABGVOperator(_Operator):
    """src/operators/ABGVOperator.cpp:46-74"""

    def __init__(self, mra, a, b):
        _lib.init()
        super().__init__(mra, _lib.load().mrx_abgv_create(mra._h, float(a), float(b)))


class PHOperator(_Operator):
    """PHOperator<3>(MRA, order): src/operators/PHOperator.cpp:40-69"""

    def __init__(self, mra, order):
        super().__init__(mra, _lib.load().mrx_ph_create(mra._h, int(order)))
        self.order = order


class BSOperator(_Operator):
    """BSOperator<3>(MRA, order): src/operators/BSOperator.cpp:40-66"""

    def __init__(self, mra, order):
        super().__init__(mra, _lib.load().mrx_bs_create(mra._h, int(order)))
        self.order = order


def poisson_kernel(epsilon, r_min, r_max):
    c = np.zeros(1000)
    e = np.zeros(1000)
    n = _lib.load().mrx_poisson_kernel(epsilon, r_min, r_max, _dp(c), _dp(e), 1000)
    return c[:n].copy(), e[:n].copy()


def helmholtz_kernel(mu, epsilon, r_min, r_max):
    c = np.zeros(1000)
    e = np.zeros(1000)
    n = _lib.load().mrx_helmholtz_kernel(mu, epsilon, r_min, r_max, _dp(c), _dp(e), 1000)
    return c[:n].copy(), e[:n].copy()


def _gauss_arrays(func):
    funcs = list(func) if isinstance(func, (list, tuple)) else [func]
    coef = np.array([g.coef for g in funcs], dtype=np.float64)
    alpha = np.array([g.beta for g in funcs], dtype=np.float64)
    pos = np.ascontiguousarray([g.pos for g in funcs], dtype=np.float64)
    power = np.ascontiguousarray([g.power for g in funcs], dtype=np.int32)
    return len(funcs), coef, alpha, pos, power


def build_grid(out, func, maxIter=-1):
    """build_grid(out, GaussFunc | GaussExp) alone (src/treebuilders/grid.cpp:78-123), or build_grid(out, FunctionTree): extend
    the grid of `out` with the nodes of another tree (grid.cpp:144-153). Host only."""
    if isinstance(func, FunctionTree):
        _lib.load().mrx_tree_build_grid_from(out._h, func._h)
        return
    n, coef, alpha, pos, power = _gauss_arrays(func)
    _lib.load().mrx_build_grid_gaussians(out._h, n, _dp(coef), _dp(alpha), _dp(pos), _ip(power), int(maxIter))


def project(prec, out, func, build_grid=True, finalize=True, device=False):
    """build_grid + project of a Gaussian (expansion): src/treebuilders/project.cpp:85-104, grid.cpp:78-123.
    device=True: the per-node quadrature, cv/mw transforms and norms run on the GPU (mrx_project_gaussians_device)."""
    n, coef, alpha, pos, power = _gauss_arrays(func)
    if device:
        _lib.load().mrx_project_gaussians_device(out._h, float(prec), n, _dp(coef), _dp(alpha), _dp(pos), _ip(power),
                                                 1 if build_grid else 0)
        return
    _lib.load().mrx_project_gaussians(out._h, float(prec), n, _dp(coef), _dp(alpha), _dp(pos), _ip(power),
                                      1 if build_grid else 0, 1 if finalize else 0)


_FUNC3 = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)


def project_function(prec, out, func, finalize=True):
    """project(prec, out, f) of an arbitrary Python callable f(x, y, z) (src/treebuilders/project.cpp:85-104): host
    quadrature through mrx_project_function; the callable is invoked from one thread."""
    cb = _FUNC3(lambda r, _u: float(func(r[0], r[1], r[2])))
    _lib.load().mrx_project_function(out._h, float(prec), C.cast(cb, C.c_void_p), None, 0, 1 if finalize else 0)


def refine_grid(out, prec=None, absPrec=False, scales=0):
    """refine_grid(out, prec, absPrec) or refine_grid(out, scales=n) (src/treebuilders/grid.cpp:271-302); returns the number of
    new nodes"""
    return _lib.load().mrx_tree_refine_grid(out._h, float(-1.0 if prec is None else prec), 1 if absPrec else 0, int(scales))


def clear_grid(out):
    """src/treebuilders/grid.cpp:180-186: keep the grid, drop the coefficients"""
    _lib.load().mrx_tree_clear_grid(out._h)


def copy_func(out, inp):
    """src/treebuilders/grid.cpp:204-208: the function `inp` on the grid `out` enters with"""
    add(-1.0, out, [(1.0, inp)])


def copy_grid(out, inp):
    """src/treebuilders/grid.cpp:150-166"""
    _lib.load().mrx_tree_copy_grid(out._h, inp._h)


class Comm:
    """One rank of a multi-GPU job: NCCL communicator created inside the library. `bcast` ships the 128-byte
    NCCL id from rank 0 to every rank (e.g. a torch.distributed / mpi4py broadcast of a bytes object)."""

    def __init__(self, rank, world, bcast):
        # the library binds to the NCCL copy already mapped into the process; make sure that is the host framework's
        # (torch-bundled) one and not an older system copy that torch itself could not live with afterwards
        try:
            import torch  # noqa: F401
            import torch.distributed  # noqa: F401
        except ImportError:
            pass
        L = _lib.load()
        buf = C.create_string_buffer(128)
        if rank == 0:
            L.mrx_comm_unique_id(buf)
        ident = bcast(bytes(buf.raw))
        self._h = L.mrx_comm_create(int(rank), int(world), C.create_string_buffer(ident, 128))
        self.rank, self.world = rank, world

    def host_arena(self, nbytes):
        """collective: host memory of `nbytes` mapped by all ranks of the node and registered with CUDA in each
        (mrx_comm_host_arena); True when every rank has it"""
        return _lib.load().mrx_comm_host_arena(self._h, int(nbytes)) == 0

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.load().mrx_comm_destroy(self._h)
            self._h = None


def shard_partition(cost, world):
    """contiguous split of one iteration's work vector over `world` ranks (mrx_shard_partition)"""
    import numpy as np
    cost = np.ascontiguousarray(cost, dtype=np.int64)
    begin = (C.c_int * (world + 1))()
    _lib.load().mrx_shard_partition(cost.ctypes.data_as(C.POINTER(C.c_longlong)), len(cost), world, begin)
    return list(begin)


def shard_cyclic(n, world, rank):
    """(count, rows): items of an n-item work vector computed by `rank` under the cyclic distribution of the sharded apply,
    and the padded items per rank of the rank-major exchange buffers (mrx_shard_cyclic)"""
    cnt, rows = C.c_int(0), C.c_int(0)
    _lib.load().mrx_shard_cyclic(int(n), int(world), int(rank), C.byref(cnt), C.byref(rows))
    return cnt.value, rows.value


def shard_block():
    """items dealt out together by the sharded apply's block-cyclic distribution (mrx_shard_block)"""
    return _lib.load().mrx_shard_block()


def shard_cyclic_row(i, n, world):
    """row of work-vector item i in the rank-major exchange buffers (mrx_shard_cyclic_row)"""
    return _lib.load().mrx_shard_cyclic_row(int(i), int(n), int(world))


def apply(prec, out, oper, inp, maxIter=-1, absPrec=False, dir=None, comm=None, precTrees=None):
    """mrcpp::apply. ConvolutionOperator form: apply(prec, out, oper, inp, maxIter, absPrec)
    (src/treebuilders/apply.cpp:68-93); derivative form: apply(None, out, D, inp, dir=d) (:379-412);
    precTrees=[trees]: apply(prec, out, oper, inp, precTrees, maxIter, absPrec) (:214-251), precision scaled per output node
    by the largest norms of the precision trees.
    comm: shard the apply over the ranks of a Comm (collective call). Returns the work counters
    (OperatorStatistics)."""
    st = ApplyStats()
    if precTrees is not None:
        h = (C.c_void_p * max(len(precTrees), 1))(*[t._h for t in precTrees])
        _lib.load().mrx_apply_prec_trees(float(prec), out._h, oper._h, inp._h, len(precTrees), h, int(maxIter), 1 if absPrec else 0,
                                         comm._h if comm is not None else None, C.byref(st))
    elif dir is not None:
        _lib.load().mrx_apply_derivative(out._h, oper._h, inp._h, int(dir), C.byref(st))
    elif comm is not None:
        _lib.load().mrx_apply_sharded(float(prec), out._h, oper._h, inp._h, int(maxIter), 1 if absPrec else 0, comm._h, C.byref(st))
    else:
        _lib.load().mrx_apply(float(prec), out._h, oper._h, inp._h, int(maxIter), 1 if absPrec else 0, C.byref(st))
    return st


def add(prec, out, inp, maxIter=-1, absPrec=False):
    """mrcpp::add(prec, out, FunctionTreeVector, maxIter, absPrec) (src/treebuilders/add.cpp:41-70) from the grid `out` enters
    with: inp = list of (coef, tree). prec < 0 or maxIter = 0: no refinement."""
    c = np.ascontiguousarray([float(ci) for ci, _ in inp], dtype=np.float64)
    h = (C.c_void_p * len(inp))(*[t._h for _, t in inp])
    _lib.load().mrx_tree_add_adaptive(float(prec), out._h, len(inp), _dp(c), h, int(maxIter), 1 if absPrec else 0)


def multiply(prec, out, inp, maxIter=-1, absPrec=False, useMaxNorms=False):
    """mrcpp::multiply(prec, out, FunctionTreeVector, maxIter, absPrec, useMaxNorms) (src/treebuilders/multiply.cpp:104-136): inp =
    list of (coef, tree); prec < 0 or maxIter = 0: no refinement of the grid `out` enters with"""
    c = np.ascontiguousarray([float(ci) for ci, _ in inp], dtype=np.float64)
    h = (C.c_void_p * len(inp))(*[t._h for _, t in inp])
    _lib.load().mrx_tree_multiply(float(prec), out._h, len(inp), _dp(c), h, int(maxIter), 1 if absPrec else 0, 1 if useMaxNorms else 0)


def power(prec, out, inp, p, maxIter=-1, absPrec=False):
    """mrcpp::power(prec, out, inp, p) (src/treebuilders/multiply.cpp:211-234): function values of `inp` raised to the power p"""
    _lib.load().mrx_tree_power(float(prec), out._h, inp._h, float(p), int(maxIter), 1 if absPrec else 0)


def dot_vectors(prec, out, inp_a, inp_b, maxIter=-1, absPrec=False):
    """mrcpp::dot(prec, out, FunctionTreeVector, FunctionTreeVector) (src/treebuilders/multiply.cpp:253-271): out = sum_d a_d b_d f_d g_d"""
    if len(inp_a) != len(inp_b):
        raise ValueError("Input length mismatch")
    parts = []
    for (ca, ta), (cb, tb) in zip(inp_a, inp_b):
        p = FunctionTree(out.mra)
        build_grid(p, out)
        multiply(prec, p, [(1.0, ta), (1.0, tb)], maxIter, absPrec, True)
        parts.append((ca * cb, p))
    for _, p in parts:
        build_grid(out, p)
    add(-1.0, out, parts, 0)


def gradient(oper, inp):
    """mrcpp::gradient(D, f) (src/treebuilders/apply.cpp:444-452): [(1.0, df/dx), (1.0, df/dy), (1.0, df/dz)]"""
    out = []
    for d in range(3):
        g = FunctionTree(inp.mra)
        apply(None, g, oper, inp, dir=d)
        out.append((1.0, g))
    return out


def divergence(out, oper, inp):
    """mrcpp::divergence(out, D, FunctionTreeVector) (src/treebuilders/apply.cpp:514-530): derivative of component d along d,
    union grid, sum. inp = list of three (coef, tree)."""
    if len(inp) != 3:
        raise ValueError("Dimension mismatch")
    parts = []
    for d, (c, t) in enumerate(inp):
        p = FunctionTree(t.mra)
        apply(None, p, oper, t, dir=d)
        parts.append((c, p))
    for _, p in parts:
        build_grid(out, p)
    add(-1.0, out, parts, 0)


def dot(bra, ket):
    """src/treebuilders/multiply.cpp:286-318"""
    return _lib.load().mrx_dot(bra._h, ket._h)
