"""ctypes loader for libmrcpp_b200.so (the C-ABI declared in include/mrcpp_b200.h).

The library is built in-tree by mrcpp_b200/build.py. Loading fails loudly if it is missing: there is
no Python or CPU fallback for the hot path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmrcpp_b200.so")  # the one library of the product; no switch, no fallback
TABLES = os.path.join(_HERE, "data", "mwtables.bin")


class ApplyStats(C.Structure):
    _fields_ = [
        ("g_nodes", C.c_longlong),
        ("f_applied", C.c_longlong),
        ("gen_nodes", C.c_longlong),
        ("iterations", C.c_int),
        ("n_nodes_out", C.c_int),
        ("ms_upload", C.c_double),
        ("ms_build", C.c_double),
        ("ms_kernel", C.c_double),
        ("ms_contract", C.c_double),
        ("ms_post", C.c_double),
        ("ms_download", C.c_double),
        ("kernel_launches", C.c_longlong),
        ("f_applied_rank", C.c_longlong),
        ("h2d_bytes", C.c_longlong),
        ("d2h_bytes", C.c_longlong),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_PI = C.POINTER(C.c_int)
_PD = C.POINTER(C.c_double)

# name -> (restype, argtypes); every symbol include/mrcpp_b200.h declares
SIGNATURES = {
    "mrx_init": (_I, [C.c_char_p, _I]),
    "mrx_device_count": (_I, []),
    "mrx_version": (C.c_char_p, []),
    "mrx_mra_create": (_P, [_I, _I, _PI, _PI, _I]),
    "mrx_mra_set_periodic": (_I, [_P, _I]),
    "mrx_mra_destroy": (None, [_P]),
    "mrx_tree_create": (_P, [_P]),
    "mrx_tree_destroy": (None, [_P]),
    "mrx_tree_n_nodes": (_I, [_P]),
    "mrx_tree_n_end_nodes": (_I, [_P]),
    "mrx_tree_square_norm": (_D, [_P]),
    "mrx_tree_clear": (None, [_P]),
    "mrx_tree_from_arrays": (_P, [_P, _I, _PI, _PI, _PI, _PI, _PD]),
    "mrx_tree_to_arrays": (_I, [_P, _PI, _PI, _PI, _PI, _PD, _PD]),
    "mrx_tree_copy_grid": (_I, [_P, _P]),
    "mrx_tree_integrate": (_D, [_P]),
    "mrx_tree_save_txt": (_I, [_P, C.c_char_p]),
    "mrx_tree_load_txt": (_I, [_P, C.c_char_p]),
    "mrx_tree_evalf": (_I, [_P, _I, _PD, _PD, _I]),
    "mrx_tree_build_grid_from": (_I, [_P, _P]),
    "mrx_tree_clear_grid": (_I, [_P]),
    "mrx_tree_add": (_I, [_P, _I, _PD, C.POINTER(C.c_void_p)]),
    "mrx_tree_refine_grid": (_I, [_P, _D, _I, _I]),
    "mrx_tree_add_inplace": (_I, [_P, _D, _P]),
    "mrx_tree_multiply": (_I, [_D, _P, _I, _PD, C.POINTER(C.c_void_p), _I, _I, _I]),
    "mrx_tree_power": (_I, [_D, _P, _P, _D, _I, _I]),
    "mrx_tree_add_adaptive": (_I, [_D, _P, _I, _PD, C.POINTER(C.c_void_p), _I, _I]),
    "mrx_build_grid_gaussians": (_I, [_P, _I, _PD, _PD, _PD, _PI, _I]),
    "mrx_project_gaussians": (_I, [_P, _D, _I, _PD, _PD, _PD, _PI, _I, _I]),
    "mrx_project_gaussians_device": (_I, [_P, _D, _I, _PD, _PD, _PD, _PI, _I]),
    "mrx_quadrature": (_I, [_I, _PD, _PD]),
    "mrx_interp_scaling": (_D, [_I, _I, _D, _I]),
    "mrx_project_function": (_I, [_P, _D, C.c_void_p, _P, _I, _I]),
    "mrx_poisson_create": (_P, [_P, _D]),
    "mrx_helmholtz_create": (_P, [_P, _D, _D]),
    "mrx_convolution_create": (_P, [_P, _I, _PD, _PD, _D]),
    "mrx_poisson_create_reach": (_P, [_P, _D, _I, _I]),
    "mrx_helmholtz_create_reach": (_P, [_P, _D, _D, _I, _I]),
    "mrx_convolution_create_reach": (_P, [_P, _I, _PD, _PD, _D, _I, _I]),
    "mrx_oper_cache_stats": (None, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "mrx_abgv_create": (_P, [_P, _D, _D]),
    "mrx_ph_create": (_P, [_P, _I]),
    "mrx_bs_create": (_P, [_P, _I]),
    "mrx_oper_from_arrays": (_P, [_P, _I, _PI, _PI, _PD, _PD, _I, _I, _D]),
    "mrx_oper_destroy": (None, [_P]),
    "mrx_oper_n_terms": (_I, [_P]),
    "mrx_oper_band_widths": (_I, [_P, _D, _PI, _I]),
    "mrx_oper_node": (_I, [_P, _I, _I, _I, _PD, _PD]),
    "mrx_oper_depth": (_I, [_P, _I]),
    "mrx_oper_max_transl": (_I, [_P, _I, _I]),
    "mrx_poisson_kernel": (_I, [_D, _D, _D, _PD, _PD, _I]),
    "mrx_helmholtz_kernel": (_I, [_D, _D, _D, _D, _PD, _PD, _I]),
    "mrx_apply": (_I, [_D, _P, _P, _P, _I, _I, C.POINTER(ApplyStats)]),
    "mrx_apply_sharded": (_I, [_D, _P, _P, _P, _I, _I, _P, C.POINTER(ApplyStats)]),
    "mrx_apply_unit_cell": (_I, [_I, _D, _P, _P, _P, _I, _I, C.POINTER(ApplyStats)]),
    "mrx_project_cosines": (_I, [_P, _D, _I, _PD, _PD, _I]),
    "mrx_apply_prec_trees": (_I, [_D, _P, _P, _P, _I, C.POINTER(C.c_void_p), _I, _I, _P, C.POINTER(ApplyStats)]),
    "mrx_comm_unique_id": (_I, [C.c_char_p]),
    "mrx_comm_create": (_P, [_I, _I, C.c_char_p]),
    "mrx_comm_destroy": (None, [_P]),
    "mrx_comm_rank": (_I, [_P]),
    "mrx_comm_size": (_I, [_P]),
    "mrx_shard_partition": (None, [C.POINTER(C.c_longlong), _I, _I, _PI]),
    "mrx_shard_cyclic": (None, [_I, _I, _I, _PI, _PI]),
    "mrx_shard_cyclic_row": (_I, [_I, _I, _I]),
    "mrx_shard_block": (_I, []),
    "mrx_apply_derivative": (_I, [_P, _P, _P, _I, C.POINTER(ApplyStats)]),
    "mrx_mw_transform": (_I, [_P, _I, _I]),
    "mrx_node_mw_transform": (_I, [_P, _I, _I, _PI]),
    "mrx_node_cv_transform": (_I, [_P, _I, _I, _PI]),
    "mrx_bench_cv_transform": (_D, [_P, _I]),
    "mrx_calc_square_norm": (_D, [_P]),
    "mrx_dot": (_D, [_P, _P]),
    "mrx_tree_rescale": (_I, [_P, _D]),
    "mrx_tree_sync_device": (_I, [_P]),
    "mrx_tree_sync_host": (_I, [_P]),
    "mrx_tree_set_host_mirror": (_I, [_P, _I]),
    "mrx_tree_set_shared_host_mirror": (_I, [_P, _P]),
    "mrx_comm_host_arena": (_I, [_P, C.c_longlong]),
    "mrx_tree_drop_device": (_I, [_P]),
    "mrx_tree_bytes": (C.c_longlong, [_P]),
    "mrx_tree_host_handle": (_P, [_P]),
    "mrx_oper_host_handle": (_P, [_P]),
    "mrx_tree_host_modified": (None, [_P]),
    "mrx_timer_start": (None, []),
    "mrx_timer_stop_ms": (_D, []),
    "mrx_bench_dmma_tflops": (_D, [_I]),
    "mrx_bench_dfma_tflops": (_D, [_I]),
    "mrx_bench_hbm_gbs": (_D, [C.c_longlong, _I]),
    "mrx_bench_mw_transform": (_D, [_P, _I, _I, _PI]),
}

_lib = None
_device = None


def load():
    """Load the shared library and bind every C-ABI symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m mrcpp_b200.build` (or __graft_entry__.build()). "
            "There is no fallback implementation.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def init(device=None):
    """Initialise tables and (optionally) the CUDA device. device=None: cuda:LOCAL_RANK if any GPU is
    visible, else host-only (hot-path calls then abort)."""
    global _device
    lib = load()
    if device is None:
        n = lib.mrx_device_count()
        device = int(os.environ.get("LOCAL_RANK", "0")) if n > 0 else -1
    if _device is not None and _device == device:
        return device
    rc = lib.mrx_init(TABLES.encode(), device)
    if rc != 0:
        raise RuntimeError(f"mrx_init failed with status {rc}")
    _device = device
    return device


def device():
    return _device
