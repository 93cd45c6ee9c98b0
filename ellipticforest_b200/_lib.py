"""ctypes binding of the C-ABI in include/efgpu.h (libefgpu.so, built in-tree by __graft_entry__.build()).

There is no fallback: if the shared library is missing or no CUDA device is visible the product
path raises.  Nothing under oracle/ is imported from this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libefgpu.so")


class TreeDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("nx", C.c_int32), ("level", C.POINTER(C.c_int32)),
                ("child", C.POINTER(C.c_int32)), ("box", C.POINTER(C.c_double))]


class Stats(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "dofs", "n_leaves", "n_nodes", "build_ms", "upwards_ms", "solve_ms", "merge_flops_canonical",
        "merge_flops_issued", "upwards_bytes", "solve_bytes", "device_bytes", "min_pivot", "max_pivot", "pivot_ratio_min", "negative_pivots", "inverse_residual")]


REFINE_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_double, C.c_void_p)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p)

# every symbol include/efgpu.h declares: (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
SIGNATURES = {
    "efgpu_create": (C.c_int, [C.POINTER(TreeDesc), C.c_int, C.POINTER(_P)]),
    "efgpu_create_ex": (C.c_int, [C.POINTER(TreeDesc), C.c_int, C.POINTER(C.c_int32), C.POINTER(_P)]),
    "efgpu_set_partition": (C.c_int, [_P, C.c_int, C.c_int]),
    "efgpu_set_allgather": (C.c_int, [_P, ALLGATHER_FN, _P]),
    "efgpu_complete_root_dtn": (C.c_int, [_P]),
    "efgpu_peer_export": (C.c_int, [_P, _P]),
    "efgpu_peer_attach": (C.c_int, [_P, _P, C.c_int]),
    "efgpu_peer_barrier": (C.c_int, [_P]),
    "efgpu_peer_broadcast": (C.c_int, [_P, _P, C.c_size_t]),
    "efgpu_set_tuning": (C.c_int, [C.c_int, C.c_int]),
    "efgpu_set_symmetric_leaves": (C.c_int, [_P, C.c_int]),
    "efgpu_is_symmetric": (C.c_int, [_P]),
    "efgpu_set_refine_inverse": (C.c_int, [_P, C.c_int]),
    "efgpu_debug_merge_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _I, _P, _P, _I, _P, _I, _P]),
    "efgpu_debug_merge_plan_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _I, _P, _P, _I, _P, _I, _P]),
    "efgpu_debug_tma_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _I, _P, _P]),
    "efgpu_build_begin": (C.c_int, [_P, C.c_uint]),
    "efgpu_build_level": (C.c_int, [_P, C.c_int, C.c_int]),
    "efgpu_build_end": (C.c_int, [_P]),
    "efgpu_max_level": (C.c_int, [_P]),
    "efgpu_destroy": (None, [_P]),
    "efgpu_last_error": (C.c_char_p, [_P]),
    "efgpu_set_leaf_constant": (C.c_int, [_P, C.c_double]),
    "efgpu_set_leaf_variable": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "efgpu_set_leaf_variable_device": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "efgpu_leaf_points_device": (C.c_int, [_P, C.c_int, _P, _P, C.c_int]),
    "efgpu_leaf_points": (C.c_int, [_P, C.c_int, _P, _P]),
    "efgpu_error_norms_device": (C.c_int, [_P, _P, _P, _D, _D, _D]),
    "efgpu_error_norms": (C.c_int, [_P, _P, _D, _D, _D]),
    "efgpu_build": (C.c_int, [_P, C.c_uint]),
    "efgpu_rebuild_from": (C.c_int, [_P, _P, C.c_uint, _D, _D]),
    "efgpu_upwards": (C.c_int, [_P, _P, C.c_double, C.c_uint]),
    "efgpu_upwards_device": (C.c_int, [_P, _P, C.c_double, C.c_uint, C.c_int]),
    "efgpu_solve_dirichlet": (C.c_int, [_P, _P, C.c_uint, _P]),
    "efgpu_solve_dirichlet_device": (C.c_int, [_P, _P, C.c_uint, _P, C.c_int]),
    "efgpu_solve_from_roots_device": (C.c_int, [_P, C.c_uint, _P, C.c_int]),
    "efgpu_operator_device": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), _I, _I]),
    "efgpu_vector_device": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), _I]),
    "efgpu_solve_robin": (C.c_int, [_P, _P, _P, _P, C.c_uint, _P]),
    "efgpu_sync": (C.c_int, [_P]),
    "efgpu_write_vtu": (C.c_int, [_P, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(_P)]),
    "efgpu_stream": (_P, [_P]),
    "efgpu_set_stream": (C.c_int, [_P, _P]),
    "efgpu_node_info": (C.c_int, [_P, C.c_int, _I, _I, _I, _I]),
    "efgpu_operator_shape": (C.c_int, [_P, C.c_int, C.c_int, _I, _I]),
    "efgpu_get_operator": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "efgpu_vector_length": (C.c_int, [_P, C.c_int, C.c_int, _I]),
    "efgpu_get_vector": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "efgpu_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "efgpu_set_profiling": (C.c_int, [_P, C.c_int]),
    "efgpu_get_profile": (C.c_int, [_P, C.c_int, _D, _D]),
    "efgpu_profile_class_name": (C.c_char_p, [C.c_int]),
    "efgpu_refine_elliptic_single": (C.c_int, [C.c_double, C.c_double, _P]),
    "efgpu_mesh_create": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, REFINE_FN, _P, C.POINTER(_P)]),
    "efgpu_mesh_desc": (C.c_int, [_P, C.POINTER(TreeDesc)]),
    "efgpu_mesh_n_leaves": (C.c_int, [_P]),
    "efgpu_mesh_leaf_nodes": (C.POINTER(C.c_int32), [_P]),
    "efgpu_mesh_path": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_size_t]),
    "efgpu_mesh_destroy": (None, [_P]),
    "efgpu_dgemm_batched": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "efgpu_dgemm_batched_tma": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
}

_lib = None


def load():
    """Load libefgpu.so (raises if it has not been built: run `python -c "import __graft_entry__ as g; g.build()"`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class EfgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("efgpu status %d: %s" % (code, msg))
        self.code = code


def check(code, handle=None):
    if code != 0:
        msg = load().efgpu_last_error(handle)
        raise EfgpuError(code, msg.decode() if msg else "")
