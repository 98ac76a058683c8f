"""ctypes binding of libilm_b200.so (include/ilm_b200.h).  Fails loudly when the
CUDA library has not been built -- there is no CPU fallback in the product."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ILM_B200_LIB") or os.path.join(_HERE, "libilm_b200.so")     # ILM_B200_LIB: tuning builds (tools/)

OK, EINVAL, ESIZE, ECUDA, ENCCL, ENOMEM = range(6)

NODES_PRIMAL, NODES_DUAL, XEDGES, YEDGES, EDGES, EDGEGRAD = range(6)
SCALAR_CACHE, VECTOR_CACHE = 0, 1
NORMAL, CROSS = 0, 1
RTLINVR, CLINVCT, GLINVD, GLINVD_CROSS = range(4)
DDF = {"yang3": 0, "m3": 1, "roma": 2, "m4prime": 3, "witchhat": 4, "goza": 5}    # goza: named, answered with ILM_EINVAL
GRID_SCALING, INDEX_SCALING = 0, 1


class ilm_grid(C.Structure):
    _fields_ = [("NX", C.c_int), ("NY", C.c_int), ("dx", C.c_double), ("I0x", C.c_int), ("I0y", C.c_int)]


class ilm_slab_info(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nranks", "rank", "Lx", "Ly", "MYp", "row0", "row1", "ntc", "tc0", "tc1", "cpw", "wlo", "whi")]


class IlmError(RuntimeError):
    """Generic library failure (ILM_ECUDA / ILM_ENOMEM / ILM_ENCCL)."""


class DimensionMismatch(ValueError):
    """ILM_ESIZE -- the reference throws DimensionMismatch / AssertionError."""


class MethodError(TypeError):
    """ILM_EINVAL -- the reference throws MethodError / ArgumentError."""


_vp, _dp, _ip, _i, _d = C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double

# name -> (restype, argtypes); every symbol declared in include/ilm_b200.h
SIGNATURES = {
    "ilm_plan_create": (_i, [C.POINTER(ilm_grid), _i, _dp, _dp, _dp, _dp, _dp, _i, _i, _dp, _i, _d, _d, _vp, C.POINTER(_vp)]),
    "ilm_plan_create_shared": (_i, [_vp, _i, _dp, _dp, _dp, _dp, _dp, _i, _i, C.POINTER(_vp)]),
    "ilm_plan_update_points": (_i, [_vp, _i, _dp, _dp, _dp, _dp, _dp]),
    "ilm_plan_destroy": (None, [_vp]),
    "ilm_plan_sync": (_i, [_vp]),
    "ilm_last_error": (C.c_char_p, []),
    "ilm_plan_npoints": (_i, [_vp]),
    "ilm_layout_size": (C.c_int64, [_vp, _i]),
    "ilm_plan_launch_count": (C.c_int64, [_vp]),
    "ilm_get_table": (_i, [_vp, _i, C.POINTER(_i), _vp, _dp, _dp]),
    "ilm_regularize": (_i, [_vp, _i, _dp, _dp]),
    "ilm_interpolate": (_i, [_vp, _i, _dp, _dp]),
    "ilm_forcing_area_add": (_i, [_vp, _i, _dp, _dp, _dp]),
    "ilm_forcing_line_add": (_i, [_vp, _i, _dp, _dp]),
    "ilm_regularize_normal": (_i, [_vp, _i, _dp, _dp]),
    "ilm_normal_interpolate": (_i, [_vp, _i, _dp, _dp]),
    "ilm_divergence": (_i, [_vp, _dp, _dp]),
    "ilm_grad": (_i, [_vp, _dp, _dp]),
    "ilm_curl_n2e": (_i, [_vp, _dp, _dp]),
    "ilm_curl_e2n": (_i, [_vp, _dp, _dp]),
    "ilm_laplacian": (_i, [_vp, _i, _dp, _dp]),
    "ilm_inverse_laplacian": (_i, [_vp, _i, _dp]),
    "ilm_inverse_laplacian_pair": (_i, [_vp, _i, _dp, _i, _dp]),
    "ilm_add_kernel": (_i, [_vp, _dp, _i, _d, _d, C.POINTER(_i)]),
    "ilm_convolve": (_i, [_vp, _i, _i, _dp]),
    "ilm_surface_divergence": (_i, [_vp, _i, _dp, _dp]),
    "ilm_surface_grad": (_i, [_vp, _i, _dp, _dp]),
    "ilm_surface_curl_s2n": (_i, [_vp, _i, _dp, _dp]),
    "ilm_surface_curl_n2s": (_i, [_vp, _i, _dp, _dp]),
    "ilm_mask": (_i, [_vp, _dp]),
    "ilm_create_schur": (_i, [_vp, _i, _d, _i, _i, _dp]),
    "ilm_create_schur_kernel": (_i, [_vp, _i, _i, _d, _i, _i, _dp]),
    "ilm_create_RTLinvR_direct": (_i, [_vp, _d, _i, _i, _dp]),
    "ilm_create_schur_direct_kernel": (_i, [_vp, _dp, _i, _d, _d, _d, _i, _i, _dp]),
    "ilm_create_nRTRn": (_i, [_vp, _d, _dp]),
    "ilm_create_surface_filter": (_i, [_vp, _dp]),
    "ilm_dense_factor": (_i, [_i, _dp, _ip, _vp]),
    "ilm_dense_solve": (_i, [_i, _dp, _ip, _i, _dp, _vp]),
    "ilm_dense_matvec_pow": (_i, [_i, _dp, _i, _dp, _vp]),
    "ilm_regularize_normal_tensor": (_i, [_vp, _i, _dp, _dp]),
    "ilm_normal_interpolate_tensor": (_i, [_vp, _i, _dp, _dp]),
    "ilm_regularize_normal_vs": (_i, [_vp, _i, _dp, _dp]),
    "ilm_normal_interpolate_vs": (_i, [_vp, _i, _dp, _dp]),
    "ilm_regularize_normal_dot_tensor": (_i, [_vp, _dp, _dp]),
    "ilm_normal_dot_interpolate_tensor": (_i, [_vp, _dp, _dp]),
    "ilm_grad_tensor": (_i, [_vp, _dp, _dp]),
    "ilm_divergence_tensor": (_i, [_vp, _dp, _dp]),
    "ilm_vsurface_divergence": (_i, [_vp, _i, _dp, _dp]),
    "ilm_vsurface_grad": (_i, [_vp, _i, _dp, _dp]),
    "ilm_vsurface_curl_s2n": (_i, [_vp, _dp, _dp]),
    "ilm_vsurface_curl_n2s": (_i, [_vp, _dp, _dp]),
    "ilm_mask_edges": (_i, [_vp, _dp]),
    "ilm_mask_product": (_i, [_vp, _i, _i, _i, _dp]),
    "ilm_create_schur_vector": (_i, [_vp, _i, _d, _i, _i, _dp]),
    "ilm_create_nRTRn_vector": (_i, [_vp, _d, _dp]),
    "ilm_helmholtz_potentials": (_i, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "ilm_vecfield_from_potentials": (_i, [_vp, _dp, _dp, _dp, _dp]),
    "ilm_vecfield_helmholtz": (_i, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "ilm_helmholtz_jump_add": (_i, [_vp, _i, _i, _dp, _dp, _dp]),
    "ilm_convective_derivative_scalar": (_i, [_vp, _dp, _dp, _dp]),
    "ilm_convective_derivative_dual": (_i, [_vp, _dp, _dp, _dp]),
    "ilm_convective_derivative_vector": (_i, [_vp, _dp, _dp, _dp]),
    "ilm_w_cross_v": (_i, [_vp, _dp, _dp, _dp]),
    "ilm_dense_launch_count": (C.c_int64, []),
    "ilm_slab_partition": (_i, [_i, _i, _i, _i, _i, C.POINTER(ilm_slab_info)]),
    "ilm_slab_counts": (_i, [_i, _i, C.POINTER(ilm_slab_info), _i, _vp, _vp]),
    "ilm_slab_buffer_doubles": (C.c_int64, [C.POINTER(ilm_slab_info)]),
    "ilm_slab_forward": (_i, [_vp, C.POINTER(ilm_slab_info), _i, _dp, _i, _dp, _dp]),
    "ilm_slab_columns": (_i, [_vp, C.POINTER(ilm_slab_info), _i, _dp, _dp]),
    "ilm_slab_inverse": (_i, [_vp, C.POINTER(ilm_slab_info), _dp, _i, _dp, _i, _dp]),
    "ilm_plan_release_spectrum": (_i, [_vp]),
    "ilm_comm_unique_id": (_i, [_vp, _i]),
    "ilm_comm_init": (_i, [_vp, _vp, _i, _i, _i]),
    "ilm_comm_destroy": (_i, [_vp]),
    "ilm_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "ilm_column_range": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "ilm_create_schur_sharded": (_i, [_vp, _i, _i, _d, _dp]),
    "ilm_dirichlet_poisson": (_i, [_vp, _dp, _dp, _dp, _dp, _dp]),
    "ilm_dirichlet_poisson_rows": (_i, [_vp, _dp, _dp, _dp, _i, _i, _dp, _dp]),
    "ilm_slab_solve": (_i, [_vp, C.POINTER(ilm_slab_info), _i, _i, _dp, _i, _dp, _dp, _dp]),
    "ilm_profile_conv": (_i, [_vp, _i, _i, C.POINTER(C.c_double * 3)]),
    "ilm_profile_conv_probe": (_i, [_vp, _i, _i, C.POINTER(C.c_double * 3)]),
    "ilm_probe_output_rows": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
}

_lib = None


def load():
    """Load (once) and return the CDLL with typed signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IlmError(
            f"{LIB_PATH} not found: build it with `make -C immersedlayers.jl_b200/csrc -j8` "
            "(or __graft_entry__.build()); ilm_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status == OK:
        return
    msg = load().ilm_last_error().decode(errors="replace")
    if status == ESIZE:
        raise DimensionMismatch(msg)
    if status == EINVAL:
        raise MethodError(msg)
    raise IlmError(f"status {status}: {msg}")
