"""ctypes binding of libmgb200.so -- the C ABI declared in include/mgb200.h.

Loading fails loudly (ImportError-like RuntimeError) when the library has not been built; creating
an engine fails loudly when no CUDA device is present.  There is no CPU fallback anywhere in this
package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgb200.so")

# constants mirrored from include/mgb200.h
OK = 0
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM, ERR_SINGULAR, ERR_UNSUPPORTED, ERR_COMM = -1, -2, -3, -4, -5, -6, -7
R_INJECTION, R_FULL_WEIGHTING, R_TRANSPOSE, R_EXPLICIT = 0, 1, 2, 3
SM_JACOBI_RJ, SM_JACOBI_A, SM_GS_LEVEL, SM_GS_MULTICOLOR = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
(ART_RJ_INDPTR, ART_RJ_INDICES, ART_RJ_VALUES, ART_DINV, ART_LEVEL_OF_ROW, ART_LEVEL_ORDER, ART_LEVEL_OFFSETS,
 ART_COLOUR_OF_ROW, ART_COLOUR_ORDER, ART_COLOUR_OFFSETS, ART_R_INDPTR, ART_R_INDICES, ART_R_VALUES,
 ART_COARSE_INVERSE, ART_A_INDPTR, ART_A_INDICES, ART_A_VALUES, ART_P_INDPTR, ART_P_INDICES, ART_P_VALUES,
 ART_INJECTION) = range(21)
BUF_V, BUF_F, BUF_R = 0, 1, 2
KERNEL_KINDS = ["jacobi", "residual", "restrict", "prolong_add", "coarse", "init_guess", "gs", "norm", "spmv", "halo", "copy", "jacobi2"]

R_MODES = {"injection": R_INJECTION, "full_weighting": R_FULL_WEIGHTING, "transpose": R_TRANSPOSE, "explicit": R_EXPLICIT}
SMOOTHERS = {"jacobi": SM_JACOBI_RJ, "jacobi_rj": SM_JACOBI_RJ, "jacobi_a": SM_JACOBI_A, "gs": SM_GS_LEVEL,
             "gs_level": SM_GS_LEVEL, "gs_color": SM_GS_MULTICOLOR, "gs_multicolor": SM_GS_MULTICOLOR}


class ProfileRecord(C.Structure):
    _fields_ = [("kind", C.c_int32), ("level", C.c_int32), ("launches", C.c_int64), ("total_ms", C.c_double),
                ("bytes", C.c_double), ("moved_bytes", C.c_double)]


# every symbol include/mgb200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = {
    "mgb_version": (_i, []),
    "mgb_create": (_i, [C.POINTER(_vp), _i]),
    "mgb_destroy": (_i, [_vp]),
    "mgb_last_error": (C.c_char_p, [_vp]),
    "mgb_set_level": (_i, [_vp, _i, _i64, _i64, _vp, _i, _vp, _vp]),
    "mgb_set_transfer": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _i, _vp, _vp, _i, _i, _vp, _i64, _vp, _i, _vp, _vp]),
    "mgb_dist_unique_id": (_i, [_vp, _i]),
    "mgb_dist_init": (_i, [_vp, _i, _i, _vp, _i]),
    "mgb_set_level_local": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _i, _vp, _vp]),
    "mgb_set_halo": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "mgb_set_gather_level": (_i, [_vp, _i, _i64, _vp]),
    "mgb_p2p_export": (_i, [_vp, _i, _vp, _i, C.POINTER(_i)]),
    "mgb_p2p_import": (_i, [_vp, _i, _i, _vp, _i]),
    "mgb_synth_poisson_level": (_i, [_vp, _i, _i, _i, _i64, _i64, _i64, _i64]),
    "mgb_synth_poisson_transfer": (_i, [_vp, _i, _i64, _i64]),
    "mgb_set_params": (_i, [_vp, _d, _i, _i, _i]),
    "mgb_set_option": (_i, [_vp, C.c_char_p, _d]),
    "mgb_finalize": (_i, [_vp]),
    "mgb_precheck": (_i, [_vp]),
    "mgb_vcycle": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp]),
    "mgb_vcycle_debug": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "mgb_vcycle_resident": (_i, [_vp, _i, _i, _vp]),
    "mgb_set_rhs": (_i, [_vp, _i, _vp, _i]),
    "mgb_set_mass_matrix": (_i, [_vp, _i, _i64, _i64, _vp, _i, _vp, _vp]),
    "mgb_fmg": (_i, [_vp, _i, _d, _i, _vp, _i, C.POINTER(_i), _vp, _i]),
    "mgb_set_exact_solution": (_i, [_vp, _i, _vp, _i]),
    "mgb_set_numbering": (_i, [_vp, _i, _i64, _vp]),
    "mgb_set_restriction": (_i, [_vp, _i, _i, _i]),
    "mgb_halo_fused": (_i, [_vp, _i, C.POINTER(_i)]),
    "mgb_set_halo_fused": (_i, [_vp, _i, _i]),
    "mgb_fmg_error_history": (_i, [_vp, _vp, _i, C.POINTER(_i)]),
    "mgb_spmv": (_i, [_vp, _i, _vp, _vp, _i]),
    "mgb_residual": (_i, [_vp, _i, _vp, _vp, _vp, _i]),
    "mgb_smooth": (_i, [_vp, _i, _vp, _vp, _i, _i]),
    "mgb_restrict": (_i, [_vp, _i, _vp, _vp, _i]),
    "mgb_prolong_add": (_i, [_vp, _i, _vp, _vp, _i]),
    "mgb_coarse_solve": (_i, [_vp, _vp, _vp, _i]),
    "mgb_norm2": (_i, [_vp, _i64, _vp, _i, C.POINTER(_d)]),
    "mgb_get_artifact": (_i, [_vp, _i, _i, _vp, _i64, C.POINTER(_i64)]),
    "mgb_level_buffer": (_i, [_vp, _i, _i, C.POINTER(_vp), C.POINTER(_i64)]),
    "mgb_get_stream": (_i, [_vp, C.POINTER(_vp)]),
    "mgb_synchronize": (_i, [_vp]),
    "mgb_launch_count": (_i, [_vp, C.POINTER(_i64)]),
    "mgb_profile_begin": (_i, [_vp]),
    "mgb_profile_end": (_i, [_vp]),
    "mgb_profile_get": (_i, [_vp, C.POINTER(ProfileRecord), _i, C.POINTER(_i)]),
    "mgb_vcycle_bytes": (_i, [_vp, _i, C.POINTER(_d)]),
    "mgb_vcycle_bytes_moved": (_i, [_vp, _i, C.POINTER(_d)]),
    "mgb_describe": (_i, [_vp, C.c_char_p, _i64]),
    "mgb_host_build_rj": (_i, [_i64, _vp, _vp, _vp, _i, C.POINTER(_i64), _vp, _vp, _vp, _vp]),
    "mgb_host_level_sets": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp, _i64]),
    "mgb_host_colouring": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i64), _vp, _i64]),
    "mgb_host_dense_inverse": (_i, [_i64, _vp, _vp, _vp, _vp]),
    "mgb_host_make_tiles": (_i, [_i64, _vp, _i64, _i64, _i, _vp, _i64, _vp, _i64, C.POINTER(_i64), _vp]),
    "mgb_host_code_operator": (_i, [_i64, _i64, _vp, _vp, _vp, _i, C.POINTER(_i), C.POINTER(_i), _vp, _vp, C.POINTER(_i), _vp]),
}

_lib = None


def load():
    """Load libmgb200.so (building is the job of multigrid_dolfinx_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m multigrid_dolfinx_b200.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MGBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mgb error {code}: {msg}")
        self.code = code
