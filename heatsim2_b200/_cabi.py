"""ctypes binding of libhs2b200.so (include/hs2_b200.h).

Loading is explicit and loud: if the CUDA library has not been built, or fails
to load, an ImportError/RuntimeError is raised - there is no fallback path.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# HS2_B200_LIB: another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("HS2_B200_LIB") or os.path.join(_PKG, "libhs2b200.so")

HS2_COEF_STRIDE = 8
HS2_LU_STRIDE = 4
ABI_VERSION = 6

c_void_p = ctypes.c_void_p
c_double_p = ctypes.POINTER(ctypes.c_double)


class AxisTables(ctypes.Structure):
    _fields_ = [
        ("d_line_id", c_void_p), ("d_lu", c_void_p),
        ("d_tab", c_void_p), ("d_GE", c_void_p), ("d_tab_il", c_void_p),
        ("h_utab", c_void_p), ("d_ucode", c_void_p),
        ("n_unique", ctypes.c_int32), ("chunk", ctypes.c_int32), ("n_chunks", ctypes.c_int32),
        ("pitch", ctypes.c_int32), ("band", ctypes.c_int32), ("xw_band", ctypes.c_int32),
        ("d_xw_tab", c_void_p), ("d_xw_code", c_void_p),
    ]


class PlanDesc(ctypes.Structure):
    _fields_ = [
        ("nz", ctypes.c_int64), ("ny", ctypes.c_int64), ("nx", ctypes.c_int64),
        ("n_classes", ctypes.c_int32), ("class_id_bytes", ctypes.c_int32),
        ("d_class_id", c_void_p), ("d_class_coef", c_void_p),
        ("axis", AxisTables * 3),
        ("device", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("z_chunk0", ctypes.c_int32), ("z_chunks_global", ctypes.c_int32),
    ]


class Source(ctypes.Structure):
    _fields_ = [("d_vol_elements", c_void_p), ("h_value", c_double_p), ("d_dense", c_void_p)]


class BuildDesc(ctypes.Structure):
    _fields_ = [
        ("nz", ctypes.c_int64), ("ny", ctypes.c_int64), ("nx", ctypes.c_int64),
        ("n_classes", ctypes.c_int32), ("class_id_bytes", ctypes.c_int32),
        ("d_class_id", c_void_p), ("h_class_coef", c_double_p),
        ("device", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("chunk", ctypes.c_int32 * 3), ("utab_axes", ctypes.c_int32),
        ("d_class_id_global", c_void_p), ("nz_global", ctypes.c_int64), ("k0", ctypes.c_int64),
    ]


class AxisInfo(ctypes.Structure):
    _fields_ = [
        ("line_length", ctypes.c_int64), ("n_lines", ctypes.c_int64),
        ("n_unique", ctypes.c_int32), ("chunk", ctypes.c_int32), ("n_chunks", ctypes.c_int32),
        ("pitch", ctypes.c_int32), ("band", ctypes.c_int32), ("xw_band", ctypes.c_int32),
    ]


# hs2_plan_copy_table selectors (HS2_TAB_*)
(TAB_LINE_ID, TAB_ROWS_LO, TAB_ROWS_DG, TAB_ROWS_HI, TAB_LU, TAB_CHUNK, TAB_GE, TAB_CHUNK_IL, TAB_UTAB, TAB_UCODE, TAB_XW,
 TAB_XW_CODE) = range(12)


class Hs2Error(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); the list is also what tests check against the header
PROTOTYPES = {
    "hs2_abi_version": (ctypes.c_int, []),
    "hs2_last_error": (ctypes.c_char_p, []),
    "hs2_sizeof": (ctypes.c_int, [ctypes.c_int]),
    "hs2_plan_create": (ctypes.c_int, [ctypes.POINTER(PlanDesc), ctypes.POINTER(c_void_p)]),
    "hs2_plan_build": (ctypes.c_int, [ctypes.POINTER(BuildDesc), ctypes.POINTER(c_void_p)]),
    "hs2_plan_axis_info": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.POINTER(AxisInfo)]),
    "hs2_plan_copy_table": (ctypes.c_int64, [c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_int64]),
    "hs2_tables_chunk": (ctypes.c_int, [c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        c_double_p, c_double_p]),
    "hs2_plan_destroy": (ctypes.c_int, [c_void_p]),
    "hs2_plan_launches_per_step": (ctypes.c_int, [c_void_p]),
    "hs2_plan_x_kernel": (ctypes.c_int, [c_void_p]),
    "hs2_plan_last_kernel": (ctypes.c_int, [c_void_p, ctypes.c_int]),
    "hs2_kernel_name": (ctypes.c_char_p, [ctypes.c_int]),
    "hs2_step": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(Source), c_void_p, c_void_p, c_void_p]),
    "hs2_run_steps": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_void_p, ctypes.c_int,
                                     c_void_p, c_void_p, ctypes.c_double, c_void_p, ctypes.c_int, c_void_p]),
    "hs2_observe": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, c_void_p, c_void_p, ctypes.c_double, c_void_p]),
    "hs2_sweep_x": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(Source), c_void_p, c_void_p, c_void_p]),
    "hs2_sweep_x_part": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(Source), c_void_p, c_void_p, ctypes.c_int,
                                        c_void_p]),
    "hs2_sweep_y": (ctypes.c_int, [c_void_p, c_void_p, c_void_p]),
    "hs2_sweep_z": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hs2_sweep_z_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, c_void_p]),
    "hs2_sweep_z_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                            ctypes.c_int64, c_void_p]),
    "hs2_peer_alloc": (ctypes.c_int, [ctypes.c_int64, ctypes.POINTER(c_void_p), c_void_p]),
    "hs2_peer_open": (ctypes.c_int, [c_void_p, ctypes.POINTER(c_void_p)]),
    "hs2_peer_close": (ctypes.c_int, [c_void_p]),
    "hs2_peer_free": (ctypes.c_int, [c_void_p]),
    "hs2_sweep_z_forward_push": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_uint64),
                                                c_void_p]),
    "hs2_sweep_z_forward_push_cols": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                                     ctypes.POINTER(ctypes.c_uint64), c_void_p]),
    "hs2_sweep_z_backward_cols": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int64,
                                                 ctypes.c_int64, c_void_p]),
    "hs2_sweep_z_fused": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.c_double, c_void_p, c_void_p]),
    "hs2_peer_fill_empty": (ctypes.c_int, [c_void_p, ctypes.c_int64, c_void_p]),
    "hs2_sweep_z_fused_tile_lines": (ctypes.c_int, [c_void_p]),
    "hs2_flag_signal": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_uint64, c_void_p]),
    "hs2_flag_wait": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_uint64, ctypes.c_double,
                                     c_void_p, c_void_p]),
    "hs2_copy_async": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int64, c_void_p]),
    "hs2_tridiag_scratch_bytes": (ctypes.c_int64, [ctypes.c_int64]),
    "hs2_tridiag_lu": (ctypes.c_int, [ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hs2_tridiag_solve": (ctypes.c_int, [ctypes.c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}


def lib():
    """The loaded library (cached).  Raises ImportError when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "heatsim2_b200: %s is missing - build it with "
                "`python heatsim2_b200/build.py` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.hs2_abi_version() != ABI_VERSION:
            raise ImportError("heatsim2_b200: ABI version mismatch (%d != %d); rebuild the library"
                              % (L.hs2_abi_version(), ABI_VERSION))
        for which, struct in enumerate((AxisTables, PlanDesc, Source, BuildDesc, AxisInfo)):
            if L.hs2_sizeof(which) != ctypes.sizeof(struct):
                raise ImportError("heatsim2_b200: ctypes mirror of %s is %d bytes, the library's struct %d"
                                  % (struct.__name__, ctypes.sizeof(struct), L.hs2_sizeof(which)))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise Hs2Error("hs2_b200 error %d: %s" % (rc, lib().hs2_last_error().decode("utf-8", "replace")))
