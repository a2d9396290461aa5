"""ctypes binding of libpyotf_b200.so (the C ABI declared in include/pyotf_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails the
error propagates.  Build it with ``python -m pyotf_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpyotf_b200.so")


class HanserParams(C.Structure):
    _fields_ = [
        ("wl", C.c_double), ("na", C.c_double), ("ni", C.c_double), ("res", C.c_double),
        ("size", C.c_int), ("vec_corr", C.c_int), ("condition", C.c_int), ("dtype", C.c_int),
        ("support", C.c_int),
    ]


class SheppardParams(C.Structure):
    _fields_ = [
        ("wl", C.c_double), ("na", C.c_double), ("ni", C.c_double), ("res", C.c_double), ("zres", C.c_double),
        ("size", C.c_int), ("zsize", C.c_int), ("vec_corr", C.c_int), ("condition", C.c_int), ("dual", C.c_int),
        ("dtype", C.c_int),
    ]


class Pyotf_b200Error(RuntimeError):
    """Raised when a C-ABI entry point reports an error."""


_lib = None

_vp, _i, _ll, _ip, _d = C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_int), C.c_double

# name -> (restype, argtypes); every symbol include/pyotf_b200.h declares
PROTOTYPES = {
    "pyotf_b200_abi_version": (_i, []),
    "pyotf_b200_version": (_i, []),
    "pyotf_b200_last_error": (C.c_char_p, []),
    "pyotf_b200_device_query": (_i, [_i, _ip, _ip, _ip, _ip]),
    "pyotf_b200_hanser_create": (_i, [C.POINTER(HanserParams), _vp, C.POINTER(_vp)]),
    "pyotf_b200_hanser_destroy": (_i, [_vp]),
    "pyotf_b200_hanser_set_overlap": (_i, [_vp, _i]),
    "pyotf_b200_hanser_info": (_i, [_vp, _ip, _ip, _ip, _ip]),
    "pyotf_b200_hanser_tables": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "pyotf_b200_pupil_support": (_i, [_vp, _i, _ip, _vp]),
    "pyotf_b200_hanser_pack_pupil": (_i, [_vp, _vp, _i, _vp, _vp]),
    "pyotf_b200_hanser_psf": (_i, [_vp, _vp, _i, _vp, _i, _ll, _vp, _vp, _vp]),
    "pyotf_b200_hanser_psf_zhost": (_i, [_vp, _vp, _i, _vp, _i, _ll, _vp, _vp, _vp]),
    "pyotf_b200_hanser_kspace": (_i, [_d, _d, _d, _i, _vp, _vp, _vp, _vp]),
    "pyotf_b200_hanser_defocus": (_i, [_vp, _ll, _vp, _i, _vp, _vp]),
    "pyotf_b200_zernike": (_i, [_vp, _vp, _ll, _vp, _vp, _i, _i, _vp, _vp]),
    "pyotf_b200_zernike_gram": (_i, [_vp, _vp, _ll, _i, _i, _vp, _vp]),
    "pyotf_b200_fftn": (_i, [_vp, _vp, _i, C.POINTER(_ll), _ip, _i, _i, _i, _i, _i, _i, _vp]),
    "pyotf_b200_abs2_sum0": (_i, [_vp, _i, _ll, _vp, _i, _vp]),
    "pyotf_b200_fft3d_shifted": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pyotf_b200_bin_lateral": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    "pyotf_b200_disk_conv_mul": (_i, [_vp, _vp, _i, _i, _d, _vp, _i, _vp]),
    "pyotf_b200_wiener_otf": (_i, [_vp, _ll, _d, _vp, _i, _vp]),
    "pyotf_b200_sim_psf": (_i, [_vp, _i, _i, _i, _d, _d, _d, _d, _d, C.POINTER(_d), _i, _i, _i, _i, _vp, _i, _vp]),
    "pyotf_b200_hanser_aberration_pupil": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp]),
    "pyotf_b200_sheppard_otfa": (_i, [C.POINTER(SheppardParams), _vp, _vp, _vp]),
    "pyotf_b200_sheppard_psf": (_i, [C.POINTER(SheppardParams), _vp, _vp, _vp, _vp]),
    "pyotf_b200_prep_data_u16": (_i, [_vp, _i, _i, _i, _i, _d, _i, _vp, _ip, _vp]),
    "pyotf_b200_pr_create": (_i, [C.POINTER(HanserParams), _vp, _i, _ll, _vp, _i, _i, _vp, C.POINTER(_vp)]),
    "pyotf_b200_pr_destroy": (_i, [_vp]),
    "pyotf_b200_pr_info": (_i, [_vp, _ip, _ip, _ip, _ip, _ip]),
    "pyotf_b200_pr_kernels_per_iteration": (_i, [_vp, _i, _ip]),
    "pyotf_b200_pr_debug_buffers": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "pyotf_b200_pr_set_pupil": (_i, [_vp, _vp, _vp]),
    "pyotf_b200_pr_get_pupil": (_i, [_vp, _i, _vp, _vp]),
    "pyotf_b200_pr_run": (_i, [_vp, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _ip, _vp]),
    "pyotf_b200_pr_iterate": (_i, [_vp, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _ip, _vp]),
    "pyotf_b200_pr_window_alloc": (_i, [_vp, _i, _vp]),
    "pyotf_b200_pr_window_open": (_i, [_vp, _vp, _i, _i]),
    "pyotf_b200_unwrap_phase_2d": (_i, [_vp, _vp, _i, _i]),
    "pyotf_b200_comm_load": (_i, [C.c_char_p]),
    "pyotf_b200_comm_unique_id": (_i, [_vp]),
    "pyotf_b200_comm_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "pyotf_b200_comm_destroy": (_i, [_vp]),
    "pyotf_b200_comm_allreduce_f64": (_i, [_vp, _vp, _ll, _vp]),
}


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Pyotf_b200Error(
            f"{LIB_PATH} is missing: the CUDA extension is not built "
            "(run `python -m pyotf_b200.build`); pyotf_b200 has no CPU fallback"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pyotf_b200_last_error()
        raise Pyotf_b200Error((msg or b"unknown error").decode())
