"""ctypes view of libsfmshim.so: flat wrappers around the C++ drop-in shim (host/sfmgpu_shim.hpp)."""
import ctypes as C
import os

import numpy as np

import sfmgpu

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsfmshim.so")
_u8 = np.ctypeslib.ndpointer(np.uint8, flags="C")
_f8 = np.ctypeslib.ndpointer(np.float64, flags="C")
_i4 = np.ctypeslib.ndpointer(np.int32, flags="C")
_i, _d, _vp = C.c_int, C.c_double, C.c_void_p


def load():
    if not os.path.exists(PATH):
        sfmgpu.build_library()
    lib = C.CDLL(PATH)
    lib.shim_last_error.restype = C.c_char_p
    lib.shim_build_pyr.argtypes = [_u8, _i, _i, _i, _u8]
    lib.shim_shi_tomasi.argtypes = [_u8, _i, _i, _i, _d, _i, _f8, _i]
    lib.shim_pair_frontend.argtypes = [_u8, _u8, _i, _i, _i, _d, _i, _i, _i, _i, _d, _i, _f8, _f8, C.POINTER(_i)]
    lib.shim_tracker_create.restype = _vp
    lib.shim_tracker_create.argtypes = [_i, _i, _d, _i, _i, _i, _i, _d]
    lib.shim_tracker_destroy.argtypes = [_vp]
    lib.shim_tracker_step.argtypes = [_vp, _u8, _i, _i, _f8, _f8, _i4, _i]
    lib.shim_tracker_tracks.argtypes = [_vp, _f8, _i4, _i]
    lib.shim_multitracker_create.restype = _vp
    lib.shim_multitracker_create.argtypes = [_i, _i, _i, _i, _i]
    lib.shim_multitracker_destroy.argtypes = [_vp]
    lib.shim_multitracker_step.argtypes = [_vp, _u8, _i, _i, _i, _f8, _f8, _i4, _i4, _i, _f8, _i4, _i4]
    lib.shim_find_E_ransac.argtypes = [_f8, _f8, _f8, _i, _i, _d, _i, _f8, _f8, _i4, C.POINTER(_i)]
    lib.shim_host_triangulate.argtypes = [_f8, _f8, _i4, _i4, _f8, _f8, _i, _f8]
    lib.shim_set_device_solver.argtypes = [_i]
    lib.shim_set_device_solver.restype = None
    lib.shim_set_speculate.argtypes = [_i]
    lib.shim_set_speculate.restype = None
    lib.shim_host_norm_points.argtypes = [_f8, _f8, _i, _f8]
    lib.shim_host_hypotheses.argtypes = [_f8, _f8, _i, _i, _f8]
    lib.shim_host_recover_pose.argtypes = [_f8, _f8, _f8, _i4, _i, _f8, _f8]
    return lib


def ck(lib, rc):
    if rc <= -1000:
        raise RuntimeError(lib.shim_last_error().decode())
    return rc
