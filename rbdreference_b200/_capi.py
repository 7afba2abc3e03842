"""ctypes binding of librbd_b200.so (include/rbd_b200.h).  No compute happens in Python."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "librbd_b200.so")

PASS_SYMBOLS = ["rnea_fpass", "rnea_bpass", "rnea_grad_fpass_dq", "rnea_grad_fpass_dqd",
                "rnea_grad_bpass_dq", "rnea_grad_bpass_dqd", "minv_bpass", "minv_fpass"]
FUSED_SYMBOLS = ["rnea", "rnea_grad", "minv", "forward_dynamics", "forward_dynamics_grad", "crba", "aba"]
PLAIN_SYMBOLS = ["rbd_abi_version", "rbd_last_error_string", "rbd_model_create", "rbd_model_destroy",
                 "rbd_model_num_dof", "rbd_model_uses_world_kernels", "rbd_set_kernel_variant",
                 "rbd_model_set_kernel_variant",
                 "rbd_measure_fma_peak", "rbd_launch_count", "rbd_trim_scratch", "rbd_prepare_device",
                 "rbd_ee_model_create", "rbd_ee_model_destroy", "rbd_ee_model_num_ee",
                 "rbd_fb_model_create", "rbd_fb_model_destroy", "rbd_fb_model_num_vel",
                 "rbd_fb_model_set_kernel_variant"]
FB_SYMBOLS = ["fb_rnea", "fb_rnea_grad", "fb_minv", "fb_forward_dynamics", "fb_forward_dynamics_grad"] + ["fb_" + p for p in PASS_SYMBOLS]
EE_SYMBOLS = ["end_effector_pose", "end_effector_pose_gradient"]


def exported_symbols():
    """Every symbol include/rbd_b200.h declares."""
    out = list(PLAIN_SYMBOLS)
    for base in FUSED_SYMBOLS + PASS_SYMBOLS + EE_SYMBOLS + FB_SYMBOLS:
        out += ["rbd_%s_f64" % base, "rbd_%s_f32" % base]
    return out


class RbdModelDesc(ctypes.Structure):
    _fields_ = [("n", c_int32),
                ("parent", POINTER(c_int32)), ("kind", POINTER(c_int32)),
                ("S", POINTER(c_double)), ("XA", POINTER(c_double)), ("XB", POINTER(c_double)),
                ("XC", POINTER(c_double)), ("I", POINTER(c_double)), ("damping", POINTER(c_double))]


class RbdEeDesc(ctypes.Structure):
    _fields_ = [("n", c_int32), ("parent", POINTER(c_int32)), ("kind", POINTER(c_int32)),
                ("TA", POINTER(c_double)), ("TB", POINTER(c_double)), ("TC", POINTER(c_double)),
                ("DA", POINTER(c_double)), ("DB", POINTER(c_double)), ("DC", POINTER(c_double)),
                ("n_ee", c_int32), ("ee_joint", POINTER(c_int32)), ("ee_final", POINTER(c_double)),
                ("offset", c_double * 4)]


class RbdFbModelDesc(ctypes.Structure):
    _fields_ = [("bodies", RbdModelDesc), ("pos_off", c_int32), ("quat_off", c_int32), ("w_first", c_int32),
                ("transpose", c_int32)]


class RbdError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load the CUDA library; raise loudly if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "rbdreference_b200: %s is missing - build it with `python -m rbdreference_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.rbd_abi_version.restype = c_int
    lib.rbd_last_error_string.restype = c_char_p
    lib.rbd_model_create.argtypes = [POINTER(RbdModelDesc), POINTER(c_void_p)]
    lib.rbd_model_destroy.argtypes = [c_void_p]
    lib.rbd_model_num_dof.argtypes = [c_void_p]
    lib.rbd_model_uses_world_kernels.argtypes = [c_void_p]
    lib.rbd_set_kernel_variant.argtypes = [c_int]
    lib.rbd_model_set_kernel_variant.argtypes = [c_void_p, c_int]
    lib.rbd_measure_fma_peak.argtypes = [c_int, POINTER(c_double), POINTER(c_double), c_void_p]
    lib.rbd_launch_count.restype = c_int64
    lib.rbd_trim_scratch.argtypes = [c_int64]
    lib.rbd_prepare_device.argtypes = [c_int]
    lib.rbd_ee_model_create.argtypes = [POINTER(RbdEeDesc), POINTER(c_void_p)]
    lib.rbd_ee_model_destroy.argtypes = [c_void_p]
    lib.rbd_ee_model_num_ee.argtypes = [c_void_p]
    lib.rbd_fb_model_create.argtypes = [POINTER(RbdFbModelDesc), POINTER(c_void_p)]
    lib.rbd_fb_model_destroy.argtypes = [c_void_p]
    lib.rbd_fb_model_num_vel.argtypes = [c_void_p]
    lib.rbd_fb_model_set_kernel_variant.argtypes = [c_void_p, c_int]
    P = c_void_p
    for suf, real in (("f64", c_double), ("f32", c_float)):
        sig = {
            "rnea": [P, c_int64, P, P, P, real, P, P, P, P, P],
            "rnea_grad": [P, c_int64, P, P, P, real, c_int, P, P, P],
            "minv": [P, c_int64, P, c_int, P, P],
            "rnea_fpass": [P, c_int64, P, P, P, real, P, P, P, P],
            "rnea_bpass": [P, c_int64, P, P, P, P],
            "rnea_grad_fpass_dq": [P, c_int64, P, P, P, P, real, P, P, P, P],
            "rnea_grad_fpass_dqd": [P, c_int64, P, P, P, P, P, P, P],
            "rnea_grad_bpass_dq": [P, c_int64, P, P, P, P, P],
            "rnea_grad_bpass_dqd": [P, c_int64, P, P, c_int, P, P],
            "minv_bpass": [P, c_int64, P, P, P, P, P, P],
            "minv_fpass": [P, c_int64, P, P, P, P, P, P],
            "forward_dynamics": [P, c_int64, P, P, P, P, P, P],
            "forward_dynamics_grad": [P, c_int64, P, P, P, P, P, P, P],
            "crba": [P, c_int64, P, P, P],
            "aba": [P, c_int64, P, P, P, real, P, P],
            "fb_rnea": [P, c_int64, P, P, P, real, P, P, P, P, P],
            "fb_rnea_grad": [P, c_int64, P, P, P, real, c_int, P, P, P],
            "fb_minv": [P, c_int64, P, c_int, P, P],
            "fb_forward_dynamics": [P, c_int64, P, P, P, P, P, P],
            "fb_forward_dynamics_grad": [P, c_int64, P, P, P, P, P, P, P],
            "end_effector_pose": [P, c_int64, P, P, P],
            "end_effector_pose_gradient": [P, c_int64, P, P, P, P],
        }
        for base in PASS_SYMBOLS:                       # the floating-base helpers share the fixed-base signatures
            sig["fb_" + base] = sig[base]
        for base, argtypes in sig.items():
            fn = getattr(lib, "rbd_%s_%s" % (base, suf))
            fn.argtypes = argtypes
            fn.restype = c_int
    if lib.rbd_abi_version() != 1:
        raise ImportError("rbdreference_b200: ABI version mismatch, rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().rbd_last_error_string()
        raise RbdError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_double))


def _iptr(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_int32))


class ModelHandle:
    """Owns an rbd_model_t* created from a compiled RobotModel."""

    def __init__(self, model):
        lib = load_library()
        self._keep = dict(
            parent=np.ascontiguousarray(model.parent, dtype=np.int32),
            kind=np.ascontiguousarray(model.kind, dtype=np.int32),
            S=np.ascontiguousarray(model.S, dtype=np.float64),
            XA=np.ascontiguousarray(model.XA, dtype=np.float64),
            XB=np.ascontiguousarray(model.XB, dtype=np.float64),
            XC=np.ascontiguousarray(model.XC, dtype=np.float64),
            I=np.ascontiguousarray(model.I, dtype=np.float64),
            damping=np.ascontiguousarray(model.damping, dtype=np.float64),
        )
        k = self._keep
        desc = RbdModelDesc(model.n, _iptr(k["parent"]), _iptr(k["kind"]), _dptr(k["S"]), _dptr(k["XA"]),
                            _dptr(k["XB"]), _dptr(k["XC"]), _dptr(k["I"]), _dptr(k["damping"]))
        handle = c_void_p()
        check(lib.rbd_model_create(ctypes.byref(desc), ctypes.byref(handle)), "rbd_model_create")
        self.ptr = handle
        self._lib = lib

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self._lib.rbd_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class EeModelHandle:
    """Owns an rbd_ee_model_t* created from a compiled EeModel (model.compile_ee_model)."""

    def __init__(self, ee):
        lib = load_library()
        k = self._keep = {name: np.ascontiguousarray(getattr(ee, name), dtype=np.float64)
                          for name in ("TA", "TB", "TC", "DA", "DB", "DC", "ee_final")}
        k["parent"] = np.ascontiguousarray(ee.parent, dtype=np.int32)
        k["kind"] = np.ascontiguousarray(ee.kind, dtype=np.int32)
        k["ee_joint"] = np.ascontiguousarray(ee.ee_joint, dtype=np.int32)
        desc = RbdEeDesc(ee.n, _iptr(k["parent"]), _iptr(k["kind"]), _dptr(k["TA"]), _dptr(k["TB"]), _dptr(k["TC"]),
                         _dptr(k["DA"]), _dptr(k["DB"]), _dptr(k["DC"]), ee.n_ee, _iptr(k["ee_joint"]),
                         _dptr(k["ee_final"]), (c_double * 4)(*[float(x) for x in ee.offset]))
        handle = c_void_p()
        check(lib.rbd_ee_model_create(ctypes.byref(desc), ctypes.byref(handle)), "rbd_ee_model_create")
        self.ptr = handle
        self.n_ee = ee.n_ee
        self._lib = lib

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self._lib.rbd_ee_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class FbModelHandle:
    """Owns an rbd_fb_model_t* created from a compiled FbModel (model.compile_fb_model)."""

    def __init__(self, fb):
        lib = load_library()
        k = self._keep = {name: np.ascontiguousarray(getattr(fb, name), dtype=np.float64)
                          for name in ("S", "XA", "XB", "XC", "I", "damping")}
        k["parent"] = np.ascontiguousarray(fb.parent, dtype=np.int32)
        k["kind"] = np.ascontiguousarray(fb.kind, dtype=np.int32)
        bodies = RbdModelDesc(fb.NB, _iptr(k["parent"]), _iptr(k["kind"]), _dptr(k["S"]), _dptr(k["XA"]),
                              _dptr(k["XB"]), _dptr(k["XC"]), _dptr(k["I"]), _dptr(k["damping"]))
        desc = RbdFbModelDesc(bodies, fb.pos_off, fb.quat_off, fb.w_first, fb.transpose)
        handle = c_void_p()
        check(lib.rbd_fb_model_create(ctypes.byref(desc), ctypes.byref(handle)), "rbd_fb_model_create")
        self.ptr = handle
        self._lib = lib

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self._lib.rbd_fb_model_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass
