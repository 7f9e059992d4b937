"""ctypes binding of the product library libkl_shell.so (C ABI: include/kl_shell.h).

The library is built in-tree by build.py / __graft_entry__.build().  There is NO fallback: if the
shared object is missing, or no CUDA device is present, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

from .problem import kl_problem, kl_bc, kl_interface, c_double_p, c_int_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KL_LIB") or os.path.join(_HERE, "libkl_shell.so")   # KL_LIB: A/B builds of the same sources

KL_ERRORS = {0: "KL_OK", -1: "KL_E_ARG", -2: "KL_E_CUDA", -3: "KL_E_NONFINITE", -4: "KL_E_JACOBIAN", -5: "KL_E_C33",
             -6: "KL_E_NOGPU"}

# every symbol include/kl_shell.h declares
SYMBOLS = ["kl_build_dofmap", "kl_create", "kl_destroy", "kl_sizes", "kl_pattern_host", "kl_pattern_device",
           "kl_jacobian", "kl_residual", "kl_al_residual", "kl_force", "kl_jacobian_device", "kl_residual_device",
           "kl_check", "kl_values_device", "kl_set_strip", "kl_last_timing", "kl_last_error", "kl_kernel_launches",
           "kl_jacobian_kernel_ms", "kl_points_kernel_ms", "kl_measure_fp64_peak", "kl_mass",
           "kl_assemble_device", "kl_cg_solve", "kl_cg_solve_device", "kl_spmv", "kl_cg_last_timing", "kl_newton_solve",
           "kl_stress_dim", "kl_eval_stress", "kl_principal_stretches", "kl_boundary_force",
           "kl_pin_values", "kl_unpin_values", "kl_fetch_values", "kl_set_values", "kl_pattern_lower_host", "kl_jacobian_lower",
           "kl_al_residual_device", "kl_alm_step", "kl_strip_begin_device", "kl_jacobian_rows_device",
           "kl_mp_build_dofmap", "kl_mp_create", "kl_mp_destroy", "kl_mp_context", "kl_mp_patch", "kl_mp_num_patches",
           "kl_mp_set_active", "kl_mp_interface_dofs", "kl_stability"]

# stress_type of constructStress (include/kl_shell.h)
STRESS_TYPES = {"displacement": 0, "membrane_force": 1, "flexural_moment": 2, "membrane": 3, "flexural": 4,
                "membrane_strain": 5, "flexural_strain": 6, "principal_stretch": 7, "principal_stretch_dir": 8,
                "principal_stress_membrane": 9, "principal_stress_flexural": 10, "principal_membrane_strain": 11,
                "principal_flexural_strain": 12, "von_mises_membrane": 13, "tension_field": 14}


class kl_newton_options(C.Structure):
    _fields_ = [("tolU", C.c_double), ("tolF", C.c_double), ("relaxation", C.c_double), ("max_it", C.c_int32),
                ("linear_start", C.c_int32), ("cg_tol", C.c_double), ("cg_max_iter", C.c_int32)]


class kl_newton_info(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("cg_iterations", C.c_int64),
                ("residual", C.c_double), ("residual_ini", C.c_double), ("dU_norm", C.c_double), ("DU_norm", C.c_double),
                ("ms_assembly", C.c_float), ("ms_solve", C.c_float)]


class kl_alm_options(C.Structure):
    _fields_ = [("tolU", C.c_double), ("tolF", C.c_double), ("max_it", C.c_int32), ("phi", C.c_double), ("relaxation", C.c_double),
                ("cg_tol", C.c_double), ("cg_max_iter", C.c_int32)]


class kl_alm_info(C.Structure):
    _fields_ = [("status", C.c_int32), ("iterations", C.c_int32), ("cg_iterations", C.c_int64), ("residueF", C.c_double),
                ("residueU", C.c_double), ("phi", C.c_double), ("DeltaL", C.c_double), ("ms_assembly", C.c_float), ("ms_solve", C.c_float)]


_LIB = None


class KLError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__(f"{KL_ERRORS.get(rc, rc)}: {msg}")
        self.rc = rc


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m gsstructuralanalysis_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.kl_build_dofmap.argtypes = [C.c_int32, C.c_int32, C.POINTER(kl_bc), c_int_p, c_int_p, c_int_p]
    L.kl_create.argtypes = [C.POINTER(kl_problem), C.c_int, C.POINTER(vp)]
    L.kl_destroy.argtypes = [vp]
    L.kl_destroy.restype = None
    L.kl_sizes.argtypes = [vp, c_int_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.kl_pattern_host.argtypes = [vp, c_int_p, c_int_p]
    L.kl_pattern_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.kl_jacobian.argtypes = [vp, c_double_p, c_double_p]
    L.kl_jacobian_lower.argtypes = [vp, c_double_p, c_double_p]
    L.kl_pattern_lower_host.argtypes = [vp, c_int_p, c_int_p, C.POINTER(C.c_int64)]
    L.kl_pin_values.argtypes = [vp, c_double_p, C.c_int64]
    L.kl_unpin_values.argtypes = [vp, c_double_p]
    L.kl_fetch_values.argtypes = [vp, c_double_p]
    L.kl_set_values.argtypes = [vp, c_double_p]
    L.kl_residual.argtypes = [vp, c_double_p, c_double_p]
    L.kl_al_residual.argtypes = [vp, c_double_p, C.c_double, c_double_p]
    L.kl_force.argtypes = [vp, c_double_p]
    L.kl_mass.argtypes = [vp, C.c_double, c_double_p, c_double_p]
    L.kl_jacobian_device.argtypes = [vp, vp, vp]
    L.kl_residual_device.argtypes = [vp, vp, C.c_double, C.c_double, vp, vp]
    L.kl_assemble_device.argtypes = [vp, vp, C.c_double, C.c_double, vp, vp]
    L.kl_check.argtypes = [vp, vp]
    L.kl_values_device.argtypes = [vp]
    L.kl_values_device.restype = vp
    L.kl_set_strip.argtypes = [vp, C.c_int32, C.c_int32]
    L.kl_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.kl_last_error.restype = C.c_char_p
    L.kl_kernel_launches.argtypes = [vp]
    L.kl_jacobian_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.kl_points_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.kl_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.kl_cg_solve.argtypes = [vp, c_double_p, c_double_p, C.c_double, C.c_int32, c_int_p, c_double_p]
    L.kl_cg_solve_device.argtypes = [vp, vp, vp, C.c_double, C.c_int32, c_int_p, c_double_p, vp]
    L.kl_spmv.argtypes = [vp, c_double_p, c_double_p]
    L.kl_cg_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.kl_newton_solve.argtypes = [vp, c_double_p, C.POINTER(kl_newton_options), C.POINTER(kl_newton_info)]
    L.kl_alm_step.argtypes = [vp, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, C.POINTER(kl_alm_options), C.POINTER(kl_alm_info)]
    L.kl_al_residual_device.argtypes = [vp, vp, C.c_double, vp, vp]
    L.kl_strip_begin_device.argtypes = [vp, vp, C.c_double, C.c_double, vp, C.c_int32, vp]
    L.kl_jacobian_rows_device.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.kl_stress_dim.argtypes = [C.c_int32]
    L.kl_eval_stress.argtypes = [vp, c_double_p, C.c_int32, C.c_int32, c_double_p, C.c_double, c_double_p]
    L.kl_principal_stretches.argtypes = [vp, c_double_p, C.c_int32, c_double_p, C.c_double, c_double_p]
    L.kl_boundary_force.argtypes = [vp, c_double_p, C.c_int32, c_double_p]
    L.kl_stability.argtypes = [vp, c_double_p, c_int_p, c_double_p]
    L.kl_mp_build_dofmap.argtypes = [C.c_int32, c_int_p, c_int_p, C.POINTER(kl_bc), C.c_int32, C.POINTER(kl_interface), c_int_p, c_int_p, c_int_p]
    L.kl_mp_create.argtypes = [C.c_int32, C.POINTER(kl_problem), C.c_int, C.POINTER(vp)]
    L.kl_mp_destroy.argtypes = [vp]
    L.kl_mp_destroy.restype = None
    L.kl_mp_context.argtypes = [vp]
    L.kl_mp_context.restype = vp
    L.kl_mp_patch.argtypes = [vp, C.c_int32]
    L.kl_mp_patch.restype = vp
    L.kl_mp_num_patches.argtypes = [vp]
    L.kl_mp_set_active.argtypes = [vp, c_int_p]
    L.kl_mp_interface_dofs.argtypes = [vp, c_int_p, c_int_p]
    _LIB = L
    return L


def check(rc):
    if rc != 0:
        raise KLError(rc, lib().kl_last_error().decode())
