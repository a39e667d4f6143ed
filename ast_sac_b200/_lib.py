"""ctypes loader of the CUDA extension (csrc/libshipenv.so, C ABI of include/shipenv.h).

There is no CPU fallback: if the shared library is missing the import of any op fails loudly, and
without a CUDA device ``shipenv_create`` returns SHIPENV_E_CUDA which is raised as RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AST_SAC_B200_LIB overrides the library path (used to compare build variants, e.g. SENV_MIN_BLOCKS)
LIB_PATH = os.environ.get("AST_SAC_B200_LIB") or os.path.join(_HERE, "csrc", "libshipenv.so")

MAX_WP, MAX_IW, MAX_POLY, MAX_VERT = 32, 30, 16, 128
ABI_VERSION = 9
MATH_STRICT, MATH_FAST = 0, 1
MODEL_SIMPLE, MODEL_DETAILED, MODEL_SIMPLIFIED = 0, 1, 2
ENV_COLAV_NONIW, ENV_COLAV_IW, ENV_RL = 0, 1, 2
COLLAV_NONE, COLLAV_SIMPLE, COLLAV_SBMPC = 0, 1, 2

INFO_EVENT_MASK = 0x7ff
INFO_TERMINAL, INFO_TEST_STOP, INFO_OBS_STOP, INFO_DONE, INFO_UNBOUND = 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20

SF = dict(north=0, east=1, yaw=2, u=3, v=4, r=5, omega=6, time=7, e_ct=8, e_ct_int=9, hdg_err_i=10,
          hdg_prev_err=11, spd_err_i=12, spd_aux=13, seg_alpha=14, seg_sin=15, seg_cos=16)
SF_COUNT = 17
LOG_COLS = ("time", "north", "east", "yaw", "rudder", "u", "v", "r", "omega", "cmd", "e_ct", "e_psi")
EF = dict(travel_dist=0, travel_time=1, acc_reward=2, n_base=3, e_base=4, log_north=5, log_east=6, sb_p_last=7,
          sb_chi_last=8, seg_new_alpha=9, seg_new_sin=10, seg_new_cos=11, seg_end_alpha=12, seg_end_sin=13,
          seg_end_cos=14)
EF_COUNT = 15
EI = dict(sampling_count=0, snapshot_info=1, flags=2)
EI_COUNT = 3

_D = C.c_double
SHIP_PARAM_DOUBLES = [
    "mass", "i_z", "x_du", "y_dv", "n_dr", "lin_damp_u", "lin_damp_v", "lin_damp_r", "ku", "kv", "kr",
    "inv_m_u", "inv_m_v", "inv_m_r", "cur_n", "cur_e", "wind_speed", "wind_dir", "cos_wind_dir", "sin_wind_dir",
    "proj_area_f", "proj_area_l", "l_ship", "c_rudder_v", "c_rudder_r",
    "init_north", "init_east", "init_yaw", "init_u", "init_v", "init_r", "init_omega",
    "dt", "sim_time", "dt_shaft", "spd_kp", "spd_kd", "spd_ki", "max_thrust",
    "kp_ship_speed", "ki_ship_speed", "kp_shaft_speed", "ki_shaft_speed", "max_shaft_speed", "init_shaft_err_i",
    "ctrl_dt", "inv_ctrl_dt", "hdg_kp", "hdg_kd", "hdg_ki", "max_rudder", "los_ra", "los_r", "los_ki", "los_limit",
    "desired_speed", "p_me", "p_el", "tq_me_max", "tq_el_max", "d_me", "d_hsg", "r_me", "r_hsg", "jp",
    "k_torque", "thrust_coeff", "k_thrust", "thrust_tau", "nav_fail_tol",
]


class ShipParams(C.Structure):
    _fields_ = [(n, _D) for n in SHIP_PARAM_DOUBLES] + [
        ("wp_north", _D * MAX_WP), ("wp_east", _D * MAX_WP), ("w_ship", _D),
        ("n_wp", C.c_int32), ("model_kind", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("ship", ShipParams * 2), ("vert_e", _D * MAX_VERT), ("vert_n", _D * MAX_VERT),
                ("map_min_n", _D), ("map_max_n", _D), ("map_min_e", _D), ("map_max_e", _D),
                ("ab_segment_length", _D), ("ab_north_segment_length", _D), ("ab_east_segment_length", _D),
                ("cos_omega", _D), ("sin_omega", _D), ("n_base0", _D), ("e_base0", _D), ("roa", _D),
                ("poly_start", C.c_int32 * (MAX_POLY + 1)), ("n_poly", C.c_int32), ("env_kind", C.c_int32),
                ("collav", C.c_int32), ("max_sampling_frequency", C.c_int32), ("abi_version", C.c_int32),
                ("math_mode", C.c_int32), ("obs_sampled_route", C.c_int32)]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ship_f64", "ship_i32", "env_f64", "env_i32", "iw_f64", "prev_f32",
                                          "obs_f32", "reward", "info_i32", "nsub_i32", "counters")]


class Layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("ship_f64", "ship_i32", "env_f64", "env_i32", "iw_f64", "prev_f32",
                                         "obs_f32", "reward", "info_i32", "nsub_i32", "counters")]


EXPORTS = [
    "shipenv_abi_version", "shipenv_sizeof_params", "shipenv_last_error", "shipenv_create", "shipenv_destroy",
    "shipenv_layout", "shipenv_bind", "shipenv_alloc", "shipenv_buffers", "shipenv_set_params",
    "shipenv_construct", "shipenv_reset", "shipenv_init_step", "shipenv_step", "shipenv_substeps",
    "shipenv_ship_rollout", "shipenv_reset_host", "shipenv_step_host", "shipenv_substeps_host",
    "shipenv_read_counters", "shipenv_measure_fp64_peak", "shipenv_selftest_math", "shipenv_set_trajectory_log",
    "shipenv_time_env_kernel", "shipenv_env_kernel_ms", "shipenv_map_query", "shipenv_map_safe_radius",
    "shipenv_register_host", "shipenv_unregister_host",
]

_lib = None


def load():
    """Load libshipenv.so; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"CUDA extension {LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
            "(or __graft_entry__.build()). ast_sac_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.shipenv_last_error.restype = C.c_char_p
    L.shipenv_create.argtypes = [C.POINTER(Params), i64, i32, C.POINTER(vp)]
    L.shipenv_destroy.argtypes = [vp]
    L.shipenv_layout.argtypes = [vp, C.POINTER(Layout)]
    L.shipenv_bind.argtypes = [vp, C.POINTER(Buffers)]
    L.shipenv_alloc.argtypes = [vp]
    L.shipenv_buffers.argtypes = [vp, C.POINTER(Buffers)]
    L.shipenv_set_params.argtypes = [vp, C.POINTER(Params)]
    L.shipenv_construct.argtypes = [vp, vp, vp]
    L.shipenv_reset.argtypes = [vp, vp, vp, vp]
    L.shipenv_init_step.argtypes = [vp, vp]
    L.shipenv_step.argtypes = [vp, vp, vp]
    L.shipenv_substeps.argtypes = [vp, i32, vp]
    L.shipenv_ship_rollout.argtypes = [vp, i32, vp]
    L.shipenv_reset_host.argtypes = [vp, vp, vp]
    L.shipenv_step_host.argtypes = [vp, vp, vp, vp, vp, vp]
    L.shipenv_substeps_host.argtypes = [vp, i32, vp, vp, vp, vp]
    L.shipenv_read_counters.argtypes = [vp, vp]
    L.shipenv_set_trajectory_log.argtypes = [vp, vp, vp, i64, i64]
    L.shipenv_time_env_kernel.argtypes = [vp, i32]
    L.shipenv_env_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.shipenv_measure_fp64_peak.argtypes = [i32, i32, C.POINTER(C.c_double)]
    L.shipenv_selftest_math.argtypes = [i32, i64, C.c_uint64, vp]
    L.shipenv_register_host.argtypes = [vp, vp, C.c_size_t]
    L.shipenv_unregister_host.argtypes = [vp, vp]
    L.shipenv_map_query.argtypes = [vp, i64, vp, vp, C.c_double, vp, vp, vp, vp]
    L.shipenv_map_safe_radius.argtypes = [vp, i64, vp, vp, vp, vp]
    if L.shipenv_abi_version() != ABI_VERSION:
        raise ImportError("libshipenv.so ABI version mismatch; rebuild the extension")
    if L.shipenv_sizeof_params() != C.sizeof(Params):
        raise ImportError("ShipEnvParams layout mismatch between include/shipenv.h and ast_sac_b200/_lib.py")
    _lib = L
    return L


def measure_fp64_peak(device: int = 0, repeats: int = 5) -> float:
    """FP64 FMA peak of the device in TFLOP/s (DFMA microbenchmark, roofline denominator)."""
    out = C.c_double()
    check(load().shipenv_measure_fp64_peak(device, repeats, C.byref(out)))
    return out.value


def selftest_math(device: int = 0, n: int = 1 << 24, seed: int = 1):
    """(sincos, atan, sqrt, division, exp, atan2, fmod, x*rsqrt(x) beyond 2 ulp, a*rsqrt(x) beyond 2 ulp) mismatch counts
    of csrc/shipenv_math.cuh vs the CUDA math library: fast build, then strict build (18 numbers, all expected 0)."""
    out = (C.c_ulonglong * 18)()
    check(load().shipenv_selftest_math(device, n, seed, out))
    return list(out)


def check(rc: int):
    if rc != 0:
        msg = load().shipenv_last_error().decode("utf-8", "replace")
        exc = {1: ValueError, 2: RuntimeError, 3: RuntimeError, 4: MemoryError}.get(rc, RuntimeError)
        raise exc(f"shipenv error {rc}: {msg}")
