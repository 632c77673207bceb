"""ctypes binding of librmab_b200.so (the C ABI declared in include/rmab.h).

There is NO CPU fallback: if the shared library is missing this module raises at first use, and every
entry point refuses non-CUDA tensors.  Build the library with `python -m robustrmab_b200.build`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RMAB_LIB_PATH") or os.path.join(HERE, "librmab_b200.so")   # override: A/B against another build

OK, E_BADARG, E_SHAPE, E_ARCH, E_LAUNCH = 0, -1, -2, -3, -4
DOMAIN_CE, DOMAIN_ARMMAN, DOMAIN_SIS = 0, 1, 2
STREAM_ENV, STREAM_ACT, STREAM_NATURE, STREAM_RESET, STREAM_CF, STREAM_INIT = 0, 1, 2, 3, 4, 7
STOP_FULL, STOP_FAST, STOP_ABS = 0, 1, 2


class RmabError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"librmab_b200 error {code}: {msg}")
        self.code = code


class RmabRng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("stream", C.c_uint32), ("t", C.c_uint32),
                ("env0", C.c_uint32), ("arm0", C.c_uint32), ("t_base", C.c_void_p)]


class RmabNetDesc(C.Structure):
    _fields_ = [("n_arms", C.c_int32), ("n_envs", C.c_int32), ("n_states", C.c_int32), ("n_actions", C.c_int32),
                ("hidden", C.c_int32), ("one_hot", C.c_int32), ("state_norm", C.c_float), ("n_extra", C.c_int32)]


class RmabHyper(C.Structure):
    _fields_ = [("clip_ratio", C.c_double), ("entropy_coeff", C.c_double), ("pi_lr", C.c_double),
                ("vf_lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("adam_eps", C.c_double),
                ("pi_iters", C.c_int32), ("v_iters", C.c_int32), ("adam_step0_pi", C.c_int32),
                ("adam_step0_v", C.c_int32)]


_P = C.c_void_p
_I = C.c_int
_D = C.c_double
_F = C.c_float

# name -> argtypes; every function returns int unless listed in _RESTYPES.  Mirrors include/rmab.h.
SIGNATURES = {
    "rmab_abi_version": [],
    "rmab_last_error": [],
    "rmab_launch_count": [],
    "rmab_reset_states": [_I, _I, _I, C.POINTER(RmabRng), _P, _P, _P],
    "rmab_env_step_table": [_I, _I, _I, _P, _P, _P, _I, _P, _P, C.POINTER(RmabRng), _P, _P, _P, _P],
    "rmab_env_step_sis": [_I, _I, _I, _P, _P, _P, _I, _P, C.POINTER(RmabRng), _P, _P, _P, _P],
    "rmab_sis_table": [_I, _I, _P, _P, _P],
    "rmab_bound_nature": [C.c_int64, _P, _P, _P, _P, _P],
    "rmab_ac_step": [C.POINTER(RmabNetDesc), _P, _P, _P, _P, _P, C.POINTER(RmabRng), _P, _P, _P, _P, _P, _P, _P],
    "rmab_lambda_forward": [_I, _I, _P, _I, _I, _P, _F, _P, _P, _P, _P, C.c_size_t, _P],
    "rmab_lambda_ws_bytes": [_I, _I],
    "rmab_lambda_sgd": [_I, _I, _P, _I, _I, _P, _F, _P, _P, _F, _P, _P, _P],
    "rmab_rollout_table": [_I, C.POINTER(RmabNetDesc), _I, _P, _P, _P, _P, _P, C.c_uint64, C.c_uint32, C.c_uint32,
                           C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P],
    "rmab_gaussian_sample": [_I, _I, _P, _P, C.POINTER(RmabRng), _P, _P, _P, _P],
    "rmab_env_counterfactual": [_I, _I, _I, _I, _P, _P, _P, _I, _P, _P, C.POINTER(RmabRng), _I, _P, _P],
    "rmab_critic_extra_iter": [C.POINTER(RmabNetDesc), C.POINTER(RmabHyper), _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P],
    "rmab_adam_step": [C.c_int64, _P, _P, _P, _P, _D, _D, _D, _D, _I, _P],
    "rmab_gae": [_I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _D, _D, _I, _P, _P, _P, _P, _P],
    "rmab_gae_explicit": [_I, _I, _I, _P, _P, _P, _P, _P, _D, _D, _I, _P, _P, _P, _P, _P],
    "rmab_cost_togo": [_I, _I, _I, _I, _P, _P, _D, _P, _P, _P],
    "rmab_ppo_update": [C.POINTER(RmabNetDesc), C.POINTER(RmabHyper), _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                        _P, _P, _P, _P],
    "rmab_stat_rows": [_I],
    "rmab_epoch_ws_bytes": [_I, _I],
    "rmab_epoch_table": [_I, C.POINTER(RmabNetDesc), _I, _P, _P, _P, _P, _P, _P, _P, _D, _D, C.c_uint64, C.c_uint32,
                         C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P, C.c_size_t, _P, _P, _P, _P],
    "rmab_epoch_table_envshard": [_I, C.POINTER(RmabNetDesc), _I, _P, _P, _P, _P, _P, _P, _P, _D, _D, C.c_uint64, C.c_uint32,
                                  C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P, _P, _P, _P, C.c_size_t, _P, _P, _P, _P],
    "rmab_ppo_update_table": [C.POINTER(RmabNetDesc), C.POINTER(RmabHyper), _I, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                              _P],
    "rmab_nat_ws_bytes": [_I, _I, _I, _I],
    "rmab_nat_l1_fwd": [_I, _I, _I, _I, _P, C.c_int64, _F, _P, _P, C.c_int64, _P, _P, _P, C.c_size_t, _P],
    "rmab_nat_mid_fwd": [_I, _I, _P, _P, _P, _P, _P, _P, _P],
    "rmab_nat_mu": [_I, _I, _I, _P, _P, _P, _P, C.c_int64, _P],
    "rmab_nat_logp": [_I, _I, _I, _P, _P, _P, _P, _P, C.c_int64, _P, _P, C.c_size_t, _P],
    "rmab_nat_pi_coef": [_I, _P, _P, _P, _F, _P, _P],
    "rmab_nat_out_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, C.c_int64, _P, _P, _P, _P, _P, _P, C.c_size_t, _P],
    "rmab_nat_mid_bwd": [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P],
    "rmab_nat_l1_bwd": [_I, _I, _I, _I, _P, C.c_int64, _F, _P, C.c_int64, _P, _P, _P, _P, C.c_size_t, _P],
    "rmab_nat_v_out": [_I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P],
    "rmab_split16_rows": [_I, _I, _P, C.c_int64, _I, _P, _P, _P, _P],
    "rmab_split16_cols": [_I, _I, _P, C.c_int64, _I, _P, _P, _P, _P, _P],
    "rmab_gemm3h": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _I, _P],
    "rmab_gemm3h_adam": [_I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _D, _D, _D, _D, _I, _I, _P],
    "rmab_select_ws_bytes": [_I, _I, _I],
    "rmab_budget_select_binary": [_I, _I, _P, _F, _F, _I, _P, _P, C.c_size_t, _P],
    "rmab_select_hist": [_I, _I, _P, _I, _P, C.c_size_t, _P, _P],
    "rmab_select_scan": [_I, _I, _I, _I, _P, _P, C.c_size_t, _P],
    "rmab_select_mark": [_I, _I, _P, _P, _I, _P, _P, C.c_size_t, _P],
    "rmab_budget_select_multi": [_I, _I, _I, _P, _P, _F, _P, _P, C.c_size_t, _P],
    "rmab_knapsack_ws_bytes": [_I, _I, _I],
    "rmab_knapsack_multichoice": [_I, _I, _I, _P, _P, _I, _P, _P, _P, C.c_size_t, _P],
    "rmab_vi_batched": [_I, _I, _I, _I, _P, _P, _P, _P, _I, _D, _D, _I, _P, _P, _P, _P, _P, _P],
    "rmab_vi_general": [_I, _I, _I, _P, _P, _D, _D, _I, _P, _P, _P, _P, _P],
    "rmab_whittle_bisect": [_I, _I, _I, _P, _P, _P, _D, _D, _I, _D, _D, _I, _P, _P],
}
_RESTYPES = {"rmab_last_error": C.c_char_p, "rmab_launch_count": C.c_uint64, "rmab_select_ws_bytes": C.c_size_t, "rmab_nat_ws_bytes": C.c_size_t, "rmab_knapsack_ws_bytes": C.c_size_t,
             "rmab_epoch_ws_bytes": C.c_size_t, "rmab_lambda_ws_bytes": C.c_size_t}

_lib = None


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built "
                "(python -m robustrmab_b200.build). robustrmab_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise RmabError(rc, lib().rmab_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(lib().rmab_launch_count())
