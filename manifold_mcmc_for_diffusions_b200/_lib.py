"""ctypes binding of the C ABI in ``include/mmd_b200.h`` (``libmmd_b200.so``).

There is no CPU fallback: importing works without a GPU (so the symbol table can be checked),
but creating a handle raises :class:`MmdError` when no CUDA device is present, and a missing
shared library raises at import of this module's ``lib()``.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMD_B200_LIB", os.path.join(_HERE, "libmmd_b200.so"))  # override: tuning builds


class MmdError(RuntimeError):
    pass


class MmdConfig(C.Structure):
    _fields_ = [
        ("model", C.c_int),
        ("num_obs", C.c_int),
        ("num_steps_per_obs", C.c_int),
        ("num_obs_per_subseq", C.c_int),
        ("dim_u", C.c_int),
        ("noise", C.c_int),
        ("sigma_fixed", C.c_double),
        ("gaussian_splitting", C.c_int),
        ("obs_interval", C.c_double),
        ("y_seq", C.POINTER(C.c_double)),
        ("n_chains", C.c_int),
        ("device", C.c_int),
    ]


class MmdIntegratorOpts(C.Structure):
    _fields_ = [
        ("solver", C.c_int),
        ("constraint_tol", C.c_double),
        ("position_tol", C.c_double),
        ("divergence_tol", C.c_double),
        ("max_iters", C.c_int),
        ("reverse_check_tol", C.c_double),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_H = C.c_void_p

# name -> (restype, argtypes); the list is the ABI: tests check every name is exported
SIGNATURES = {
    "mmd_create": (C.c_int, [C.POINTER(MmdConfig), C.POINTER(_H)]),
    "mmd_destroy": (C.c_int, [_H]),
    "mmd_last_error_string": (C.c_char_p, []),
    "mmd_dim_q": (C.c_int, [_H]),
    "mmd_num_partition": (C.c_int, [_H]),
    "mmd_num_constraints": (C.c_int, [_H, C.c_int]),
    "mmd_num_blocks": (C.c_int, [_H, C.c_int]),
    "mmd_n_chains": (C.c_int, [_H]),
    "mmd_set_state": (C.c_int, [_H, _dp, _dp, _dp, C.c_int]),
    "mmd_set_state_async": (C.c_int, [_H, _dp, _dp, _dp, C.c_int]),
    "mmd_set_chain_regrouping": (C.c_int, [_H, C.c_int]),
    "mmd_get_slot_chains": (C.c_int, [_H, C.POINTER(C.c_int)]),
    "mmd_get_state": (C.c_int, [_H, _dp, _dp, _dp]),
    "mmd_set_momentum": (C.c_int, [_H, _dp]),
    "mmd_set_state_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "mmd_get_state_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmd_get_stream": (C.c_void_p, [_H]),
    "mmd_get_device": (C.c_int, [_H]),
    "mmd_wait_stream": (C.c_int, [_H, C.c_void_p]),
    "mmd_stream_wait": (C.c_int, [_H, C.c_void_p]),
    "mmd_get_state_async": (C.c_int, [_H, _dp, _dp, _dp]),
    "mmd_get_head": (C.c_int, [_H, _dp]),
    "mmd_chains_per_tile": (C.c_int, [_H]),
    "mmd_set_chain_offset": (C.c_int, [_H, C.c_int]),
    "mmd_get_partition": (C.c_int, [_H]),
    "mmd_linearize": (C.c_int, [_H, C.c_int]),
    "mmd_constr": (C.c_int, [_H, _dp]),
    "mmd_log_det_sqrt_gram": (C.c_int, [_H, _dp]),
    "mmd_grad_log_det_sqrt_gram": (C.c_int, [_H, _dp]),
    "mmd_hamiltonian": (C.c_int, [_H, _dp]),
    "mmd_project_momentum": (C.c_int, [_H]),
    "mmd_normal_space_component": (C.c_int, [_H, _dp, _dp]),
    "mmd_update_x_obs_seq": (C.c_int, [_H]),
    "mmd_switch_partition": (C.c_int, [_H]),
    "mmd_sample_momentum": (C.c_int, [_H, C.c_uint64, C.c_uint64]),
    "mmd_get_factor": (C.c_int, [_H, C.c_char_p, _dp, _ip]),
    "mmd_default_integrator_opts": (None, [C.POINTER(MmdIntegratorOpts)]),
    "mmd_num_generator_params": (C.c_int, [_H]),
    "mmd_get_generator_params": (C.c_int, [_H, _dp]),
    "mmd_set_generator_params": (C.c_int, [_H, _dp, C.c_int]),
    "mmd_leapfrog_step": (C.c_int, [_H, C.c_double, C.POINTER(MmdIntegratorOpts)]),
    "mmd_leapfrog_step_inner": (C.c_int, [_H, C.c_double, C.c_int, C.POINTER(MmdIntegratorOpts)]),
    "mmd_get_step_info": (C.c_int, [_H, _ip, _ip, _ip, _dp]),
    "mmd_project_quasi_newton": (
        C.c_int,
        [_H, _dp, C.c_double, C.POINTER(MmdIntegratorOpts), _dp, _ip, _ip],
    ),
    "mmd_init_linear_interpolation": (C.c_int, [_H, _dp, _dp, _dp, C.c_int]),
    "mmd_hmc_transition": (
        C.c_int,
        [_H, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(MmdIntegratorOpts), C.c_int],
    ),
    "mmd_get_transition_stats": (C.c_int, [_H, _ip, _dp, _ip]),
    "mmd_transition_begin": (C.c_int, [_H, C.c_uint64, C.c_uint64]),
    "mmd_transition_steps": (C.c_int, [_H, C.c_double, C.c_int, C.POINTER(MmdIntegratorOpts)]),
    "mmd_transition_end": (C.c_int, [_H, C.c_uint64, C.c_uint64, C.c_int]),
    "mmd_set_step_sizes": (C.c_int, [_H, _dp]),
    "mmd_get_step_sizes": (C.c_int, [_H, _dp]),
    "mmd_adapt_start": (C.c_int, [_H, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "mmd_adapt_start_per_chain": (C.c_int, [_H, _dp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "mmd_adapt_stop": (C.c_int, [_H, C.c_int]),
    "mmd_adapt_update": (C.c_int, [_H, _dp]),
    "mmd_aux_reserve": (C.c_int, [_H, C.c_int]),
    "mmd_vec_axpby": (C.c_int, [_H, C.c_int, C.c_int, C.c_double, C.c_double, _ip]),
    "mmd_vec_uturn": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]),
    "mmd_set_inactive": (C.c_int, [_H, _ip, C.c_int]),
    "mmd_relinearize": (C.c_int, [_H]),
    "mmd_successful_steps": (C.c_longlong, [_H, C.c_int]),
    "mmd_total_qn_iterations": (C.c_longlong, [_H, C.c_int]),
    "mmd_debug_phase_cycles": (C.c_int, [_H, C.POINTER(C.c_ulonglong), C.c_int]),
    "mmd_profile_enable": (C.c_int, [_H, C.c_int, C.c_int]),
    "mmd_profile_summary": (C.c_int, [_H, C.c_int, _ip, _dp]),
    "mmd_hmc_dim": (C.c_int, [_H]),
    "mmd_hmc_target": (C.c_int, [_H, _dp, C.c_int, _dp, _dp, _dp]),
    "mmd_adam_begin": (C.c_int, [_H, _dp, _ip]),
    "mmd_adam_eval": (C.c_int, [_H, _dp, _dp]),
    "mmd_adam_update": (C.c_int, [_H, C.c_double, _ip]),
    "mmd_adam_get": (C.c_int, [_H, _dp, _dp]),
    "mmd_launch_count": (C.c_longlong, [_H]),
    "mmd_timer_start": (C.c_int, [_H]),
    "mmd_timer_stop_ms": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "mmd_synchronize": (C.c_int, [_H]),
}

_lib = None


def lib():
    """Load ``libmmd_b200.so`` (built in-tree by ``__graft_entry__.build()``); raise if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmdError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise MmdError(lib().mmd_last_error_string().decode())
