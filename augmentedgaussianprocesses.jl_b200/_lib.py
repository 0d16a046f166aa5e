"""ctypes binding of libagp_b200.so (the C ABI of include/agp_b200.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded this module raises, and
every compute entry point fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libagp_b200.so")

AGP_OK, AGP_ERR_BAD_ARG, AGP_ERR_CUDA, AGP_ERR_KTILDE_NONPOS, AGP_ERR_NOT_POSDEF, AGP_ERR_STATE = range(6)
KERNEL_SQEXP, KERNEL_MATERN32, KERNEL_MATERN52 = 0, 1, 2
LIK_GAUSSIAN, LIK_LOGISTIC, LIK_STUDENTT, LIK_LOGISTICSOFTMAX = 0, 1, 2, 3
LIK_LAPLACE, LIK_BAYESIANSVM, LIK_NEGBINOMIAL, LIK_POISSON, LIK_HETEROSCEDASTIC = 4, 5, 6, 7, 8
MODEL_SVGP, MODEL_MOSVGP, MODEL_VGP, MODEL_MOVGP = 0, 1, 2, 3
PREC_F64, PREC_F32, PREC_TF32X3 = 0, 1, 2
DTYPE_F64, DTYPE_F32 = 0, 1
LAYOUT_COLMAJOR, LAYOUT_ROWMAJOR = 0, 1
Y_REAL, Y_CLASS = 0, 1
PRECISIONS = {"f64": PREC_F64, "f32": PREC_F32, "tf32x3": PREC_TF32X3}

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class ModelDesc(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int32),
        ("n_latent_global", C.c_int32),
        ("latent_begin", C.c_int32),
        ("n_latent_local", C.c_int32),
        ("m", C.c_int32),
        ("D", C.c_int32),
        ("batch_capacity", C.c_int32),
        ("precision", C.c_int32),
        ("stochastic", C.c_int32),
        ("rm_kappa", C.c_double),
        ("rm_tau", C.c_double),
        ("jitter", C.c_double),
        ("n_task", C.c_int32),
        ("lik_kind", c_int32_p),
        ("lik_p0", c_double_p),
        ("lik_p1", c_double_p),
        ("A", c_double_p),
        ("kernel_kind", c_int32_p),
        ("kernel_scale", c_double_p),
        ("kernel_variance", c_double_p),
        ("Z", c_double_p),
        ("mu0", c_double_p),
    ]


# every symbol include/agp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "agp_abi_version": (C.c_int, []),
    "agp_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "agp_ctx_destroy": (None, [C.c_void_p]),
    "agp_last_error": (C.c_char_p, [C.c_void_p]),
    "agp_model_create": (C.c_int, [C.c_void_p, C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "agp_model_destroy": (None, [C.c_void_p]),
    "agp_data_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_int]),
    "agp_minibatches_upload": (C.c_int, [C.c_void_p, c_int64_p, C.c_int64, C.c_int32, C.c_int32]),
    "agp_refresh_K": (C.c_int, [C.c_void_p]),
    "agp_state_reset": (C.c_int, [C.c_void_p]),
    "agp_set_kernel": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double]),
    "agp_step": (C.c_int, [C.c_void_p, c_int64_p, C.c_int32, C.c_int32, C.c_double]),
    "agp_step_async": (C.c_int, [C.c_void_p, c_int64_p, C.c_int32, C.c_int32, C.c_double]),
    "agp_step_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_double]),
    "agp_step_batch_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_double,
                                       c_int64_p]),
    "agp_result_wait": (C.c_int, [C.c_void_p, C.c_int64, c_double_p]),
    "agp_sync": (C.c_int, [C.c_void_p]),
    "agp_step_moments_async": (C.c_int, [C.c_void_p, c_int64_p, C.c_int32, C.c_int32]),
    "agp_step_update_async": (C.c_int, [C.c_void_p, C.c_double]),
    "agp_moments_devptr": (C.c_void_p, [C.c_void_p, C.c_int32, c_int64_p]),
    "agp_elbo": (C.c_int, [C.c_void_p, C.c_double, c_double_p]),
    "agp_elbo_moments_async": (C.c_int, [C.c_void_p]),
    "agp_get_posterior": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]),
    "agp_set_posterior": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p]),
    "agp_get_counters": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "agp_set_counters": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "agp_get_local": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, c_double_p, C.c_int32]),
    "agp_get_kernel_matrices": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p, C.c_int32]),
    "agp_get_Kinv": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p]),
    "agp_predict_f": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, c_double_p, c_double_p]),
    "agp_proba_logistic": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p, C.c_int32, c_double_p, c_double_p]),
    "agp_set_step_size": (C.c_int, [C.c_void_p, C.c_double]),
    "agp_set_noise_optimiser": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]),
    "agp_hyper_grads": (C.c_int, [C.c_void_p, C.c_double, c_double_p, c_double_p, c_double_p]),
    "agp_set_Z": (C.c_int, [C.c_void_p, C.c_int32, c_double_p]),
    "agp_keep_stale_K": (C.c_int, [C.c_void_p, C.c_int32]),
    "agp_set_A_optimiser": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]),
    "agp_get_A": (C.c_int, [C.c_void_p, c_double_p]),
    "agp_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "agp_peer_attach": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "agp_peer_detach": (C.c_int, [C.c_void_p]),
    "agp_set_quadrature": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int32]),
    "agp_get_lik_param": (C.c_int, [C.c_void_p, C.c_int32, c_double_p]),
    "agp_set_lik_param": (C.c_int, [C.c_void_p, C.c_int32, C.c_double]),
    "agp_proba_link": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "agp_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "agp_profile_read": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), c_double_p, c_int64_p]),
    "agp_launch_count": (C.c_int64, [C.c_void_p]),
    "agp_time_kernel": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p]),
    "agp_use_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "agp_predict_f_cov": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "agp_online_carry": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, C.c_int32, c_double_p, c_double_p, C.c_double]),
    "agp_online_extra_kl": (C.c_int, [C.c_void_p, c_double_p]),
    "agp_local_updates_async": (C.c_int, [C.c_void_p]),
    "agp_step_with_gradients": (C.c_int, [C.c_void_p, c_int64_p, C.c_int32, C.c_int32, c_double_p, c_double_p]),
    "agp_experimental_ns_refine": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p, C.c_int32, C.c_int32, c_double_p, c_double_p]),
}

_lib = None


class AGPError(RuntimeError):
    """Raised for every non-zero status of the C ABI; `.code` holds the AGP_ERR_* value."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class KtildeError(AGPError):
    """gpblocks/latentgp.jl:213  error("K̃ has negative values")"""


class PosDefException(AGPError):
    """LinearAlgebra.PosDefException raised by cholesky() in the reference"""


def load():
    """dlopen libagp_b200.so and type every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python augmentedgaussianprocesses.jl_b200/build.py` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.agp_abi_version() != 1:
        raise ImportError("libagp_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(ctx, code: int):
    if code == AGP_OK:
        return
    msg = load().agp_last_error(ctx)
    msg = msg.decode("utf-8", "replace") if msg else f"agp error {code}"
    if code == AGP_ERR_KTILDE_NONPOS:
        raise KtildeError(code, msg)
    if code == AGP_ERR_NOT_POSDEF:
        raise PosDefException(code, msg)
    if code == AGP_ERR_CUDA and not msg:
        msg = "CUDA error (is a GPU present? there is no CPU fallback)"
    raise AGPError(code, msg)


def dptr(a):
    return a.ctypes.data_as(c_double_p)
