"""ctypes binding of libsvae_b200.so (C ABI declared in include/svae_b200.h).

The library is the product: there is no eager / CPU fallback.  If the shared object is missing
the import of this module raises, and every call checks the return code and raises with the
library's own error text.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvae_b200.so")

MAX_LAYERS = 8

ACT_TANH, ACT_LEAKYRELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
LIK_BERNOULLI, LIK_GAUSS, LIK_GAUSS_FITNOISE = 0, 1, 2
PRECISION_PARITY, PRECISION_FAST, PRECISION_PARITY_TC = 0, 1, 2
ENC_RESID = 0x100          # OR-ed into the activation argument of svae_encoder_forward/backward

ACT_CODES = {"tanh": ACT_TANH, "leakyrelu": ACT_LEAKYRELU, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID}
PRECISION_CODES = {"parity": PRECISION_PARITY, "fast": PRECISION_FAST, "parity_tc": PRECISION_PARITY_TC}

EXPORTS = [
    "svae_version", "svae_last_error", "svae_launch_count", "svae_device_sm_count", "svae_workspace_bytes",
    "svae_encoder_forward", "svae_encoder_backward", "svae_decoder_forward", "svae_decoder_backward",
    "svae_step", "svae_adam_step", "svae_adam_tick", "svae_adam_step_graph", "svae_gather_rows", "svae_rotation_matrices", "svae_rotate_bicubic", "svae_ctf_filter", "svae_sm_clock_probe", "svae_gemm_bf16",
    "svae_gemm_dx_moments", "svae_gemm_dw_top", "svae_resid_linear_forward", "svae_resid_linear_backward",
]


class SvaeShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "P", "n_rows", "n_cols", "C", "Cin", "Z", "I", "H", "L", "Hq", "Lq", "k_ctf")]


class SvaeConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("rotate", "translate", "likelihood", "theta_kl_mean", "activation", "precision", "softplus",
                 "chunk_images")] + \
               [(n, C.c_float) for n in ("theta_prior", "dx_scale", "z_scale", "grad_scale")] + \
               [(n, C.c_int32) for n in ("resid", "expand_coords", "bilinear")]


class SvaeDecoderParams(C.Structure):
    _fields_ = [("coord_w", C.c_void_p), ("coord_b", C.c_void_p), ("latent_w", C.c_void_p),
                ("hidden_w", C.c_void_p * MAX_LAYERS), ("hidden_b", C.c_void_p * MAX_LAYERS),
                ("out_w", C.c_void_p), ("out_b", C.c_void_p), ("bilinear_w", C.c_void_p)]


class SvaeEncoderParams(C.Structure):
    _fields_ = [("w", C.c_void_p * (MAX_LAYERS + 1)), ("b", C.c_void_p * (MAX_LAYERS + 1))]


class SvaeStepInputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("grid", "y", "y_enc", "theta_offset", "eps", "ctf", "mask", "rng_step")] + \
               [("rng_seed", C.c_uint64), ("rng_image_offset", C.c_int64), ("decoder_grads_event", C.c_void_p)]


class SvaeStepOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("stats", "y_hat", "latent", "stats_sum")]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (or `make -C spatial-vae_b200/csrc`). "
            "spatial_vae (B200) has no CPU or eager fallback.")
    return declare(C.CDLL(LIB_PATH))


def declare(lib):
    """Attach the argument / return types of include/svae_b200.h to a loaded library object."""
    vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    P = C.POINTER
    lib.svae_version.restype = i32
    lib.svae_last_error.argtypes = [C.c_char_p, i32]
    lib.svae_device_sm_count.restype = i32
    lib.svae_launch_count.restype = C.c_ulonglong
    lib.svae_workspace_bytes.argtypes = [P(SvaeShape), P(SvaeConfig), P(sz)]
    lib.svae_encoder_forward.argtypes = [P(SvaeShape), i32, P(SvaeEncoderParams), vp, vp, vp, vp]
    lib.svae_encoder_backward.argtypes = [P(SvaeShape), i32, P(SvaeEncoderParams), vp, vp, vp, P(SvaeEncoderParams),
                                          vp, vp, vp]
    lib.svae_decoder_forward.argtypes = [P(SvaeShape), P(SvaeConfig), P(SvaeDecoderParams), vp, vp, vp, vp, sz, vp]
    lib.svae_decoder_backward.argtypes = [P(SvaeShape), P(SvaeConfig), P(SvaeDecoderParams), vp, vp, vp,
                                          P(SvaeDecoderParams), vp, vp, vp, sz, vp]
    lib.svae_step.argtypes = [P(SvaeShape), P(SvaeConfig), P(SvaeDecoderParams), P(SvaeEncoderParams),
                              P(SvaeStepInputs), P(SvaeStepOutputs), P(SvaeDecoderParams), P(SvaeEncoderParams),
                              vp, sz, vp]
    lib.svae_adam_step.argtypes = [vp, vp, vp, vp, sz, f32, f32, f32, f32, i32, i32, vp]
    lib.svae_adam_tick.argtypes = [vp, vp, f32, f32, vp]
    lib.svae_adam_step_graph.argtypes = [vp, vp, vp, vp, sz, f32, f32, f32, f32, vp, i32, vp]
    lib.svae_gather_rows.argtypes = [vp, vp, vp, C.c_int64, C.c_int64, vp]
    lib.svae_rotation_matrices.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.svae_rotate_bicubic.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.svae_ctf_filter.argtypes = [vp, i32, i32, i32, C.c_double, vp, vp]
    lib.svae_sm_clock_probe.argtypes = [vp, vp]
    lib.svae_gemm_bf16.argtypes = [i32, i32, i32, i32, vp, i32, vp, i32, vp, vp, i32, i32, vp, i32, vp]
    lib.svae_gemm_dx_moments.argtypes = [i32, i32, i32, vp, i32, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp]
    lib.svae_gemm_dw_top.argtypes = [i32, i32, i32, vp, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.svae_resid_linear_forward.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.svae_resid_linear_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError if the library lacks a declared symbol
        if name not in ("svae_version", "svae_device_sm_count", "svae_launch_count"):
            getattr(lib, name).restype = i32
    return lib


lib = _load()


class SvaeError(RuntimeError):
    pass


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib.svae_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        kind = {-1: "SVAE_EINVAL", -2: "SVAE_EALIGN", -3: "SVAE_ECUDA", -4: "SVAE_ENOSPACE"}.get(rc, str(rc))
        msg = f"{what} failed with {kind}: {last_error()}"
        if rc == -1:
            raise SvaeError(msg)
        raise SvaeError(msg)
