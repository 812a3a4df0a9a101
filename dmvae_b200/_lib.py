"""ctypes binding of libdmvae_b200.so (the C ABI declared in include/dmvae_b200.h).

There is no fallback: if the shared library is missing the import of any compute entry point raises, and a
compute call on a device that is not sm_100 raises (dmvae_check_device).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmvae_b200.so")

F32, BF16 = 0, 1
ABI_VERSION = 8          # must equal dmvae_abi_version() of the loaded library (include/dmvae_b200.h: DMVAE_ABI_VERSION)
_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes  (every function returns int except dmvae_last_error)
SIGNATURES = {
    "dmvae_abi_version": [],
    "dmvae_check_device": [],
    "dmvae_dmd_mix_xt": [_p, _p, _p, _p, _i64, _i64, _i, _p],
    "dmvae_dmd_loss_fwd_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _f, _i, _f, _i, _i, _p],
    "dmvae_l1l2_fwd": [_p, _p, _p, _i64, _p],
    "dmvae_l1l2_bwd": [_p, _p, _p, _p, _p, _i64, _f, _f, _p],
    "dmvae_l1l2_fwd_bwd": [_p, _p, _p, _p, _i64, _f, _f, _p],
    "dmvae_lpips_dist_fwd": [_p, _p, _p, _p, _i64, _i64, _i, _i, _i, _p],
    "dmvae_lpips_dist_bwd": [_p, _p, _p, _p, _p, _i64, _i64, _i, _f, _i, _p],
    "dmvae_reparam_kl_fwd": [_p, _p, _p, _p, _i64, _i64, _i, _p],
    "dmvae_reparam_kl_bwd": [_p, _p, _p, _p, _p, _f, _i64, _i64, _i, _p],
    "dmvae_gn_stats": [_p, _p, _i64, _i64, _i, _p],
    "dmvae_gn_apply": [_p, _p, _p, _p, _p, _i64, _i64, _i, _f, _i, _p],
    "dmvae_gn_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _f, _i, _p],
    "dmvae_pack_weights": [_p, _p, _p, _i, _i, _i, _i, _p],
    "dmvae_conv_tc_supported": [_i] * 7,
    "dmvae_conv_tc_fwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "dmvae_conv_tc_set_tile_mode": [_i],
    "dmvae_conv_tc_last_kernel": [],
    "dmvae_conv_tc_strided_supported": [_i] * 10,
    "dmvae_conv_tc_fwd_strided": [_p, _p, _p, _p] + [_i] * 12 + [_p],
    "dmvae_conv_tc_wgrad_strided": [_p, _p, _p] + [_i] * 12 + [_p],
    "dmvae_conv_up2x_supported": [_i] * 5,
    "dmvae_subpixel_pack": [_p, _i64, _i64, _i64, _p, _p, _i, _i, _p],
    "dmvae_conv_up2x_fwd": [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dmvae_conv_up2x_dgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dmvae_conv_up2x_wgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "dmvae_subpixel_fold_wgrad": [_p, _p, _i64, _i64, _i64, _i, _i, _p],
    "dmvae_zero_insert2x": [_p, _p, _i64, _i, _i, _i, _p],
    "dmvae_conv_tc_wgrad_supported": [_i] * 7,
    "dmvae_conv_tc_wgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p],
    "dmvae_wgrad_unpack": [_p, _p, _i, _i, _i, _i, _p],
    "dmvae_grad_patches": [_p, _p, _i64, _i, _i, _i, _i, _i, _i, _i, _p],
    "dmvae_conv_direct_fwd": [_p, _p, _p, _p, _p] + [_i] * 13 + [_p],
    "dmvae_maxpool2x2_fwd": [_p, _p, _i64, _i, _i, _i, _p],
    "dmvae_pool_tap_bwd": [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _p],
    "dmvae_relu_mask": [_p, _p, _p, _i64, _p],
    "dmvae_conv_direct_dgrad_strided": [_p, _p, _p] + [_i] * 12 + [_p],
    "dmvae_conv_direct_wgrad": [_p, _p, _p] + [_i] * 12 + [_p],
    "dmvae_bias_grad": [_p, _p, _i64, _i, _p],
    "dmvae_upsample2x_fwd": [_p, _p, _i64, _i, _i, _i, _p],
    "dmvae_upsample2x_bwd": [_p, _p, _i64, _i, _i, _i, _p],
    "dmvae_nchw_to_nhwc": [_p, _p, _i64, _i, _i64, _i, _p],
    "dmvae_nhwc_to_nchw": [_p, _p, _i64, _i, _i64, _i, _p],
    "dmvae_add_bf16": [_p, _p, _p, _i64, _p],
    "dmvae_scale_residual": [_p, _p, _p, _i64, _i, _p],
    "dmvae_layernorm_bf16": [_p, _p, _p, _p, _i64, _i, _f, _p],
    "dmvae_scale_residual_layernorm": [_p, _p, _p, _p, _p, _p, _i64, _i, _f, _p],
    "dmvae_rmsnorm_modulate": [_p, _i, _p, _p, _p, _i64, _p, _i64, _i, _i, _f, _p],
    "dmvae_qk_norm_rope": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _f, _p],
    "dmvae_grad_sumsq": [_p, _p, _i64, _p],
    "dmvae_adamw_ema_step": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i, _f, _f, _p],
    "dmvae_cast_bf16": [_p, _p, _i64, _p],
    "dmvae_pack_dgrad_bf16": [_p, _p, _i, _i, _i, _p],
    "dmvae_pack_dgrad_batched": [_p, _p, _p, _i, _i64, _p],
}

_lib: Optional[C.CDLL] = None
_device_ok = False


class DmvaeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (CPU-safe: touches no device)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DmvaeError(
                f"{LIB_PATH} is missing: build it with `python -m dmvae_b200.build` (nvcc, sm_100a). "
                "dmvae_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.dmvae_last_error.restype = C.c_char_p
        lib.dmvae_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = argtypes
        have = lib.dmvae_abi_version()
        if have != ABI_VERSION:
            raise DmvaeError(f"{LIB_PATH} exports ABI version {have}, this package binds version {ABI_VERSION}: "
                             "rebuild with `python -m dmvae_b200.build --force`")
        _lib = lib
    return _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise DmvaeError(f"unsupported dtype {t.dtype} (float32 / bfloat16 only)")


# kernels launched per entry point (everything else launches exactly one)
_LAUNCHES = {"dmvae_gn_bwd": 2}


class Stats:
    """Launch accounting and optional per-entry-point device timing (CUDA events on the launching stream).
    bench.py turns ``timing`` on for a profiled step; the timed region runs with it off."""
    launches = 0
    timing = False
    events = []          # (name, start_event, end_event, work) ; work = algorithmic flops or bytes, see bench.py
    args_log = []        # integer arguments of each timed call (shapes), parallel to ``events``
    kernel_ids = []      # dmvae_conv_tc_last_kernel() after each timed conv call (0 otherwise), parallel to ``events``
    work_fn = None       # callable(name, args) -> float

    @classmethod
    def reset(cls):
        cls.launches = 0
        cls.events = []
        cls.args_log = []
        cls.kernel_ids = []


def call(name: str, *args) -> None:
    """Invoke a compute entry point on the current CUDA stream; raises DmvaeError on failure."""
    global _device_ok
    lib = load()
    if not _device_ok:
        if not torch.cuda.is_available():
            raise DmvaeError("dmvae_b200 requires a CUDA device (sm_100a); no CPU fallback exists")
        if lib.dmvae_check_device() != 0:
            raise DmvaeError(lib.dmvae_last_error().decode())
        _device_ok = True
    Stats.launches += _LAUNCHES.get(name, 1)
    if Stats.timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, _stream())
        e1.record()
        Stats.events.append((name, e0, e1, Stats.work_fn(name, args) if Stats.work_fn else 0.0))
        Stats.kernel_ids.append(lib.dmvae_conv_tc_last_kernel() if name.startswith(("dmvae_conv_tc", "dmvae_conv_up2x")) else 0)
        Stats.args_log.append(tuple(a for a in args if isinstance(a, int) and abs(a) < (1 << 20)))
    else:
        rc = getattr(lib, name)(*args, _stream())
    if rc != 0:
        raise DmvaeError(f"{name} failed ({rc}): {lib.dmvae_last_error().decode()}")


def query(name: str, *args) -> int:
    """Pure host-side predicate (no device work)."""
    return getattr(load(), name)(*args)
