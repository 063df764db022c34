"""ctypes binding of libnerf_b200.so (C ABI declared in include/nerf_b200.h).

There is no CPU fallback: if the library is missing or a call fails the
product raises -- it never routes through PyTorch eager ops or the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NERFB_VARIANT selects an experiment build (tools/, never set in production); the default is the product library
_VARIANT = os.environ.get("NERFB_VARIANT", "")
LIB_PATH = os.path.join(_HERE, f"libnerf_b200_{_VARIANT}.so" if _VARIANT else "libnerf_b200.so")

NERFB_NO_VIEWS = -1000
NERFB_NUM_PARAMS = 26
NERFB_ADAM_MAX = 64

_vp, _i, _f, _d, _ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong

_SIGS = {
    "nerfb_abi_version": ([], _i),
    "nerfb_last_error": ([], C.c_char_p),
    "nerfb_device_info": ([C.POINTER(_i)] * 3, _i),
    "nerfb_get_rays": ([_i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp], _i),
    "nerfb_ndc_rays": ([_i, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_pack_rays": ([_i, _vp, _vp, _vp, _i, _i, _i, _i, _d, _f, _f, _vp, _vp], _i),
    "nerfb_gather_ray_batch": ([_i, _i, _i, _f, _f, _f, _f, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp], _i),
    "nerfb_coarse_z": ([_i, _i, _vp, _i, _i, _vp, _vp, _vp], _i),
    "nerfb_embed": ([_i, _vp, _i, _i, _vp, _i, _i, _vp], _i),
    "nerfb_embed_samples": ([_i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp], _i),
    "nerfb_linear_fwd": ([_i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp], _i),
    "nerfb_linear_bwd_data": ([_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _vp], _i),
    "nerfb_linear_bwd_weight": ([_i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp], _i),
    "nerfb_raw2outputs_fwd": ([_i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_raw2outputs_bwd": ([_i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_sample_pdf": ([_i, _i, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "nerfb_hierarchical_z": ([_i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_composite_resample": ([_i, _i, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  _vp], _i),
    "nerfb_packed_weights_bytes": ([_i], C.c_size_t),
    "nerfb_pack_weights": ([C.POINTER(_vp), _i, _i, _vp, _vp], _i),
    "nerfb_mlp_fwd_bf16": ([_i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp], _i),
    "nerfb_mlp_fwd_composite_bf16": ([_i, _i, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp], _i),
    "nerfb_composite_finish": ([_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_resample_records": ([_i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_train_store_bytes": ([_ll, _i], C.c_size_t),
    "nerfb_mlp_fwd_save_bf16": ([_i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_packed_weights_bwd_bytes": ([], C.c_size_t),
    "nerfb_pack_weights_bwd": ([C.POINTER(_vp), _vp, _vp], _i),
    "nerfb_mlp_bwd_chain_bf16": ([_ll, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_mlp_wgrad_bf16": ([_ll, _vp, _vp, _vp, C.POINTER(_vp), _vp], _i),
    "nerfb_mlp_bwd_fused_bf16": ([_ll, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp, _i, _i, _vp], _i),
    "nerfb_mlp_fwd_save_nv_bf16": ([_i, _i, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_pack_weights_bwd_nv": ([C.POINTER(_vp), _i, _vp, _vp], _i),
    "nerfb_mlp_bwd_chain_nv_bf16": ([_ll, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_mlp_wgrad_nv_bf16": ([_ll, _i, _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp], _i),
    "nerfb_umma_selftest": ([_i, _i, _i, _vp, _vp, _vp, _vp], _i),
    "nerfb_umma_selftest_mn": ([_i, _i, _vp, _vp, _vp, _vp], _i),
    "nerfb_umma_selftest_bias": ([_i, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_bulk_probe": ([_vp, _i, _i, _i, _i, _i, _i, _vp, _vp], _i),
    "nerfb_contention_probe": ([_vp, _i, _i, _i, _i, _i, _vp, _vp], _i),
    "nerfb_debug_schedule": ([_i] + [C.POINTER(_i)] * 6, _i),
    "nerfb_watchdog_record": ([], C.POINTER(C.c_ulonglong)),
    "nerfb_set_trace": ([_vp], _i),
    "nerfb_keyed_rng": ([C.c_ulonglong, _ll, _i, _i, _i, _i, _f, _vp, _vp], _i),
    "nerfb_render_rays_rng": ([C.c_ulonglong, _ll, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp], _i),
    "nerfb_image_metrics_workspace_bytes": ([_i, _i], C.c_size_t),
    "nerfb_image_metrics": ([_i, _i, _i, _vp, _vp, _f, _i, _f, _f, _f, _vp, _vp, _vp, _vp], _i),
    "nerfb_adam_step": ([_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_ll),
                         _f, _f, _f, _f, _i, _f, _vp], _i),
}

_lib = None


class NerfB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NerfB200Error(
            f"{LIB_PATH} is missing: build it with `python -m nerf_projects_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in _SIGS.items():
        if _VARIANT and not hasattr(lib, name):
            continue             # experiment builds of older sources may lack newer entry points
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = args
        fn.restype = res
    if lib.nerfb_abi_version() != 1:
        raise NerfB200Error("libnerf_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc: int, what: str):
    if rc != 0:
        msg = load().nerfb_last_error().decode(errors="replace")
        raise NerfB200Error(f"{what} failed (code {rc}): {msg}")


def watchdog_record():
    """Decoded watchdog record of the tensor-core kernels (None if it never fired)."""
    rec = load().nerfb_watchdog_record()
    if not rec or rec[0] != 0xDEAD:
        return None
    return dict(block=rec[1] >> 32, thread=rec[1] & 0xFFFFFFFF, tag=rec[2], bar=rec[3] >> 32, parity=rec[3] & 0xFFFFFFFF)


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Kernels are enqueued on the CURRENT device's current stream, so a
    CUDA tensor on another device is refused instead of being dereferenced by the wrong GPU (the public entry points --
    render, create_nerf's networks, train -- switch to the tensors' device themselves, see ``on_device``)."""
    if t is None:
        return None
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise NerfB200Error(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                            "wrap the call in `with torch.cuda.device(tensor.device):`")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def on_device(dev):
    """Context manager: make ``dev`` (a torch.device or a CUDA tensor's device) the current CUDA device."""
    import contextlib
    dev = torch.device(dev)
    if dev.type != "cuda":
        return contextlib.nullcontext()
    return torch.cuda.device(dev)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NerfB200Error("nerf_projects_b200 kernels need CUDA tensors; got a tensor on %s "
                                "(there is no CPU path)" % t.device)


def f32c(t):
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
