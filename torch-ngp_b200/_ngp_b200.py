"""ctypes binding of libngp_b200.so — the C-ABI declared in include/ngp_b200.h.

This is the only place the host-side packages (gridencoder, ffmlp, shencoder, raymarching) touch
native code.  There is NO fallback: if the library is missing or a call fails, a RuntimeError is
raised (the reference raises RuntimeError from its extensions as well, SURVEY §8b "Errors").
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libngp_b200.so")

_c = ctypes
_vp, _u32, _i32, _f32, _sz = _c.c_void_p, _c.c_uint32, _c.c_int, _c.c_float, _c.c_size_t

# name -> argtypes, in include/ngp_b200.h order
_SIGNATURES = {
    "ngp_grid_encode_forward": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _vp, _u32, _i32, _u32, _i32, _i32, _vp],
    "ngp_grid_encode_backward": [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _f32, _u32, _vp, _vp, _u32, _i32, _u32, _i32, _i32, _vp],
    "ngp_grad_total_variation": [_vp, _vp, _vp, _vp, _f32, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i32, _i32, _vp],
    "ngp_grid_level_scales": [_vp, _u32, _f32, _u32, _vp],
    "ngp_sh_encode_forward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp],
    "ngp_sh_encode_backward": [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp],
    "ngp_ffmlp_forward": [_vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "ngp_ffmlp_inference": [_vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp],
    "ngp_ffmlp_backward": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _i32, _vp, _vp, _vp, _vp, _sz, _vp],
    "ngp_ffmlp_backward_ex": [_vp, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _i32, _vp, _vp, _vp, _vp, _sz, _u32, _vp],
    "ngp_ffmlp_wgrad_finalize": [_vp, _vp, _u32, _i32, _vp],
    "ngp_ffmlp_allocate_splitk": [_sz],
    "ngp_ffmlp_free_splitk": [],
    "ngp_field_sigma_forward": [_vp, _vp, _vp, _u32, _f32, _u32, _u32, _i32, _vp, _u32, _u32, _i32, _vp, _vp, _vp, _vp, _vp],
    "ngp_field_color_forward": [_vp, _vp, _vp, _u32, _u32, _i32, _vp, _vp, _vp],
    "ngp_field_color_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp, _sz, _vp],
    "ngp_field_color_forward_ex": [_vp, _vp, _vp, _vp, _u32, _u32, _i32, _vp, _vp, _vp, _vp],
    "ngp_field_color_backward_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp, _sz, _u32, _vp],
    "ngp_infer_init": [_u32, _vp, _vp, _vp],
    "ngp_march_rays_dev": [_vp, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ngp_composite_rays_dev": [_vp, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ngp_compact_rays_dev": [_vp, _u32, _u32, _vp, _vp, _vp],
    "ngp_field_sigma_forward_dev": [_vp, _f32, _vp, _vp, _vp, _u32, _f32, _u32, _u32, _i32, _vp, _u32, _u32, _vp, _vp, _vp],
    "ngp_field_color_forward_dev": [_vp, _vp, _vp, _vp, _u32, _u32, _vp, _vp],
    "ngp_debug_umma": [_vp, _vp, _vp, _i32, _vp],
    "ngp_debug_red_probe": [_vp, _u32, _u32, _u32, _i32, _vp],
    "ngp_optim_check_finite": [_vp, _i32, _c.c_uint64, _vp, _vp],
    "ngp_optim_adam_step": [_vp, _vp, _vp, _vp, _i32, _vp, _c.c_uint64, _f32, _f32, _f32, _f32, _vp, _i32, _vp],
    "ngp_optim_scaler_update": [_vp, _f32, _f32, _i32, _vp],
    "ngp_exchange_barrier": [_vp, _u32, _u32, _u32, _vp, _vp, _u32, _vp],
    "ngp_exchange_reduce": [_vp, _u32, _u32, _c.c_uint64, _c.c_uint64, _vp, _vp],
    "ngp_exchange_adam": [_vp, _vp, _vp, _vp, _vp, _u32, _c.c_uint64, _c.c_uint64, _c.c_uint64, _f32, _f32, _f32, _f32, _vp, _vp],
    "ngp_exchange_zero": [_vp, _c.c_uint64, _vp],
    "ngp_exchange_reduce_fused": [_vp, _vp, _vp, _u32, _u32, _c.c_uint64, _c.c_uint64, _vp, _u32, _vp],
    "ngp_exchange_adam_fused": [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _c.c_uint64, _c.c_uint64, _c.c_uint64,
                                _f32, _f32, _f32, _f32, _vp, _u32, _vp],
    "ngp_exchange_finish": [_vp, _u32, _u32, _vp, _f32, _f32, _i32, _u32, _vp],
    "ngp_composite_rays_train_forward_mse": [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ngp_step_counter_push": [_vp, _vp, _vp, _vp, _vp],
    "ngp_density_grid_mark_untrained": [_vp, _u32, _f32, _f32, _f32, _f32, _f32, _u32, _u32, _vp, _vp, _vp, _vp],
    "ngp_density_grid_occupied": [_vp, _u32, _u32, _vp, _vp, _vp, _vp],
    "ngp_density_grid_sample_full": [_u32, _u32, _f32, _vp, _vp, _vp],
    "ngp_density_grid_sample_partial": [_u32, _u32, _f32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ngp_density_grid_update": [_vp, _vp, _vp, _vp, _u32, _f32, _f32, _f32, _u32, _u32, _vp, _vp, _vp, _vp],
    "ngp_get_rays": [_vp, _u32, _f32, _f32, _f32, _f32, _u32, _u32, _u32, _vp, _u32, _vp, _vp, _vp],
    "ngp_gather_pixels": [_vp, _i32, _vp, _u32, _u32, _u32, _u32, _u32, _vp, _u32, _i32, _vp, _f32, _vp, _vp, _vp],
    "ngp_freq_encode_forward": [_vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "ngp_freq_encode_backward": [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp],
    "ngp_near_far_from_aabb": [_vp, _vp, _vp, _u32, _f32, _vp, _vp, _vp],
    "ngp_sph_from_ray": [_vp, _vp, _f32, _u32, _vp, _vp],
    "ngp_morton3D": [_vp, _u32, _vp, _vp],
    "ngp_morton3D_invert": [_vp, _u32, _vp, _vp],
    "ngp_packbits": [_vp, _u32, _f32, _vp, _vp],
    "ngp_march_rays_train": [_vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ngp_composite_rays_train_forward": [_vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp, _vp],
    "ngp_composite_rays_train_backward": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _f32, _vp, _vp, _vp],
    "ngp_march_rays": [_u32, _u32, _vp, _vp, _vp, _vp, _f32, _f32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ngp_composite_rays": [_u32, _u32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
}
# every symbol include/ngp_b200.h declares (tests check the .so exports all of them)
EXPORTED = sorted(list(_SIGNATURES) + ["ngp_debug_set_mlp_backward", "ngp_debug_set_sigma_gather", "ngp_last_error", "ngp_version", "ngp_build_arch", "ngp_launch_count",
                                       "ngp_reset_launch_count", "ngp_ffmlp_backward_workspace_bytes",
                                       "ngp_density_grid_occupied_scratch_bytes", "ngp_density_grid_update_scratch_bytes",
                                       "ngp_peer_alloc", "ngp_peer_open", "ngp_peer_close", "ngp_peer_free", "ngp_exchange_pad_bytes",
                                       "ngp_exchange_error"])

_lib = None


def load():
    """Load (once) and return the ctypes handle.  Raises RuntimeError if the .so is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libngp_b200.so not found at {LIB_PATH}: build it with `python torch-ngp_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _c.c_int
    lib.ngp_last_error.restype = _c.c_char_p
    lib.ngp_build_arch.restype = _c.c_char_p
    lib.ngp_version.restype = _c.c_int
    lib.ngp_launch_count.restype = _c.c_uint64
    lib.ngp_reset_launch_count.restype = None
    lib.ngp_debug_set_mlp_backward.argtypes = [_i32]
    lib.ngp_debug_set_mlp_backward.restype = _c.c_int
    lib.ngp_debug_set_sigma_gather.argtypes = [_i32, _i32]
    lib.ngp_debug_set_sigma_gather.restype = _c.c_int
    lib.ngp_ffmlp_backward_workspace_bytes.argtypes = [_u32, _u32, _u32, _u32, _u32]
    lib.ngp_ffmlp_backward_workspace_bytes.restype = _sz
    for name in ("ngp_density_grid_occupied_scratch_bytes", "ngp_density_grid_update_scratch_bytes"):
        getattr(lib, name).argtypes = [_u32, _u32]
        getattr(lib, name).restype = _sz
    # peer-visible memory (CUDA IPC) and the exchange's host-side helpers: no stream argument
    lib.ngp_peer_alloc.argtypes = [_sz, _c.POINTER(_vp), _vp]
    lib.ngp_peer_open.argtypes = [_vp, _c.POINTER(_vp)]
    lib.ngp_peer_close.argtypes = [_vp]
    lib.ngp_peer_free.argtypes = [_vp]
    lib.ngp_exchange_pad_bytes.argtypes = []
    lib.ngp_exchange_pad_bytes.restype = _sz
    lib.ngp_exchange_error.argtypes = [_vp, _c.POINTER(_u32)]
    for name in ("ngp_peer_alloc", "ngp_peer_open", "ngp_peer_close", "ngp_peer_free", "ngp_exchange_error"):
        getattr(lib, name).restype = _c.c_int
    _lib = lib
    return lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


_profile = None   # when a list: (name, args, start_event, end_event) per call (bench.py roofline timing)


def profile_begin():
    global _profile
    _profile = []


def profile_end():
    global _profile
    rec, _profile = _profile, None
    return rec


def call(name, *args):
    """Invoke a C-ABI entry point on the current stream; raise RuntimeError on a non-zero code."""
    lib = load()
    if _profile is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream())
        e1.record()
        _profile.append((name, args, e0, e1))
    else:
        rc = getattr(lib, name)(*args, stream())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.ngp_last_error().decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ngp_b200: expected a CUDA tensor (this build has no CPU path)")


def launch_count():
    return int(load().ngp_launch_count())


def reset_launch_count():
    load().ngp_reset_launch_count()
