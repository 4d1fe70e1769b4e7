"""ctypes binding of libfmc_b200.so (include/fmc_b200.h).  The library is the product: if it is missing or an
entry point is absent the import of any op fails loudly -- there is no fallback path."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# FMC_B200_LIB: A/B kernel experiments load another build of the same library (profiles/variants/*.so)
LIB_PATH = os.environ.get("FMC_B200_LIB") or os.path.join(_HERE, "libfmc_b200.so")

P, I, L, F, D = c_void_p, c_int, c_longlong, c_float, c_double
ABI_VERSION = 3  # FMC_B200_ABI_VERSION of include/fmc_b200.h

# name -> argtypes, in the order of include/fmc_b200.h
SIGNATURES = {
    "fmc_gemm_bf16": [P, L, P, L, P, L, I, I, I, P, P, L, P, I, L, I, I, P],
    "fmc_gemm_ln_bf16": [P, L, P, L, P, L, I, I, I, P, P, P, I, I, P],
    "fmc_rowstats_bf16": [P, L, P, L, I, F, P],
    "fmc_conv3x3_bf16": [P, P, P, P, P, I, I, I, I, I, I, I, P],
    "fmc_spatial_attn_bf16": [P, L, I, L, P, L, I, P, L, I, L, I, P, L, I, I, I, I, I, I, I, F, P],
    "fmc_spatial_attn_vf16": [P, L, I, L, P, L, I, P, L, I, L, I, P, L, I, I, I, I, I, I, I, F, P],
    "fmc_spatial_attn_lse_bf16": [P, L, I, L, P, L, I, P, L, I, L, I, P, L, P, I, I, I, I, I, I, I, F, P],
    "fmc_temporal_attn_bf16": [P, L, I, I, I, I, P, L, I, I, I, I, I, F, P],
    "fmc_temporal_qkv_attn_bf16": [P, L, P, L, P, L, I, I, I, I, I, F, P],
    "fmc_debug_set_timeline": [P],
    "fmc_layernorm_bf16": [P, L, P, P, F, P, L, P, I, I, P, L, P, L, L, I, P],
    "fmc_groupnorm_bf16": [P, L, P, P, F, P, L, P, I, I, I, I, I, P, L, I, P],
    "fmc_add_bf16": [P, L, P, L, P, I, L, P, L, L, I, I, P],
    "fmc_resize_nearest_bf16": [P, P, I, I, I, I, I, I, P],
    "fmc_avgpool2_bf16": [P, P, I, I, I, I, P],
    "fmc_copy2d_bf16": [P, L, P, L, L, I, P],
    "fmc_ncfhw_f32_to_cl_bf16": [P, P, I, I, I, L, I, P],
    "fmc_cl_bf16_to_ncfhw_f32": [P, L, P, I, I, I, L, P],
    "fmc_cast_act_bf16": [P, P, L, I, P],
    "fmc_timestep_embedding_bf16": [P, P, I, I, P],
    "fmc_plucker_f32": [P, P, P, I, I, I, P],
    "fmc_plucker_unshuffle_bf16": [P, P, P, I, I, I, P],
    "fmc_traj_scatter_f32": [P, P, P, P, I, I, I, I, P],
    "fmc_traj_scatter_unshuffle_bf16": [P, P, P, P, I, I, I, I, P],
    "fmc_sphere_mask_f32": [P, P, I, I, I, I, P],
    "fmc_traj_scatter_circles_unshuffle_bf16": [P, P, P, P, I, I, I, I, P],
    "fmc_mask_modulate_bf16": [P, P, P, P, P, I, I, I, I, I, I, P],
    "fmc_cfg_ddim_step_f32": [P, P, F, P, P, P, F, F, L, P],
    "fmc_window_combine_ddim_f32": [P, I, I, F, P, P, I, I, I, L, I, I, F, F, P],
    # backward (csrc/backward.cu)
    "fmc_transpose_bf16": [P, L, P, L, L, I, P],
    "fmc_colsum_f32": [P, L, I, P, P, L, I, I, P],
    "fmc_layernorm_bwd_bf16": [P, L, P, L, P, F, P, L, P, L, I, P],
    "fmc_wgrad_bf16": [P, L, P, L, P, L, P, L, I, I, I, P],
    # relative poses (csrc/pose.cu)
    "fmc_pose_relative_to_first_f64": [P, L, P, I, I, D, P],
    "fmc_pose_absolute_from_relative_f64": [P, P, P, I, I, D, P],
    "fmc_pose_objects_relative_f64": [P, L, P, L, P, I, I, D, P],
    # pipeline edges (csrc/edge.cu)
    "fmc_softmax_rows": [P, L, P, L, I, L, I, F, P],
    "fmc_small_mha": [P, L, I, I, I, P, L, I, I, I, I, I, F, I, P],
    "fmc_quick_gelu": [P, L, P, L, I, L, I, P],
    "fmc_embed_tokens": [P, P, P, P, L, I, L, I, I, I, P],
    "fmc_vae_sample_f32": [P, L, I, P, P, I, I, L, F, P],
    "fmc_cl_to_video_f32": [P, L, I, P, I, I, I, L, F, F, F, F, P],
    "fmc_groupnorm_bwd_bf16": [P, L, P, L, P, P, F, P, L, P, I, I, I, I, I, P, L, I, P],
    "fmc_geglu_fwd_bf16": [P, L, P, L, L, I, P],
    "fmc_geglu_bwd_bf16": [P, L, P, L, P, L, L, I, P],
    "fmc_relu_bwd_bf16": [P, P, P, L, P],
    "fmc_resize_nearest_bwd_bf16": [P, P, I, I, I, I, I, I, P],
    "fmc_avgpool2_bwd_bf16": [P, P, I, I, I, I, P],
    "fmc_attention_bwd_bf16": [P, L, I, P, L, I, P, L, I, I, P, L, P, L, P, L, I, P, L, I, P, L, I, P, P, I, I, I, I, I, I, I, I, I, F, P],
    "fmc_grad_norm_f32": [P, L, F, F, P, P, P],
    "fmc_adamw_step_f32": [P, P, P, P, L, F, F, F, F, F, I, P, P],
    # reference-precision mode (csrc/precise.cu)
    "fmc_gemm_tf32": [P, L, P, L, P, L, I, I, I, P, P, L, P, I, L, I, I, P],
    "fmc_split_tf32": [P, L, P, L, L, I, P],
    "fmc_attention_f32": [P, L, I, P, L, I, P, L, I, P, L, I, I, I, I, I, I, I, I, F, P],
    "fmc_layernorm_f32": [P, L, P, P, F, P, L, P, I, I, P, L, P, L, L, I, P],
    "fmc_groupnorm_f32": [P, L, P, P, F, P, L, P, I, I, I, I, I, P, L, I, P],
    "fmc_im2col3x3_f32": [P, P, I, I, I, I, I, P],
    "fmc_add_f32": [P, L, P, L, P, I, L, P, L, L, I, I, P],
    "fmc_resize_nearest_f32": [P, P, I, I, I, I, I, I, P],
    "fmc_avgpool2_f32": [P, P, I, I, I, I, P],
    "fmc_copy2d_f32": [P, L, P, L, L, I, P],
    "fmc_ncfhw_f32_to_cl_f32": [P, P, I, I, I, L, I, P],
    "fmc_cl_f32_to_ncfhw_f32": [P, L, P, I, I, I, L, P],
    "fmc_silu_f32": [P, P, L, P],
    "fmc_timestep_embedding_f32": [P, P, I, I, P],
    "fmc_mask_modulate_f32": [P, P, P, P, P, I, I, I, I, I, I, P],
}

_lib = None


class FmcError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FmcError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(make -C synfmc_b200/csrc). There is no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        handle.fmc_last_error_string.restype = c_char_p
        handle.fmc_last_error_string.argtypes = []
        handle.fmc_abi_version.restype = c_int
        handle.fmc_abi_version.argtypes = []
        handle.fmc_groupnorm_launches.restype = c_int
        handle.fmc_groupnorm_launches.argtypes = [c_int, c_int, c_int]
        handle.fmc_grad_norm_workspace_floats.restype = c_int
        handle.fmc_grad_norm_workspace_floats.argtypes = []
        handle.fmc_wgrad_workspace_floats.restype = c_longlong
        handle.fmc_wgrad_workspace_floats.argtypes = [c_longlong, c_int, c_int]
        handle.fmc_groupnorm_bwd_workspace_floats.restype = c_longlong
        handle.fmc_groupnorm_bwd_workspace_floats.argtypes = [c_int, c_int, c_int]
        handle.fmc_colsum_workspace_floats.restype = c_int
        handle.fmc_colsum_workspace_floats.argtypes = [c_longlong, c_int]
        handle.fmc_layernorm_bwd_blocks.restype = c_int
        handle.fmc_layernorm_bwd_blocks.argtypes = [c_longlong]
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export what the header declares
            fn.restype = c_int
            fn.argtypes = argtypes
        if handle.fmc_abi_version() != ABI_VERSION:  # a stale in-tree build (or FMC_B200_LIB) against newer bindings
            raise FmcError(f"{LIB_PATH} reports ABI version {handle.fmc_abi_version()}, the bindings expect {ABI_VERSION}: "
                           "rebuild with `make -C synfmc_b200/csrc`")
        _lib = handle
    return _lib


# kernels launched by one call of each entry point: everything launches one except GroupNorm, which is either the
# single-pass cluster kernel (1) or partial sums + finalize + apply (3) depending on the shape -- the library says which
def _kernels_per_call(handle, name, args):
    if name == "fmc_groupnorm_bf16":
        return handle.fmc_groupnorm_launches(args[9], args[10], args[11])  # HW, C, groups
    if name in ("fmc_groupnorm_f32", "fmc_grad_norm_f32", "fmc_colsum_f32"):
        return 2  # statistics + apply / partial sums + finalize
    if name == "fmc_groupnorm_bwd_bf16":
        return 3  # statistics, gradient means, dx
    if name == "fmc_wgrad_bf16":  # T, M, N = args[7:10]; one direct kernel when the token axis is not split, else + the fold
        splits = handle.fmc_wgrad_workspace_floats(args[7], args[8], args[9]) // (args[8] * args[9])
        return 1 if splits == 1 and not args[10] else 2
    if name == "fmc_attention_bwd_bf16":
        if args[17] and args[-7] == 16 and args[-6] == 16:
            return 1  # 16-frame self-attention: one warp per sequence
        return 2 if args[17] else 1  # dQ kernel (+ dK / dV kernel)
    return 1
launch_count = 0  # kernels of this library launched by this process (bench.py reports it as gpu_launches)
trace = None      # when set to a list by bench.py: (name, args, start_event, end_event) per call, CUDA events on the
                  # launching stream -- the per-kernel timing behind the roofline figures


def call(name, *args):
    """Invoke an entry point; a non-zero return code becomes an exception carrying the library's message."""
    global launch_count
    handle = lib()
    if trace is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(handle, name)(*args)
        e1.record()
        trace.append((name, args, e0, e1))
    else:
        rc = getattr(handle, name)(*args)
    if rc != 0:
        raise FmcError(f"{name} failed (rc={rc}): {handle.fmc_last_error_string().decode()}")
    launch_count += _kernels_per_call(handle, name, args)
