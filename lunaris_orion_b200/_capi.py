"""ctypes binding of include/lunaris_b200.h. Fails loudly when the library is missing (no fallback)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LUNARIS_B200_LIB selects another build of the SAME library (tools/gpu A/B runs of a kernel change against its
# predecessor on one box); it is not a fallback: a missing file raises like the default path does.
LIB_PATH = os.environ.get("LUNARIS_B200_LIB") or os.path.join(_HERE, "_lib", "liblunaris_b200.so")

_ERRORS = {
    2: "LUN_E_SHAPE: unsupported channel count / block size",
    3: "LUN_E_TAPS: tap list empty or longer than 16",
    4: "LUN_E_GRID: spatial grid is not a power of two",
    5: "LUN_E_BOX: TMA box would exceed 256 elements",
    6: "LUN_E_STATS", 7: "LUN_E_ALIGN: output channel stride/offset not 16-byte aligned",
    8: "LUN_E_ATTR: cudaFuncSetAttribute failed", 9: "LUN_E_LAUNCH: kernel launch failed",
    101: "LUN_E_DRIVER: cuTensorMapEncodeTiled unavailable", 102: "LUN_E_TMAP: tensor-map encode failed",
    103: "LUN_E_TMAP: tensor-map encode failed (2d)",
}


class LunarisB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load liblunaris_b200.so once. Raises if it was not built (python -m lunaris_orion_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LunarisB200Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "lunaris_orion_b200 has no CPU / library fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


c_int, c_float, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p
c_long, c_double, c_u64 = ctypes.c_long, ctypes.c_double, ctypes.c_ulonglong
c_pp = ctypes.POINTER(ctypes.c_void_p)
c_int_p = ctypes.POINTER(ctypes.c_int)

# name -> argtypes; every function returns int. Kept in sync with include/lunaris_b200.h
# (tests/test_capi_symbols.py parses the header and checks this table against it).
SIGNATURES = {
    "lun_num_sms": [],
    "lun_abi_version": [],
    "lun_launch_count": [],
    "lun_conv_taps_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                           c_int, c_int, c_int, c_int, c_int, c_int_p, c_int_p, c_int_p,
                           c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_int, c_float, c_void_p, c_void_p],
    "lun_convT4x4s2_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                            c_void_p],
    "lun_convT4x4s2_halo_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p],
    "lun_conv_taps_dropsum_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_int_p, c_int_p, c_int_p,
                                   c_void_p, c_int, c_int, c_int, c_void_p, c_u64, c_float, c_void_p],
    "lun_wgrad_taps_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int_p, c_int_p, c_int_p, c_void_p, c_void_p],
    "lun_channel_stats_bf16": [c_void_p, c_long, c_int, c_void_p, c_void_p],
    "lun_bn_finalize": [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float,
                        c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "lun_affine_fwd_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_u64, c_float, c_int, c_int, c_int, c_float, c_void_p],
    "lun_block_bwd_reduce_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p],
    "lun_block_bwd_apply_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                 c_float, c_void_p],
    "lun_attn_ref_rows_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_float, c_void_p],
    "lun_attn_ref_rows_split_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_float,
                                     c_void_p],
    "lun_gather_query_rows_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "lun_attn_fold_rows_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_u64, c_float, c_void_p],
    "lun_gather_query_rows_affine_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_int, c_void_p],
    "lun_proj_expand_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_float,
                             c_void_p],
    "lun_proj_bwd_gather_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_u64, c_float,
                                 c_void_p],
    "lun_fe_conv1": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p],
    "lun_image_channel_stats_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "lun_gn_mish_fwd_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                             c_int, c_float, c_void_p],
    "lun_gn_mish_bwd_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "lun_conv3x3_c3_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "lun_conv3x3_c3_wgrad": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "lun_gn_mish_final_conv_tanh_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int, c_int, c_int, c_int, c_float, c_void_p],
    "lun_final_conv_tanh_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "lun_final_conv_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                           c_void_p],
    "lun_reparam_fwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "lun_reparam_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "lun_vae_loss_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_void_p],
    "lun_vae_loss_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_void_p],
    "lun_sprites_u8_to_f32": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "lun_flash_attn2d_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                              c_void_p, c_void_p],
    "lun_flash_attn2d_bwd_prep_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_void_p],
    "lun_flash_attn2d_dv_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "lun_flash_attn2d_dqk_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_void_p],
    "lun_pack_weight_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "lun_head_linear_bf16": [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "lun_heads_buffer_sizes": [c_int_p, c_int_p],
    "lun_heads_fwd": [c_pp, c_int_p, ctypes.POINTER(c_u64), c_float, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "lun_heads_bwd": [c_pp, c_int_p, ctypes.POINTER(c_u64), c_float, c_float, c_float, c_void_p, c_void_p, c_void_p,
                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                      c_pp, c_void_p],
    "lun_multi_grad_sumsq": [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p],
    "lun_multi_clip_adamw": [c_void_p, c_void_p, c_int, c_void_p, c_float, c_float, c_void_p],
    "lun_fe_branches": [c_void_p, c_void_p, c_void_p, c_pp, c_pp, c_pp, c_pp, c_void_p, c_int, c_int, c_int, c_float,
                        c_void_p],
}


def _declare(l):
    older_build = bool(os.environ.get("LUNARIS_B200_LIB"))      # an A/B build may predate the newest entry points
    for name, argtypes in SIGNATURES.items():
        if older_build and not hasattr(l, name):
            continue
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_longlong if name == "lun_launch_count" else c_int


def check(rc, what):
    if rc != 0:
        raise LunarisB200Error(f"{what} failed with code {rc}: {_ERRORS.get(rc, 'unknown error')}")


_int_arrays = {}


def int_array(values):
    """ctypes int array of a (short, recurring) Python sequence - memoised: the tap lists of the ~20 conv geometries
    recur on every launch and building a ctypes array costs more than the launch call itself."""
    key = tuple(values)
    arr = _int_arrays.get(key)
    if arr is None:
        arr = _int_arrays[key] = (ctypes.c_int * len(key))(*key)
    return arr


def raw_stream():
    """cudaStream_t of torch's current stream on the current device (two C calls; torch.cuda.current_stream() builds a
    Python Stream object and resolves the device by name each time - 15 us, a fifth of the host time of a step)."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
