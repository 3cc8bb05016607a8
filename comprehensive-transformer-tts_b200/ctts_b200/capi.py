"""ctypes binding of libctts_b200.so (the C ABI declared in include/ctts_b200.h).

There is deliberately no fallback: if the shared library is missing, was not built for this
machine, or the device is not sm_100+, every entry point raises -- the product path must never
silently run on PyTorch eager or on the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CTTS_B200_LIB: load another build of the same ABI (A/B runs of a kernel change on one box)
LIB_PATH = os.environ.get("CTTS_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libctts_b200.so")

ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH, ACT_SWISH = 0, 1, 2, 3, 4

_P, _I, _F, _Z, _L = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong

# name -> argtypes, in the order of include/ctts_b200.h (restype is int unless noted)
SIGNATURES = {
    "ctts_abi_version": [],
    "ctts_last_error": [],
    "ctts_device_arch": [],
    "ctts_embed_tokens": [_P, _P, _P, _I, _F, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "ctts_add_positions": [_P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _P],
    "ctts_layernorm": [_P, _P, _P, _F, _P, _I, _I, _I, _P, _P],
    "ctts_conv1d_gemm": [_P, _P, _P, _F, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P],
    "ctts_pack_conv_weight": [_P, _I, _I, _I, _P, _P],
    "ctts_attention": [_P, _P, _I, _I, _I, _I, _F, _P, _P],
    "ctts_decode_durations": [_P, _F, _I, _P, _P],
    "ctts_length_scan": [_P, _P, _P, _I, _I, _P, _P, _P, _P],
    "ctts_length_expand": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P],
    "ctts_cwt_to_pitch": [_P, _I, _P, _P, _P, _I, _F, _F, _P, _I, _I, _I, _P, _P, _P, _P],
    "ctts_f0_to_pitch": [_P, _P, _I, _P, _P, _P],
    "ctts_frame_pitch": [_P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "ctts_gather_index": [_P, _P, _I, _I, _I, _P, _P],
    "ctts_phoneme_pitch": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "ctts_gather_add": [_P, _P, _I, _I, _I, _P, _P],
    "ctts_bucketize": [_P, _F, _P, _I, _I, _P, _P],
    "ctts_add_row_broadcast": [_P, _P, _I, _I, _I, _P, _P],
    "ctts_gru_bidir": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "ctts_linear_smallk": [_P, _P, _P, _P, _I, _I, _I, _P, _P],
    "ctts_aligner_attention": [_P, _P, _P, _P, _F, _I, _I, _I, _I, _P, _P, _P],
    "ctts_mas": [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "ctts_phoneme_energy": [_P, _P, _P, _I, _I, _I, _P, _P, _P],
    "ctts_batched_gemm_fp32": [_P, _P, _F, _P, _I, _I, _I, _I, _I, _I, _L, _L, _I, _L, _L, _I, _L, _L, _I, _P, _P],
    "ctts_fastformer_pool": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "ctts_binary": [_P, _P, _I, _I, _P, _I, _I, _I, _P, _P],
    "ctts_glu": [_P, _I, _I, _P, _P],
    "ctts_dwconv_bn_swish": [_P, _P, _I, _P, _P, _I, _I, _I, _P, _P],
    "ctts_relshift_softmax": [_P, _P, _I, _I, _I, _F, _P, _P],
    "ctts_relshift_softmax_planes": [_P, _P, _I, _I, _I, _I, _F, _I, _P, _P],
    "ctts_pad_heads_planes": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "ctts_transpose_heads": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "ctts_gemm_bf16x3": [_P, _P, _P, _P, _P, _F, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ctts_attention_bf16x3": [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "ctts_gemm_split": [_I, _P, _P, _P, _F, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "ctts_split_planes": [_P, _Z, _I, _P, _P],
    "ctts_layernorm_planes": [_P, _P, _P, _F, _P, _I, _I, _I, _P, _I, _P, _P],
    "ctts_gemm_split_ln": [_P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _F, _I, _P, _P, _P],
    "ctts_attention_split": [_I, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P],
    "ctts_attention_small": [_P, _P, _I, _I, _I, _I, _F, _P, _P],
    "ctts_flash_attention_bf16x3": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P],
    "ctts_transpose_v_planes": [_I, _P, _I, _I, _I, _I, _P, _P],
    "ctts_debug_set_timing_buffer": [_P],
    "ctts_split_bf16": [_P, _Z, _P, _P, _P],
    "ctts_layernorm_split": [_P, _P, _P, _F, _P, _I, _I, _I, _P, _P, _P, _P],
    # ---- training step (include/ctts_b200.h, "TRAINING STEP") ----
    "ctts_gemm_generic": [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _F, _I, _P],
    "ctts_act_bwd": [_P, _P, _I, _F, _P, _I, _I, _I, _I, _P, _P, _P],
    "ctts_act_bwd_planes": [_P, _P, _I, _F, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P],
    "ctts_dropout_add": [_P, _P, _P, _I, _I, _I, _F, ctypes.c_ulonglong, ctypes.c_ulonglong, _P, _P, _P],
    "ctts_layernorm_bwd": [_P, _P, _P, _F, _P, _I, _I, _I, _P, _I, _P, _P, _P],
    "ctts_mask_rows": [_P, _P, _I, _I, _I, _P],
    "ctts_axpy": [_P, _F, _Z, _I, _P, _P],
    "ctts_rowscale_axpy": [_P, _P, _F, _I, _I, _I, _P, _P],
    "ctts_scatter_add_rows": [_P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P],
    "ctts_length_expand_bwd": [_P, _P, _I, _I, _I, _I, _I, _P, _P],
    "ctts_add_positions_bwd": [_P, _P, _P, _I, _P, _I, _I, _I, _I, _P, _P],
    "ctts_masked_softmax": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "ctts_softmax_bwd": [_P, _P, _I, _I, _I, _I, _F, _P, _P],
    "ctts_bn_stats": [_P, _I, _I, _P, _P, _P],
    "ctts_bn_act_fwd": [_P, _P, _P, _P, _P, _F, _I, _I, _I, _P, _I, _P, _P],
    "ctts_bn_update_running": [_P, _P, _I, _F, _I, _P, _P, _P, _P],
    "ctts_bn_bwd": [_P, _P, _P, _P, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P, _P],
    "ctts_dropout": [_P, _Z, _F, ctypes.c_ulonglong, ctypes.c_ulonglong, _P, _P, _P],
    "ctts_pack_conv_weight_dgrad": [_P, _I, _I, _I, _P, _P],
    "ctts_unpack_conv_wgrad": [_P, _I, _I, _I, _I, _P, _P],
    "ctts_split_transpose": [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "ctts_gemm_wgrad": [_I, _P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _P, _P],
    "ctts_gemm_wgrad_rowmajor": [_I, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P, _P],
    "ctts_gemm_batched_planes": [_I, _P, _P, _P, _P, _P, _L, _L, _F, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "ctts_aligner_attention_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P],
    "ctts_glu_bwd": [_P, _P, _I, _I, _P, _P],
    "ctts_dwconv": [_P, _P, _I, _I, _I, _I, _P, _P],
    "ctts_dwconv_bwd": [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "ctts_relshift_bwd": [_P, _I, _I, _I, _I, _F, _P, _P, _P],
    "ctts_fastformer_pool_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P],
    "ctts_mul_bwd": [_P, _P, _P, _I, _P, _I, _I, _I, _P, _P, _P],
    "ctts_act_fwd": [_P, _Z, _I, _P, _I, _P, _P],
    "ctts_merge_planes": [_I, _P, _Z, _P, _P],
    "ctts_copy_rows": [_P, _L, _I, _I, _P, _L, _I, _P],
    "ctts_add_coords": [_P, _I, _I, _I, _P, _P],
    "ctts_im2col_3x3_s12": [_P, _I, _I, _I, _I, _P, _P],
    "ctts_col2im_3x3_s12": [_P, _I, _I, _I, _I, _P, _P],
    "ctts_permute_last2": [_P, _I, _I, _I, _P, _P],
    "ctts_gru_bwd": [_P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P],
}

_lib = None


class CttsError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and set the prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CttsError("libctts_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / eager fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = ctypes.c_char_p if name == "ctts_last_error" else ctypes.c_int
    if lib.ctts_abi_version() != 1:
        raise CttsError("libctts_b200.so ABI %d, binding expects 1" % lib.ctts_abi_version())
    _lib = lib
    return lib


def ptr_array(tensors):
    """Host array of device pointers (for the `const void* const*` plane arguments)."""
    arr = (ctypes.c_void_p * 3)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    arr._keepalive = list(tensors)
    return arr


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


LAUNCHES = 0          # number of kernel-launching entry points called so far (bench.py reports the delta)
TIMING_HOOK = None    # optional callable(name, args) -> context manager, used by bench.py to time one op class


def call(name, *args):
    """Call an entry point with torch tensors (-> device pointers) / python scalars; raise on error."""
    global LAUNCHES
    lib = load()
    LAUNCHES += 1
    conv = [(_ptr(a) if (a is None or hasattr(a, "data_ptr")) else (ctypes.cast(a, ctypes.c_void_p) if isinstance(
        a, ctypes.Array) else a)) for a in args]
    rc = getattr(lib, name)(*conv)
    if rc != 0:
        raise CttsError("%s failed (%d): %s" % (name, rc, lib.ctts_last_error().decode()))


def require_cuda_tensor(t):
    """The product path has no CPU implementation: refuse host tensors instead of computing somewhere else."""
    if not t.is_cuda:
        raise CttsError("CompTransTTS (B200) runs on CUDA tensors only; got %s -- there is no CPU path" % t.device)


_arch_checked = False
_device_seen = None


def require_device():
    """Fail loudly unless a CUDA device of compute capability >= 10.0 is current -- and the SAME device as before: the
    library keeps per-process state that is really per-device (raised shared-memory limits, SM count, cluster occupancy),
    so the model is one process per GPU (train.py:251-252 does the same with mp.spawn)."""
    global _arch_checked, _device_seen
    if _arch_checked:
        import torch
        cur = torch.cuda.current_device()
        if cur != _device_seen:
            raise CttsError("libctts_b200 was initialised on cuda:%d and is now called on cuda:%d; use one process per GPU"
                            % (_device_seen, cur))
        return
    lib = load()
    arch = lib.ctts_device_arch()
    if arch < 0:
        raise CttsError("no usable CUDA device: %s" % lib.ctts_last_error().decode())
    if arch < 100:
        raise CttsError("libctts_b200 is built for sm_100a only; current device is sm_%d" % arch)
    import torch
    _device_seen = torch.cuda.current_device()
    _arch_checked = True
