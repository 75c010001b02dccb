"""ctypes binding of libvtb200.so (C-ABI declared in include/vtb200.h).

The library is the product: there is NO CPU / eager fallback.  Importing this module never needs a GPU
(so the `-m "not gpu"` suite can check that the library loads and exports every symbol); calling any
op without CUDA raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VTB_LIB=<file name next to this module>: A/B of two builds of the library on the same box (tools/gpu_ab_lib.sh)
LIB_PATH = os.path.join(_HERE, os.environ.get("VTB_LIB", "libvtb200.so"))

EPI_NONE, EPI_SILU_DUAL, EPI_SILU_GRAD = 0, 1, 2
ATTN_GLOBAL, ATTN_WINDOW, ATTN_HALO = 0, 1, 2

# every symbol include/vtb200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "vtb_last_error", "vtb_version", "vtb_init", "vtb_set_option", "vtb_gemm_bf16", "vtb_layernorm_fwd",
    "vtb_layernorm_bwd", "vtb_attention_fwd", "vtb_attention_bwd", "vtb_attention_bwd_workspace_bytes", "vtb_cast_f32_bf16",
    "vtb_cast_f32_bf16_2d", "vtb_scale_cast_bf16", "vtb_scale_cast_colsum_bf16", "vtb_colsum_bf16", "vtb_dropout", "vtb_patch_gather",
    "vtb_patch_scatter", "vtb_transpose_hw", "vtb_dwconv3x3_fwd", "vtb_dwconv3x3_bwd", "vtb_vit_assemble_tokens", "vtb_fill_rows", "vtb_rowgroup_sum", "vtb_mean_rows_fwd", "vtb_mean_rows_bwd",
    "vtb_silu_fwd", "vtb_silu_bwd", "vtb_dino_loss", "vtb_mt_num_chunks", "vtb_mt_cast_f32_bf16", "vtb_mt_ema",
    "vtb_mt_grad_norm", "vtb_mt_scale", "vtb_mt_agc", "vtb_mt_adamw", "vtb_mix_loss", "vtb_l2norm_fwd", "vtb_l2norm_bwd",
    "vtb_weight_norm_fwd", "vtb_weight_norm_bwd", "vtb_gelu_fwd", "vtb_gelu_bwd", "vtb_input_batch",
    "vtb_split3_bf16", "vtb_attention_fwd_f32", "vtb_silu_fwd_exact", "vtb_patch_gather_f32",
]


class GemmParams(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int32), ("a_mn_major", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int32), ("b_mn_major", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int32), ("out_f32", C.c_int32),
        ("out2", C.c_void_p),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int32),
        ("row_scale", C.c_void_p), ("rows_per_scale", C.c_int32),
        ("aux", C.c_void_p), ("ldaux", C.c_int32),
        ("epilogue", C.c_int32), ("splits", C.c_int32), ("accumulate", C.c_int32),
        ("alpha", C.c_float),
        ("a_colsum", C.c_void_p),
    ]


class AttnParams(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("batch", C.c_int32), ("heads", C.c_int32), ("dh", C.c_int32),
        ("nq", C.c_int32), ("nkv", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32),
        ("window", C.c_int32), ("shift", C.c_int32), ("halo", C.c_int32),
        ("scale", C.c_float),
        ("q", C.c_void_p), ("ldq", C.c_int32),
        ("k", C.c_void_p), ("ldk", C.c_int32),
        ("v", C.c_void_p), ("ldv", C.c_int32),
        ("o", C.c_void_p), ("ldo", C.c_int32),
        ("lse", C.c_void_p),
        ("rel_bias", C.c_void_p),
        ("pos", C.c_void_p), ("n_pos", C.c_int32),
        ("mask", C.c_void_p), ("n_mask", C.c_int32), ("mask_ld", C.c_int32),
        ("dout", C.c_void_p), ("lddo", C.c_int32),
        ("dq", C.c_void_p), ("lddq", C.c_int32),
        ("dk", C.c_void_p), ("lddk", C.c_int32),
        ("dv", C.c_void_p), ("lddv", C.c_int32),
        ("dkv_f32", C.c_int32),
        ("delta", C.c_void_p),
        ("drel_bias", C.c_void_p),
        ("mask_bits", C.c_void_p),
        ("ws", C.c_void_p), ("ws_bytes", C.c_int64),
    ]


_lib = None
_inited = False


def load():
    """dlopen the library (no GPU needed) and set prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"vtb200: {LIB_PATH} is missing — build it with `python __graft_entry__.py build` "
            "(or `make -C vision-transformers-pytorch_b200/csrc`). There is no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.vtb_last_error.restype = C.c_char_p
    lib.vtb_last_error.argtypes = []
    lib.vtb_version.restype = i32
    lib.vtb_init.restype = i32
    lib.vtb_set_option.argtypes = [C.c_char_p, i32]
    lib.vtb_gemm_bf16.argtypes = [C.POINTER(GemmParams), vp]
    lib.vtb_layernorm_fwd.argtypes = [vp, vp, vp, f32, i64, i32, i32, i32, i32, vp, i32, vp, vp, vp, i32, vp]
    lib.vtb_layernorm_bwd.argtypes = [vp, i32, vp, vp, vp, vp, i64, i32, i32, i32, i32, vp, vp, vp, vp,
                                      i32, vp, vp, vp, vp]
    lib.vtb_attention_fwd.argtypes = [C.POINTER(AttnParams), vp]
    lib.vtb_attention_bwd.argtypes = [C.POINTER(AttnParams), vp]
    lib.vtb_attention_bwd_workspace_bytes.argtypes = [C.POINTER(AttnParams)]
    lib.vtb_attention_bwd_workspace_bytes.restype = i64
    lib.vtb_attention_fwd_f32.argtypes = [C.POINTER(AttnParams), vp]
    lib.vtb_split3_bf16.argtypes = [vp, i64, i64, i32, i32, vp, vp]
    lib.vtb_silu_fwd_exact.argtypes = [vp, vp, i64, vp]
    lib.vtb_patch_gather_f32.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.vtb_cast_f32_bf16.argtypes = [vp, vp, i64, vp]
    lib.vtb_cast_f32_bf16_2d.argtypes = [vp, i64, vp, i64, i64, i32, vp]
    lib.vtb_scale_cast_bf16.argtypes = [vp, vp, i32, i64, i32, vp, vp]
    lib.vtb_scale_cast_colsum_bf16.argtypes = [vp, vp, i32, i64, i32, vp, vp, vp]
    lib.vtb_colsum_bf16.argtypes = [vp, i64, i32, i32, vp, vp]
    lib.vtb_dropout.argtypes = [vp, vp, f32, i64, i32, vp, vp, i64, vp, vp]
    lib.vtb_patch_gather.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.vtb_patch_scatter.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, vp]
    lib.vtb_transpose_hw.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.vtb_dwconv3x3_fwd.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    lib.vtb_dwconv3x3_bwd.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.vtb_vit_assemble_tokens.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    lib.vtb_fill_rows.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    lib.vtb_rowgroup_sum.argtypes = [vp, i64, i32, i32, i32, vp, vp]
    lib.vtb_mean_rows_fwd.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.vtb_mean_rows_bwd.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.vtb_silu_fwd.argtypes = [vp, vp, i64, vp]
    lib.vtb_silu_bwd.argtypes = [vp, vp, vp, i64, vp]
    lib.vtb_dino_loss.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, vp, vp, vp]
    f64 = C.c_double
    lib.vtb_l2norm_fwd.argtypes = [vp, i64, i32, f32, vp, vp, vp]
    lib.vtb_l2norm_bwd.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.vtb_weight_norm_fwd.argtypes = [vp, vp, i64, i32, vp, vp, vp]
    lib.vtb_weight_norm_bwd.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
    lib.vtb_gelu_fwd.argtypes = [vp, vp, vp, i64, vp]
    lib.vtb_gelu_bwd.argtypes = [vp, vp, vp, i64, vp]
    lib.vtb_mt_num_chunks.argtypes = [vp, i32]
    lib.vtb_mt_cast_f32_bf16.argtypes = [vp, vp, vp, i32, vp]
    lib.vtb_mt_ema.argtypes = [vp, vp, vp, i32, f64, vp]
    lib.vtb_mt_grad_norm.argtypes = [vp, vp, i32, f32, vp, vp, vp]
    lib.vtb_mt_scale.argtypes = [vp, vp, i32, vp, vp]
    lib.vtb_mt_agc.argtypes = [vp, vp, vp, vp, i32, f32, f32, vp]
    lib.vtb_mt_adamw.argtypes = [vp, vp, vp, vp, vp, vp, i32, f64, f64, f64, f64, f64, i64, vp, vp]
    lib.vtb_mix_loss.argtypes = [vp, i64, vp, vp, vp, i32, i32, f64, f32, vp, vp, vp, vp, i32, vp]
    lib.vtb_input_batch.argtypes = [vp, i32, vp, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("vtb_last_error",):
            fn.restype = i32
    lib.vtb_mt_num_chunks.restype = i64
    _lib = lib
    return lib


def get():
    """Library handle, initialised for the current CUDA device.  Raises without a GPU."""
    global _inited
    lib = load()
    if not _inited:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError(
                "vtb200: CUDA device required — the sm_100a kernels are the only implementation "
                "(no CPU fallback by design)."
            )
        torch.cuda.current_device()  # make sure a context exists
        torch.zeros(1, device="cuda")
        check(lib.vtb_init(), lib)
        _inited = True
        # A/B switches from the environment, e.g. VTB_OPTS=attn_tc_fwd_version=1,attn_tc_bwd_version=1
        for kv in filter(None, os.environ.get("VTB_OPTS", "").split(",")):
            name, _, val = kv.partition("=")
            check(lib.vtb_set_option(name.strip().encode(), int(val)), lib)
    return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        msg = lib.vtb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"vtb200 error {rc}: {msg}")


def set_option(name, value):
    lib = get()
    check(lib.vtb_set_option(name.encode(), int(value)), lib)
