"""ctypes loader for libopadpo_b200.so (the C ABI declared in include/opadpo_b200.h).

The product path has no CPU or PyTorch fallback: if the CUDA library is missing, loading raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libopadpo_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

HASH_PATH = LIB_PATH + ".srchash"
EXPECTED_ABI = 7     # opadpo_abi_version() of the library these bindings were written against

_lib = None


def source_hash() -> str:
    """sha256 over every file the library is compiled from (csrc/* and the public header)."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if not f.startswith("."))
    files.append(os.path.join(os.path.dirname(_HERE), "include", "opadpo_b200.h"))
    for p in files:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


class OpadpoError(RuntimeError):
    pass


class OpadpoUnknownName(OpadpoError):
    """OPADPO_ENOENT: the engine holds no tensor / module of that name (as opposed to a shape or dtype error on one it does)."""


def build_library(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libopadpo_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(max(1, (os.cpu_count() or 2)))]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise OpadpoError("building libopadpo_b200.so failed (see output above)")
    with open(HASH_PATH, "w") as f:
        f.write(source_hash() + "\n")
    return LIB_PATH


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K1", ctypes.c_int32), ("K2", ctypes.c_int32),
        ("a_mn_major", ctypes.c_int32), ("b_mn_major", ctypes.c_int32),
        ("A1", ctypes.c_void_p), ("lda1", ctypes.c_int64),
        ("B1", ctypes.c_void_p), ("ldb1", ctypes.c_int64),
        ("A2", ctypes.c_void_p), ("lda2", ctypes.c_int64),
        ("B2", ctypes.c_void_p), ("ldb2", ctypes.c_int64),
        ("a2_group_n", ctypes.c_int32), ("a2_group_stride", ctypes.c_int32),
        ("splits", ctypes.c_int32), ("epilogue", ctypes.c_int32),
        ("alpha", ctypes.c_float),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_int64),
        ("bias", ctypes.c_void_p),
        ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int64),
        ("act", ctypes.c_int32),
        ("labels", ctypes.c_void_p),
        ("lse_partials", ctypes.c_void_p),
        ("label_logit", ctypes.c_void_p),
        ("lse", ctypes.c_void_p),
        ("dlp", ctypes.c_void_p),
        ("max_ctas", ctypes.c_int32),
        ("cluster_mode", ctypes.c_int32),
        ("dent", ctypes.c_void_p),
        ("ent", ctypes.c_void_p),
        ("rope_pos", ctypes.c_void_p),
        ("rope_cs", ctypes.c_void_p),
        ("rope_cols", ctypes.c_int32), ("rope_S", ctypes.c_int32),
    ]


class AttnDesc(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int32), ("S", ctypes.c_int32), ("H", ctypes.c_int32), ("HD", ctypes.c_int32),
        ("causal", ctypes.c_int32), ("scale", ctypes.c_float),
        ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("ld_qkv", ctypes.c_int64),
        ("o", ctypes.c_void_p), ("ld_o", ctypes.c_int64),
        ("lse", ctypes.c_void_p), ("kbits", ctypes.c_void_p),
        ("dO", ctypes.c_void_p), ("ld_do", ctypes.c_int64),
        ("delta", ctypes.c_void_p),
        ("dq", ctypes.c_void_p), ("dk", ctypes.c_void_p), ("dv", ctypes.c_void_p), ("ld_dqkv", ctypes.c_int64),
        ("prefix_len", ctypes.c_int32), ("segment_len", ctypes.c_int32),
    ]


LAYER_CB = ctypes.CFUNCTYPE(None, ctypes.c_int32, ctypes.c_void_p)


class DpoPairDesc(ctypes.Structure):
    _fields_ = [
        ("pi_c", ctypes.c_void_p), ("pi_r", ctypes.c_void_p), ("ref_c", ctypes.c_void_p), ("ref_r", ctypes.c_void_p),
        ("s_c", ctypes.c_void_p), ("s_r", ctypes.c_void_p),
        ("d_c", ctypes.c_void_p), ("d_r", ctypes.c_void_p),
        ("out", ctypes.c_void_p),
        ("n", ctypes.c_int64),
        ("beta", ctypes.c_float), ("label_smoothing", ctypes.c_float),
        ("reference_free", ctypes.c_int32), ("divergence", ctypes.c_int32),
        ("alpha_div", ctypes.c_float), ("loss_coef", ctypes.c_float),
        ("anchor_mode", ctypes.c_int32),
        ("anchor_sign", ctypes.c_float), ("anchor_value", ctypes.c_float), ("anchor_coef", ctypes.c_float),
        ("grad_scale", ctypes.c_float),
    ]


class Segments(ctypes.Structure):
    _fields_ = [("tokens", ctypes.c_void_p), ("seg_offsets", ctypes.c_void_p), ("seg_score", ctypes.c_void_p),
                ("seg_relation", ctypes.c_void_p), ("seg_flags", ctypes.c_void_p), ("row_offsets", ctypes.c_void_p),
                ("b", ctypes.c_int32)]


class Sampler(ctypes.Structure):
    _fields_ = [("do_sample", ctypes.c_int32), ("top_k", ctypes.c_int32), ("top_p", ctypes.c_float),
                ("temperature", ctypes.c_float), ("seed", ctypes.c_uint64), ("n_suppress", ctypes.c_int32),
                ("suppress_tokens", ctypes.c_int32 * 8)]


MAX_RANKS = 8            # OPADPO_MAX_RANKS
AR_MAX_BLOCKS = 256      # OPADPO_AR_MAX_BLOCKS


class AllReduceDesc(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32),
                ("grad", ctypes.c_void_p), ("stage_local", ctypes.c_void_p), ("stage_mc", ctypes.c_uint64),
                ("stage_peer", ctypes.c_uint64 * MAX_RANKS), ("signal_peer", ctypes.c_uint64 * MAX_RANKS),
                ("count", ctypes.c_int64), ("epoch", ctypes.c_uint32), ("blocks", ctypes.c_int32),
                ("err_flag", ctypes.c_void_p), ("timeout_ms", ctypes.c_int32), ("fp32_in_place", ctypes.c_int32)]


class Config(ctypes.Structure):
    _fields_ = [
        ("d_model", ctypes.c_int32), ("n_layers", ctypes.c_int32), ("n_heads", ctypes.c_int32),
        ("d_ffn", ctypes.c_int32), ("vocab", ctypes.c_int32),
        ("rms_eps", ctypes.c_float), ("rope_theta", ctypes.c_float),
        ("max_positions", ctypes.c_int32),
        ("v_d", ctypes.c_int32), ("v_layers", ctypes.c_int32), ("v_heads", ctypes.c_int32),
        ("v_ffn", ctypes.c_int32), ("v_image", ctypes.c_int32), ("v_patch", ctypes.c_int32),
        ("v_eps", ctypes.c_float),
        ("lora_r", ctypes.c_int32), ("lora_alpha", ctypes.c_float),
        ("pad_id", ctypes.c_int32), ("eos_id", ctypes.c_int32), ("image_token", ctypes.c_int32),
    ]


# every symbol include/opadpo_b200.h declares (tests check the built library exports all of them)
EXPORTED_SYMBOLS = [
    "opadpo_last_error", "opadpo_abi_version", "opadpo_op_gemm", "opadpo_gemm_n_tiles", "opadpo_op_decode_gemm", "opadpo_op_attn_fwd",
    "opadpo_op_attn_bwd", "opadpo_op_pack_mask", "opadpo_op_dpo_pair", "opadpo_engine_create", "opadpo_engine_destroy",
    "opadpo_engine_load_tensor", "opadpo_engine_load_lora", "opadpo_engine_export_lora", "opadpo_engine_lora_shape",
    "opadpo_engine_set_trainable", "opadpo_engine_buffers", "opadpo_engine_lora_offsets", "opadpo_engine_sync_shadow",
    "opadpo_engine_set_option", "opadpo_engine_set_backward_callback", "opadpo_engine_layer_grad_range",
    "opadpo_engine_adopt_grad_buffer",
    "opadpo_forward_logprobs", "opadpo_backward", "opadpo_backward_ent", "opadpo_zero_grad", "opadpo_optimizer_step", "opadpo_generate",
    "opadpo_launch_count", "opadpo_prof_begin", "opadpo_prof_end",
    "opadpo_op_collate_responses", "opadpo_op_collate_queries", "opadpo_op_mask_images", "opadpo_op_attention_holes",
    "opadpo_op_preprocess_images", "opadpo_op_truncate_after_eos", "opadpo_op_allreduce_bf16wire",
]


def _declare_engine(lib):
    c = ctypes
    vp, i32, i64, f32 = c.c_void_p, c.c_int32, c.c_int64, c.c_float
    lib.opadpo_op_dpo_pair.argtypes = [c.POINTER(DpoPairDesc), vp]
    lib.opadpo_engine_create.argtypes = [c.POINTER(Config), c.POINTER(vp)]
    lib.opadpo_engine_destroy.argtypes = [vp]
    lib.opadpo_engine_load_tensor.argtypes = [vp, c.c_char_p, vp, i32, i64, i64, vp]
    lib.opadpo_engine_load_lora.argtypes = [vp, i32, c.c_char_p, vp, vp, i32, vp]
    lib.opadpo_engine_export_lora.argtypes = [vp, i32, c.c_char_p, vp, vp, vp]
    lib.opadpo_engine_lora_shape.argtypes = [vp, c.c_char_p, c.POINTER(i64), c.POINTER(i64)]
    lib.opadpo_engine_set_trainable.argtypes = [vp, i32]
    lib.opadpo_engine_buffers.argtypes = [vp, c.POINTER(vp), c.POINTER(vp), c.POINTER(vp), c.POINTER(vp),
                                          c.POINTER(i64), c.POINTER(i64)]
    lib.opadpo_engine_lora_offsets.argtypes = [vp, c.c_char_p, c.POINTER(i64), c.POINTER(i64)]
    lib.opadpo_engine_sync_shadow.argtypes = [vp, vp]
    lib.opadpo_engine_set_option.argtypes = [vp, c.c_char_p, i64]
    lib.opadpo_engine_set_backward_callback.argtypes = [vp, LAYER_CB, vp]
    lib.opadpo_engine_layer_grad_range.argtypes = [vp, i32, c.POINTER(i64), c.POINTER(i64)]
    lib.opadpo_engine_adopt_grad_buffer.argtypes = [vp, vp, i64]
    lib.opadpo_forward_logprobs.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, i32, c.POINTER(vp), i32, f32, vp, vp,
                                            i32, vp]
    lib.opadpo_backward.argtypes = [vp, i32, vp, vp]
    lib.opadpo_backward_ent.argtypes = [vp, i32, vp, vp, vp]
    lib.opadpo_zero_grad.argtypes = [vp, vp]
    lib.opadpo_optimizer_step.argtypes = [vp, f32, f32, f32, f32, f32, f32, f32, vp, vp]
    lib.opadpo_generate.argtypes = [vp, i32, vp, i32, vp, vp, i32, i32, c.POINTER(Sampler), vp, vp]
    lib.opadpo_prof_begin.argtypes = []
    lib.opadpo_op_collate_responses.argtypes = [c.POINTER(Segments), i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.opadpo_op_collate_queries.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.opadpo_op_mask_images.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp]
    lib.opadpo_op_attention_holes.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.opadpo_op_truncate_after_eos.argtypes = [vp, vp, i32, i32, i64, i64, vp, i32, vp]
    lib.opadpo_op_preprocess_images.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, vp, vp]
    lib.opadpo_op_allreduce_bf16wire.argtypes = [c.POINTER(AllReduceDesc), vp]
    lib.opadpo_prof_end.argtypes = [c.POINTER(c.c_double), c.POINTER(c.c_double), c.POINTER(i64)]
    lib.opadpo_launch_count.argtypes = []
    lib.opadpo_launch_count.restype = c.c_int64
    for name in EXPORTED_SYMBOLS:
        if name not in ("opadpo_last_error", "opadpo_launch_count"):
            getattr(lib, name).restype = c.c_int


def _declare(lib):
    _declare_engine(lib)
    lib.opadpo_last_error.restype = ctypes.c_char_p
    lib.opadpo_abi_version.restype = ctypes.c_int
    lib.opadpo_gemm_n_tiles.argtypes = [ctypes.c_int32]
    lib.opadpo_gemm_n_tiles.restype = ctypes.c_int
    lib.opadpo_op_gemm.argtypes = [ctypes.POINTER(GemmDesc), ctypes.c_void_p]
    lib.opadpo_op_gemm.restype = ctypes.c_int
    lib.opadpo_op_decode_gemm.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                                          ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_float,
                                          ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_void_p]
    lib.opadpo_op_decode_gemm.restype = ctypes.c_int
    for name in ("opadpo_op_attn_fwd", "opadpo_op_attn_bwd"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.POINTER(AttnDesc), ctypes.c_void_p]
        fn.restype = ctypes.c_int
    lib.opadpo_op_pack_mask.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.opadpo_op_pack_mask.restype = ctypes.c_int


def load(build_if_missing: bool = False):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            build_library()
        else:
            raise OpadpoError(
                f"{LIB_PATH} is missing: the CUDA extension is the only implementation of the hot path "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`)")
    _check_fresh()
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
    lib.opadpo_abi_version.restype = ctypes.c_int
    got = lib.opadpo_abi_version()
    if got != EXPECTED_ABI:
        raise OpadpoError(f"{LIB_PATH} exports ABI version {got}, these bindings expect {EXPECTED_ABI}: rebuild it "
                          "(`python -c 'import __graft_entry__ as g; g.build()'`)")
    _declare(lib)
    _lib = lib
    return lib


def _check_fresh():
    """A library built from other sources than the ones beside it would be loaded silently and run stale kernels under new
    struct layouts: refuse. Compared by content hash (build_library records it), so copies and checkouts that reset
    mtimes do not matter. OPADPO_ALLOW_STALE_LIB=1 overrides."""
    try:
        with open(HASH_PATH) as f:
            built = f.read().strip()
    except OSError:
        built = None
    if built != source_hash() and os.environ.get("OPADPO_ALLOW_STALE_LIB") != "1":
        raise OpadpoError(f"{LIB_PATH} was not built from the sources in {CSRC_DIR} (source hash mismatch): rebuild it "
                          "(`python -c 'import __graft_entry__ as g; g.build()'`), or set OPADPO_ALLOW_STALE_LIB=1")


ENOENT = -5     # OPADPO_ENOENT


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().opadpo_last_error()
        cls = OpadpoUnknownName if rc == ENOENT else OpadpoError
        raise cls(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
