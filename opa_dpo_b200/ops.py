"""Operator-level bindings: torch tensors in, C ABI calls out (device pointers + sizes only cross the boundary).

These wrap `opadpo_op_*` one-to-one; they are what tests/ uses for kernel parity and what the host-side mirror of
the reference interface composes. No function here computes anything in PyTorch.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import AttnDesc, GemmDesc

EPI_BF16, EPI_F32_ATOMIC, EPI_LSE, EPI_CEGRAD, EPI_F32_PARTIAL, EPI_BF16_ROPE, EPI_SWIGLU_BF16, EPI_SWIGLU_BWD = 0, 1, 2, 3, 4, 5, 6, 7
ACT_NONE, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk2d(t: torch.Tensor, dtype=torch.bfloat16):
    assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1 and t.dtype == dtype, (t.shape, t.stride(), t.dtype)


def gemm_n_tiles(n: int) -> int:
    return _lib.load().opadpo_gemm_n_tiles(n)


def gemm(a1: torch.Tensor, b1: torch.Tensor, out: torch.Tensor, *, a2=None, b2=None, a_mn=False, b_mn=False,
         alpha: float = 1.0, epilogue: int = EPI_BF16, bias=None, residual=None, act: int = ACT_NONE,
         a2_group_n: int = 0, a2_group_stride: int = 0, splits: int = 1, labels=None, lse_partials=None,
         label_logit=None, lse=None, dlp=None, dent=None, ent=None, rope_pos=None, rope_cs=None, rope_cols: int = 0,
         rope_S: int = 0, M=None, N=None, K1=None, K2=None, max_ctas: int = 0, cluster: int = -1):
    """out[M,N] = epilogue(alpha*(A1 B1^T + A2 B2^T)). Operand shapes follow their 'major':
    K-major A is [M,K], MN-major A is [K,M]; K-major B is [N,K], MN-major B is [K,N]."""
    lib = _lib.load()
    _chk2d(a1), _chk2d(b1)
    if M is None:
        M = a1.shape[1] if a_mn else a1.shape[0]
    if K1 is None:
        K1 = a1.shape[0] if a_mn else a1.shape[1]
    if N is None:
        N = b1.shape[1] if b_mn else b1.shape[0]
    d = GemmDesc()
    d.M, d.N, d.K1 = M, N, K1
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.A1, d.lda1 = a1.data_ptr(), a1.stride(0)
    d.B1, d.ldb1 = b1.data_ptr(), b1.stride(0)
    if a2 is not None:
        _chk2d(a2), _chk2d(b2)
        if K2 is None:
            K2 = b2.shape[0] if b_mn else b2.shape[1]
        d.K2 = K2
        d.A2, d.lda2 = a2.data_ptr(), a2.stride(0)
        d.B2, d.ldb2 = b2.data_ptr(), b2.stride(0)
    else:
        d.K2 = 0
    d.a2_group_n, d.a2_group_stride = a2_group_n, a2_group_stride
    d.splits, d.epilogue, d.alpha = splits, epilogue, alpha
    if epilogue != EPI_LSE:
        d.C, d.ldc = out.data_ptr(), out.stride(0)
    d.bias = _ptr(bias)
    if residual is not None:
        d.residual, d.ldr = residual.data_ptr(), residual.stride(0)
    d.act = act
    d.labels = _ptr(labels)
    d.lse_partials = _ptr(lse_partials)
    d.label_logit = _ptr(label_logit)
    d.lse = _ptr(lse)
    d.dlp = _ptr(dlp)
    d.dent, d.ent = _ptr(dent), _ptr(ent)
    d.rope_pos, d.rope_cs, d.rope_cols, d.rope_S = _ptr(rope_pos), _ptr(rope_cs), rope_cols, rope_S
    d.max_ctas = max_ctas
    d.cluster_mode = cluster + 1   # -1 auto, 0 off, 1 on
    _lib.check(lib.opadpo_op_gemm(ctypes.byref(d), ctypes.c_void_p(_stream())), "opadpo_op_gemm")
    return out


def decode_gemm(x: torch.Tensor, w: torch.Tensor, *, splits: int = 0, max_splits: int = 0, alpha: float = 1.0):
    """Weight-streaming decode GEMM: x [Mb <= 128, K] bf16, w [N, K] bf16 -> fp32 slices [splits, Mb, N] (sum over dim 0 =
    alpha * x w^T). splits = 0 lets the library pick."""
    lib = _lib.load()
    _chk2d(x), _chk2d(w)
    Mb, K = x.shape
    N = w.shape[0]
    cap = splits if splits > 0 else (max_splits if max_splits > 0 else 16)
    part = torch.empty((cap, Mb, N), device=x.device, dtype=torch.float32)
    used = ctypes.c_int32(0)
    _lib.check(lib.opadpo_op_decode_gemm(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), Mb, N, K, splits, max_splits, alpha,
                                         part.data_ptr(), ctypes.byref(used), ctypes.c_void_p(_stream())), "opadpo_op_decode_gemm")
    return part[:used.value]


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """[N, S] bool/uint8 key-validity mask -> packed u64 bit words [N, ceil(S/64)] (int64 storage)."""
    lib = _lib.load()
    m = mask.to(torch.uint8).contiguous()
    N, S = m.shape
    bits = torch.empty((N, (S + 63) // 64), device=m.device, dtype=torch.int64)
    _lib.check(lib.opadpo_op_pack_mask(m.data_ptr(), bits.data_ptr(), N, S, ctypes.c_void_p(_stream())), "pack_mask")
    return bits


def _attn_desc(q, k, v, o, lse, N, S, H, HD, causal, scale, kbits, prefix_len=0, segment_len=0):
    d = AttnDesc()
    d.prefix_len, d.segment_len = prefix_len, segment_len
    d.N, d.S, d.H, d.HD, d.causal, d.scale = N, S, H, HD, int(causal), scale
    assert q.stride(0) == k.stride(0) == v.stride(0)
    d.q, d.k, d.v, d.ld_qkv = q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(0)
    d.o, d.ld_o = o.data_ptr(), o.stride(0)
    d.lse = _ptr(lse)
    d.kbits = _ptr(kbits)
    return d


def attn_fwd(q, k, v, o, lse, *, N, S, H, HD, causal, scale, kbits=None, prefix_len=0, segment_len=0):
    """q,k,v,o: [N*S, H*HD] bf16 views; lse: [N*S, H] f32 (or None)."""
    lib = _lib.load()
    d = _attn_desc(q, k, v, o, lse, N, S, H, HD, causal, scale, kbits, prefix_len, segment_len)
    _lib.check(lib.opadpo_op_attn_fwd(ctypes.byref(d), ctypes.c_void_p(_stream())), "opadpo_op_attn_fwd")
    return o


def attn_bwd(q, k, v, o, lse, dO, dq, dk, dv, delta, *, N, S, H, HD, causal, scale, kbits=None, prefix_len=0,
             segment_len=0):
    lib = _lib.load()
    d = _attn_desc(q, k, v, o, lse, N, S, H, HD, causal, scale, kbits, prefix_len, segment_len)
    d.dO, d.ld_do = dO.data_ptr(), dO.stride(0)
    d.delta = delta.data_ptr()
    assert dq.stride(0) == dk.stride(0) == dv.stride(0)
    d.dq, d.dk, d.dv, d.ld_dqkv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dq.stride(0)
    _lib.check(lib.opadpo_op_attn_bwd(ctypes.byref(d), ctypes.c_void_p(_stream())), "opadpo_op_attn_bwd")
