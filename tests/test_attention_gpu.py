"""Kernel parity: tcgen05 flash attention fwd/bwd (through the C ABI) vs plain PyTorch fp32 attention + autograd.

Tolerances: outputs are bf16 (rel 2^-8) and P / dS are rounded to bf16 before the second MMA (as flash-attn does)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attn(q, k, v, mask, causal, scale):
    """q,k,v: [N,S,H,D] fp32; mask: [N,S] bool key validity. Returns o [N,S,H,D] and lse [N,S,H]."""
    N, S, H, D = q.shape
    s = torch.einsum("nihd,njhd->nhij", q, k) * scale
    allow = mask[:, None, None, :].expand(N, H, S, S).clone()
    if causal:
        allow &= torch.ones(S, S, dtype=torch.bool, device=q.device).tril()[None, None]
    s = s.masked_fill(~allow, float("-inf"))
    lse = torch.logsumexp(s, dim=-1)                       # [N,H,S]
    p = torch.exp(s - lse[..., None])
    p = torch.nan_to_num(p, nan=0.0)                        # fully masked rows -> 0
    o = torch.einsum("nhij,njhd->nihd", p, v)
    return o, lse.permute(0, 2, 1)


def _run(N, S, H, D, causal, lens=None, left=None, seed=0, bwd=True):
    from opa_dpo_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    T = N * S
    qkv = (torch.randn((T, 3 * H * D), generator=g, device="cuda") * 1.0).to(torch.bfloat16)
    q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
    mask = torch.ones((N, S), dtype=torch.bool, device="cuda")
    if lens is not None:
        for n in range(N):
            lo = left[n] if left is not None else 0
            mask[n, :lo] = False
            mask[n, lo + lens[n]:] = False
    kbits = ops.pack_mask(mask)
    o = torch.zeros((T, H * D), device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros((T, H), device="cuda", dtype=torch.float32)
    scale = 1.0 / math.sqrt(D)
    ops.attn_fwd(q, k, v, o, lse, N=N, S=S, H=H, HD=D, causal=causal, scale=scale, kbits=kbits)
    torch.cuda.synchronize()

    qf = q.float().reshape(N, S, H, D).requires_grad_(True)
    kf = k.float().reshape(N, S, H, D).requires_grad_(True)
    vf = v.float().reshape(N, S, H, D).requires_grad_(True)
    ro, rlse = _ref_attn(qf, kf, vf, mask, causal, scale)
    # rows whose query is itself a pad position are don't-care in the reference path (zeroed downstream)
    if causal:
        qvalid = mask
    else:
        qvalid = torch.ones_like(mask)
    sel = qvalid[:, :, None, None].expand_as(ro)
    err = ((o.float().reshape(N, S, H, D) - ro) * sel).abs().max().item()
    assert err < 2e-2, f"attn fwd max err {err}"
    lerr = ((lse.reshape(N, S, H) - rlse) * qvalid[:, :, None]).nan_to_num(posinf=0, neginf=0).abs().max().item()
    assert lerr < 2e-3, f"lse max err {lerr}"
    if not bwd:
        return
    dO = (torch.randn((T, H * D), generator=g, device="cuda") * qvalid.reshape(T, 1)).to(torch.bfloat16)
    dqkv = torch.full((T, 3 * H * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    dq, dk, dv = dqkv[:, :H * D], dqkv[:, H * D:2 * H * D], dqkv[:, 2 * H * D:]
    delta = torch.empty((T, H), device="cuda", dtype=torch.float32)
    ops.attn_bwd(q, k, v, o, lse, dO, dq, dk, dv, delta, N=N, S=S, H=H, HD=D, causal=causal, scale=scale, kbits=kbits)
    torch.cuda.synchronize()
    (ro * sel).backward(dO.float().reshape(N, S, H, D))
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        got = got.float().reshape(N, S, H, D)
        assert torch.isfinite(got).all(), f"{name} has non-finite values"
        tol = 2e-2 * max(1.0, ref.abs().max().item())
        e = (got - ref).abs().max().item()
        assert e < tol, f"{name} max err {e} (tol {tol})"


@pytest.mark.parametrize("S", [64, 128, 200, 577, 1215])
def test_attn_causal_hd128(S):
    _run(2, S, 2, 128, True, seed=S)


def test_attn_causal_padding():
    # left-padded query + right-padded response: one contiguous valid span per row (rl_models.py:99-102)
    _run(3, 333, 2, 128, True, lens=[333, 200, 150], left=[0, 17, 60], seed=7)


def test_attn_noncausal_hd64_clip():
    _run(2, 577, 4, 64, False, seed=11, bwd=False)


def test_attn_noncausal_hd128():
    _run(1, 300, 2, 128, False, seed=13)


def test_attn_masked_holes():
    """CoPO 'attention' mode punches random holes into the image-token keys (non-contiguous mask)."""
    from opa_dpo_b200 import ops
    N, S, H, D = 2, 260, 2, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn((N * S, 3 * H * D), generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
    mask = torch.rand((N, S), generator=g, device="cuda") > 0.3
    mask[:, 0] = True
    o = torch.zeros((N * S, H * D), device="cuda", dtype=torch.bfloat16)
    ops.attn_fwd(q, k, v, o, None, N=N, S=S, H=H, HD=D, causal=True, scale=1 / math.sqrt(D), kbits=ops.pack_mask(mask))
    torch.cuda.synchronize()
    ro, _ = _ref_attn(q.float().reshape(N, S, H, D), k.float().reshape(N, S, H, D), v.float().reshape(N, S, H, D), mask,
                      True, 1 / math.sqrt(D))
    assert (o.float().reshape(N, S, H, D) - ro).abs().max().item() < 2e-2


@pytest.mark.parametrize("S0,R,K,pads", [(703, 512, 2, False), (100, 60, 3, True), (64, 128, 2, True), (37, 200, 2, False)])
def test_attn_shared_prefix_segments(S0, R, K, pads):
    """Packed rows [prefix | resp_0 | ... | resp_{K-1}]: response queries see the prefix and their own segment only
    (tile skipping + in-tile exclusion), forward and backward, against a dense fp32 reference with the same tree mask."""
    from opa_dpo_b200 import ops
    N, H, D = 2, 2, 128
    S = S0 + K * R
    T = N * S
    g = torch.Generator(device="cuda").manual_seed(S0 + R)
    qkv = torch.randn((T, 3 * H * D), generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:]
    mask = torch.ones((N, S), dtype=torch.bool, device="cuda")
    if pads:
        mask[1, :5] = False                                  # left pads in the prefix
        for kk in range(K):                                  # ragged responses: right pads inside each segment
            mask[1, S0 + kk * R + R - 7 - kk:S0 + (kk + 1) * R] = False
    scale = 1.0 / math.sqrt(D)
    o = torch.zeros((T, H * D), device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros((T, H), device="cuda", dtype=torch.float32)
    kbits = ops.pack_mask(mask)
    ops.attn_fwd(q, k, v, o, lse, N=N, S=S, H=H, HD=D, causal=True, scale=scale, kbits=kbits, prefix_len=S0, segment_len=R)
    torch.cuda.synchronize()
    # dense reference
    i = torch.arange(S, device="cuda")[:, None]
    j = torch.arange(S, device="cuda")[None, :]
    seg_start = torch.where(i >= S0, S0 + ((i - S0) // R) * R, torch.zeros_like(i))
    tree = (j <= i) & ~((i >= S0) & (j >= S0) & (j < seg_start))
    qf = q.float().reshape(N, S, H, D).requires_grad_(True)
    kf = k.float().reshape(N, S, H, D).requires_grad_(True)
    vf = v.float().reshape(N, S, H, D).requires_grad_(True)
    s = torch.einsum("nihd,njhd->nhij", qf, kf) * scale
    allow = tree[None, None] & mask[:, None, None, :]
    s = s.masked_fill(~allow, float("-inf"))
    p = torch.nan_to_num(torch.softmax(s, dim=-1), nan=0.0)
    ro = torch.einsum("nhij,njhd->nihd", p, vf)
    sel = mask[:, :, None, None].expand_as(ro)
    err = ((o.float().reshape(N, S, H, D) - ro) * sel).abs().max().item()
    assert err < 2e-2, f"segment attn fwd max err {err}"
    dO = (torch.randn((T, H * D), generator=g, device="cuda") * mask.reshape(T, 1)).to(torch.bfloat16)
    dqkv = torch.full((T, 3 * H * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    dq, dk, dv = dqkv[:, :H * D], dqkv[:, H * D:2 * H * D], dqkv[:, 2 * H * D:]
    delta = torch.empty((T, H), device="cuda", dtype=torch.float32)
    ops.attn_bwd(q, k, v, o, lse, dO, dq, dk, dv, delta, N=N, S=S, H=H, HD=D, causal=True, scale=scale, kbits=kbits,
                 prefix_len=S0, segment_len=R)
    torch.cuda.synchronize()
    (ro * sel).backward(dO.float().reshape(N, S, H, D))
    for name, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        got = got.float().reshape(N, S, H, D)
        assert torch.isfinite(got).all(), f"{name} has non-finite values"
        tol = 2e-2 * max(1.0, ref.abs().max().item())
        e = (got - ref).abs().max().item()
        assert e < tol, f"{name} max err {e} (tol {tol})"
