"""Kernel parity: tcgen05 GEMM (through the C ABI) vs a plain PyTorch fp32 matmul of the same bf16 operands.

Tolerance: bf16 output rounding (rel 2^-8) on top of fp32 accumulation; stated per test."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(torch.bfloat16)


def _close(out, ref, rtol=1e-2, atol=None, what=""):
    ref = ref.float()
    out = out.float()
    if atol is None:
        atol = 1e-2 * ref.abs().max().item() + 1e-6
    err = (out - ref).abs()
    bad = err > (atol + rtol * ref.abs())
    assert not bad.any(), f"{what}: {int(bad.sum())} / {bad.numel()} mismatches, max err {err.max().item():.4g}, " \
                          f"ref max {ref.abs().max().item():.4g}, first bad idx {bad.nonzero()[0].tolist()}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 128), (200, 264, 136),
                                   (1215, 768, 320), (4096, 4096, 4096)])
def test_gemm_kk(M, N, K):
    from opa_dpo_b200 import ops
    a, b = _mk((M, K), 1), _mk((N, K), 2)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().t(), what=f"kk {M}x{N}x{K}")


def test_gemm_kk_strided_views():
    from opa_dpo_b200 import ops
    M, N, K = 300, 512, 192
    abuf, bbuf = _mk((M, K + 64), 3), _mk((N, K + 8), 4)
    a, b = abuf[:, 64:], bbuf[:, :K]
    cbuf = torch.zeros((M, N + 256), device="cuda", dtype=torch.bfloat16)
    out = cbuf[:, 256:]
    ops.gemm(a, b, out)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().t(), what="strided")
    assert cbuf[:, :256].abs().max().item() == 0


def test_gemm_epilogue_bias_act_residual():
    from opa_dpo_b200 import ops
    M, N, K = 333, 512, 256
    a, b = _mk((M, K), 5), _mk((N, K), 6, 0.1)
    bias, res = _mk((N,), 7), _mk((M, N), 8)
    for act, fn in [(ops.ACT_NONE, lambda x: x), (ops.ACT_GELU, torch.nn.functional.gelu),
                    (ops.ACT_QUICK_GELU, lambda x: x * torch.sigmoid(1.702 * x))]:
        out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, b, out, bias=bias, residual=res, act=act, alpha=0.5)
        torch.cuda.synchronize()
        ref = fn(0.5 * (a.float() @ b.float().t()) + bias.float()) + res.float()
        _close(out, ref, what=f"epilogue act={act}")


def test_gemm_lora_second_segment_and_groups():
    """[x | t] [W^T ; B^T] with per-group t columns (fused q|k|v): y_g = x W_g^T + t_g B_g^T."""
    from opa_dpo_b200 import ops
    M, K, r, G, Ng = 500, 320, 64, 3, 256
    x, W = _mk((M, K), 9), _mk((G * Ng, K), 10, 0.05)
    t, B = _mk((M, G * r), 11), _mk((G * Ng, r), 12, 0.05)
    out = torch.empty((M, G * Ng), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, W, out, a2=t, b2=B, a2_group_n=Ng, a2_group_stride=r, K2=r)
    torch.cuda.synchronize()
    ref = x.float() @ W.float().t()
    for g in range(G):
        ref[:, g * Ng:(g + 1) * Ng] += t[:, g * r:(g + 1) * r].float() @ B[g * Ng:(g + 1) * Ng].float().t()
    _close(out, ref, what="lora groups")
    # ungrouped, tiny rank (K2 = 8 -> zero-filled k tail)
    t8, B8 = _mk((M, 8), 13), _mk((G * Ng, 8), 14)
    ops.gemm(x, W, out, a2=t8, b2=B8)
    torch.cuda.synchronize()
    _close(out, x.float() @ W.float().t() + t8.float() @ B8.float().t(), what="lora r=8")


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 520, 200), (2048, 4096, 1024)])
def test_gemm_b_mn_major(M, N, K):
    """dgrad form: dx = dy W with W stored [K(contraction), N(output)] row-major."""
    from opa_dpo_b200 import ops
    a, b = _mk((M, K), 15), _mk((K, N), 16)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, b_mn=True)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float(), what=f"k-mn {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K,splits", [(128, 256, 64, 1), (256, 520, 1000, 1), (256, 4096, 5000, 7), (40, 264, 700, 3)])
def test_gemm_wgrad_mn_mn_atomic(M, N, K, splits):
    """LoRA wgrad form: dW[M,N] += alpha * A^T B with A [K,M], B [K,N] (contraction over tokens), fp32 atomics."""
    from opa_dpo_b200 import ops
    a, b = _mk((K, M), 17), _mk((K, N), 18)
    init = torch.randn((M, N), device="cuda", dtype=torch.float32)
    out = init.clone()
    ops.gemm(a, b, out, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_ATOMIC, splits=splits, alpha=2.0)
    torch.cuda.synchronize()
    ref = init + 2.0 * (a.float().t() @ b.float())
    _close(out, ref, rtol=1e-4, atol=1e-3 * ref.abs().max().item(), what="wgrad")


def test_gemm_lse_and_cegrad():
    """Fused lm_head epilogues vs torch log_softmax / autograd."""
    from opa_dpo_b200 import ops
    M, V, K, T = 700, 1000, 256, 0.7
    h, W = _mk((M, K), 19), _mk((V, K), 20, 0.2)
    labels = torch.randint(0, V, (M,), device="cuda", dtype=torch.int32)
    labels[::7] = -1
    nt = ops.gemm_n_tiles(V)
    part = torch.zeros((M, nt, 4), device="cuda", dtype=torch.float32)
    lab = torch.zeros((M,), device="cuda", dtype=torch.float32)
    ops.gemm(h, W, None, epilogue=ops.EPI_LSE, alpha=1.0 / T, labels=labels, lse_partials=part, label_logit=lab)
    torch.cuda.synchronize()
    z = (h.float() @ W.float().t()) / T
    # combine partials (base-2 domain)
    mx = part[..., 0].max(dim=1).values
    s = (part[..., 1] * torch.exp2(part[..., 0] - mx[:, None])).sum(1)
    zs = (part[..., 2] * torch.exp2(part[..., 0] - mx[:, None])).sum(1)
    lse = (mx + torch.log2(s)) * math.log(2.0)
    ref_lse = torch.logsumexp(z, dim=-1)
    assert (lse - ref_lse).abs().max().item() < 2e-4 * max(1.0, ref_lse.abs().max().item())
    valid = labels >= 0
    ref_lab = z.gather(1, labels.clamp(min=0).long()[:, None])[:, 0]
    assert (lab - ref_lab)[valid].abs().max().item() < 1e-4 * max(1.0, ref_lab.abs().max().item())
    ent = lse - (zs / s) * math.log(2.0)
    lsm = torch.log_softmax(z, -1)
    ref_ent = -(lsm.exp() * lsm).sum(-1)
    assert (ent - ref_ent).abs().max().item() < 1e-3
    # CE backward tile
    dlp = torch.randn((M,), device="cuda", dtype=torch.float32) * valid
    G = torch.empty((M, V), device="cuda", dtype=torch.bfloat16)
    ops.gemm(h, W, G, epilogue=ops.EPI_CEGRAD, alpha=1.0 / T, labels=labels, lse=lse.contiguous(), dlp=dlp)
    torch.cuda.synchronize()
    onehot = torch.zeros_like(z)
    onehot[valid, labels[valid].long()] = 1.0
    refG = dlp[:, None] * (onehot - lsm.exp()) / T
    _close(G, refG, what="cegrad")
    # entropy head: d(sum_m dent[m] * H_m)/d(logits) next to the log-prob gradient (opa_trainer.py:63,82-85), vs autograd
    dent = torch.randn((M,), device="cuda", dtype=torch.float32) * valid
    dent[::5] = 0.0
    hw = (h.float() @ W.float().t()).requires_grad_(True)
    lsm2 = torch.log_softmax(hw / T, -1)
    H = -(lsm2.exp() * lsm2).sum(-1)
    lp_lab = lsm2.gather(1, labels.clamp(min=0).long()[:, None])[:, 0]
    ((dent * H).sum() + (dlp * lp_lab * valid).sum()).backward()
    G2 = torch.full((M, V), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(h, W, G2, epilogue=ops.EPI_CEGRAD, alpha=1.0 / T, labels=labels, lse=lse.contiguous(), dlp=dlp, dent=dent,
             ent=ent.contiguous())
    torch.cuda.synchronize()
    _close(G2, hw.grad, what="cegrad+entropy")
    G3 = torch.full((M, V), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(h, W, G3, epilogue=ops.EPI_CEGRAD, alpha=1.0 / T, labels=labels, lse=lse.contiguous(), dent=dent,
             ent=ent.contiguous())   # entropy gradient alone (dlp = NULL)
    torch.cuda.synchronize()
    hw.grad = None
    lsm3 = torch.log_softmax(hw / T, -1)
    (dent * -(lsm3.exp() * lsm3).sum(-1)).sum().backward()
    _close(G3, hw.grad, what="entropy only")


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (1100, 768, 320), (1215 * 3, 512, 256), (2048, 264, 136)])
def test_gemm_cluster_multicast_kk(M, N, K, mode):
    """2-CTA variants — mode 1: cluster with the B tile multicast; mode 2: CTA pair (tcgen05 cta_group::2, one 256x256
    MMA over two SMs): same results as the single-CTA kernel, incl. an odd number of m blocks."""
    from opa_dpo_b200 import ops
    a, b = _mk((M, K), 41), _mk((N, K), 42)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out, cluster=mode)
    torch.cuda.synchronize()
    _close(out, a.float() @ b.float().t(), what=f"cluster kk {M}x{N}x{K}")
    out1 = torch.empty_like(out)
    ops.gemm(a, b, out1, cluster=0)
    torch.cuda.synchronize()
    assert torch.equal(out, out1), "cluster and single-CTA kernels must agree bit for bit (same k order)"


@pytest.mark.parametrize("mode", [1, 2])
def test_gemm_cluster_lora_groups_bias_residual_and_mn_major(mode):
    from opa_dpo_b200 import ops
    M, K, r, G, Ng = 1500, 320, 64, 3, 256
    x, W = _mk((M, K), 43), _mk((G * Ng, K), 44, 0.05)
    t, B = _mk((M, G * r), 45), _mk((G * Ng, r), 46, 0.05)
    bias, res = _mk((G * Ng,), 47), _mk((M, G * Ng), 48)
    out = torch.empty((M, G * Ng), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, W, out, a2=t, b2=B, a2_group_n=Ng, a2_group_stride=r, K2=r, bias=bias, residual=res, cluster=mode)
    torch.cuda.synchronize()
    ref = x.float() @ W.float().t() + bias.float() + res.float()
    for g in range(G):
        ref[:, g * Ng:(g + 1) * Ng] += t[:, g * r:(g + 1) * r].float() @ B[g * Ng:(g + 1) * Ng].float().t()
    _close(out, ref, what="cluster lora groups")
    # dgrad form: B MN-major, two segments
    dy, Wt = _mk((M, 512), 49), _mk((512, 384), 50, 0.05)
    u, At = _mk((M, 64), 51), _mk((64, 384), 52, 0.05)
    dx = torch.empty((M, 384), device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, Wt, dx, a2=u, b2=At, b_mn=True, K2=64, cluster=mode)
    torch.cuda.synchronize()
    _close(dx, dy.float() @ Wt.float() + u.float() @ At.float(), what="cluster dgrad")
    # wgrad form: both MN-major, fp32 atomic accumulate, no split
    T = 1300
    a, b = _mk((T, 1104), 53)[:, :1100], _mk((T, 304), 54)[:, :300]
    acc2 = torch.zeros((1100, 304), device="cuda", dtype=torch.float32)[:, :300]
    ops.gemm(a, b, acc2, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_ATOMIC, cluster=1)
    torch.cuda.synchronize()
    _close(acc2, a.float().t() @ b.float(), what="cluster wgrad")


@pytest.mark.parametrize("cluster", [0, 2])
@pytest.mark.parametrize("with_pos", [True, False])
def test_gemm_rope_epilogue(cluster, with_pos):
    """EPI_BF16_ROPE: q|k|v projection (with the LoRA second segment) whose epilogue applies HF LLaMA's rotary embedding to
    the q and k columns, vs an fp32 matmul + apply_rotary_pos_emb (rotate_half pairing j <-> j + 64 inside 128-wide heads)."""
    from opa_dpo_b200 import ops
    M, d, K, r, S = 1100, 256, 320, 64, 275
    x, W = _mk((M, K), 31), _mk((3 * d, K), 32, 0.1)
    t, B = _mk((M, 3 * r), 33), _mk((3 * d, r), 34, 0.1)
    g = torch.Generator(device="cuda").manual_seed(35)
    pos = torch.randint(0, 500, (M,), generator=g, device="cuda", dtype=torch.int32) if with_pos else None
    inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, device="cuda", dtype=torch.float32) / 128))
    ang = torch.arange(512, device="cuda", dtype=torch.float32)[:, None] * inv[None, :]
    cs = torch.stack([ang.cos(), ang.sin()], -1).contiguous()              # [positions, 64, 2] = float2 table
    out = torch.full((M, 3 * d), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, W, out, a2=t, b2=B, a2_group_n=d, a2_group_stride=r, epilogue=ops.EPI_BF16_ROPE, rope_pos=pos, rope_cs=cs,
             rope_cols=2 * d, rope_S=S, cluster=cluster)
    torch.cuda.synchronize()
    ref = x.float() @ W.float().t()
    for gidx in range(3):
        ref[:, gidx * d:(gidx + 1) * d] += t.float()[:, gidx * r:(gidx + 1) * r] @ B.float()[gidx * d:(gidx + 1) * d].t()
    p = pos.long() if with_pos else torch.arange(M, device="cuda") % S
    c, s = cs[p, :, 0], cs[p, :, 1]                                        # [M, 64]
    qk = ref[:, :2 * d].reshape(M, 2 * d // 128, 128)
    lo, hi = qk[..., :64], qk[..., 64:]
    rot = torch.cat([lo * c[:, None, :] - hi * s[:, None, :], hi * c[:, None, :] + lo * s[:, None, :]], -1).reshape(M, 2 * d)
    _close(out[:, :2 * d], rot, what="rope q|k")
    _close(out[:, 2 * d:], ref[:, 2 * d:], what="v untouched")


@pytest.mark.parametrize("cluster", [0, 2])
def test_gemm_swiglu_epilogue_bit_identical_to_unfused(cluster):
    """EPI_SWIGLU_BF16 over a weight whose gate / up rows are interleaved in blocks of 128 == plain projection followed by
    SwiGLU with the unfused path's roundings (gate, up -> bf16; silu -> bf16; product -> bf16), bit for bit."""
    from opa_dpo_b200 import ops
    M, F, K = 1100, 384, 320
    x, W = _mk((M, K), 41), _mk((2 * F, K), 42, 0.3)
    Wil = torch.empty_like(W)
    for t in range(F // 128):
        Wil[t * 256:t * 256 + 128] = W[t * 128:(t + 1) * 128]
        Wil[t * 256 + 128:(t + 1) * 256] = W[F + t * 128:F + (t + 1) * 128]
    gu = torch.empty((M, 2 * F), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, W, gu, cluster=cluster)
    act = torch.full((M, F), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, Wil.contiguous(), act, epilogue=ops.EPI_SWIGLU_BF16, N=2 * F, cluster=cluster)
    torch.cuda.synchronize()
    g, u = gu[:, :F].float(), gu[:, F:].float()
    want = ((g / (1.0 + torch.exp(-g))).to(torch.bfloat16).float() * u).to(torch.bfloat16)
    diff = (act.float() - want.float()).abs()
    # __expf vs torch.exp differ in the last bits of the exponential: allow one bf16 ulp on a few elements, none larger
    assert (diff > 0).float().mean().item() < 2e-3 and diff.max().item() <= 2.0 ** -7 * want.float().abs().max().item()


@pytest.mark.parametrize("cluster", [0, 2])
def test_gemm_swiglu_backward_epilogue_equals_unfused(cluster):
    """EPI_SWIGLU_BWD: the down projection's dgrad (dy W, B operand MN-major, LoRA as a second K segment) with the SwiGLU
    backward in its epilogue == the same GEMM writing d(act) as bf16 followed by the SwiGLU-backward formula on it (d gate
    = d act * up * sig * (1 + gate (1 - sig)), d up = d act * silu(gate)); ragged M, F a multiple of 32 only."""
    from opa_dpo_b200 import ops
    M, F, K, r = 1100, 416, 320, 64
    dy, W = _mk((M, K), 51, 0.5), _mk((K, F), 52, 0.2)          # W = down-projection weight [out = K, in = F]
    u, A = _mk((M, r), 53, 0.5), _mk((r, F), 54, 0.2)           # LoRA: dx += u A
    gu = _mk((M, 2 * F), 55, 1.5)
    dact = torch.empty((M, F), device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, W, dact, a2=u, b2=A, b_mn=True, cluster=cluster)
    dgu = torch.full((M, 2 * F), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, W, dgu, a2=u, b2=A, b_mn=True, epilogue=ops.EPI_SWIGLU_BWD, residual=gu, cluster=cluster)
    torch.cuda.synchronize()
    g, up, da = gu[:, :F].float(), gu[:, F:].float(), dact.float()
    sg = 1.0 / (1.0 + torch.exp(-g))
    want = torch.cat([da * up * sg * (1.0 + g * (1.0 - sg)), da * g * sg], 1).to(torch.bfloat16)
    assert torch.isfinite(dgu.float()).all()
    diff = (dgu.float() - want.float()).abs()
    # __expf vs torch.exp: a last-bit difference in the exponential may move a result by one bf16 ulp, never more
    assert (diff > 0).float().mean().item() < 5e-3 and diff.max().item() <= 2.0 ** -7 * want.float().abs().max().item()


@pytest.mark.parametrize("Mb,N,K,splits", [(32, 4096, 4096, 0), (32, 12288, 4096, 0), (1, 256, 64, 1), (5, 136, 200, 2),
                                           (32, 22016, 4096, 0), (32, 4096, 11008, 0), (64, 520, 1000, 3), (100, 1000, 320, 0),
                                           (128, 32000, 4096, 2), (17, 128, 4096, 16)])
def test_decode_gemm_weight_streaming(Mb, N, K, splits):
    """The rollout decode step's weight-streaming GEMM (weight rows on the MMA's M, activation rows on its N): the sum of
    its fp32 split-K slices equals x W^T (fp32 matmul of the same bf16 operands, fp32 tolerance: 1e-3 of the row scale);
    ragged N / K / Mb edges, every activation-tile width, run-to-run bit-identical."""
    from opa_dpo_b200 import ops
    x, w = _mk((Mb, K), 3, 0.5), _mk((N, K), 4, 0.05)
    part = ops.decode_gemm(x, w, splits=splits, alpha=0.5)
    part2 = ops.decode_gemm(x, w, splits=splits, alpha=0.5)
    torch.cuda.synchronize()
    assert part.shape[1:] == (Mb, N) and part.shape[0] >= 1 and (splits == 0 or part.shape[0] == splits)
    assert torch.equal(part, part2)
    ref = 0.5 * (x.float() @ w.float().t())
    got = part.sum(0)
    tol = 1e-3 * ref.abs().max().item() + 1e-6
    assert (got - ref).abs().max().item() <= tol, (part.shape[0], (got - ref).abs().max().item(), tol)
    # strided operands (views into wider buffers), as the engine passes them
    xw = _mk((Mb, K + 64), 5, 0.5)
    part3 = ops.decode_gemm(xw[:, 64:], w, splits=splits)
    torch.cuda.synchronize()
    ref3 = xw[:, 64:].float() @ w.float().t()
    assert (part3.sum(0) - ref3).abs().max().item() <= 1e-3 * ref3.abs().max().item() + 1e-6
