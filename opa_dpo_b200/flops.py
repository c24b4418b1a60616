"""Algorithmic FLOPs of the OPA-DPO step: multiply-add = 2 FLOPs, causal attention at half of full,
lm_head on the R response rows only, CLIP executed layers only, LoRA included; no recompute, no padding savings, no
prefix sharing, no image dedupe (the reference repeats the image per response: rl_models.py:107-111)."""
from __future__ import annotations

from .engine import EngineConfig


def forward_flops_per_sequence(cfg: EngineConfig, Q: int, R: int) -> dict:
    d, F, L, V, r = cfg.d_model, cfg.d_ffn, cfg.n_layers, cfg.vocab, cfg.lora_r
    P = cfg.n_patches
    S = Q - 1 + P + R
    lin = L * (4 * d * d + 3 * d * F)
    lora = L * r * (4 * (d + d) + 2 * (d + F) + (F + d))
    vd, vF, vL, Sv = cfg.v_d, cfg.v_ffn, cfg.v_layers, P + 1
    clip_lin = vL * (4 * vd * vd + 2 * vd * vF)
    clip_lora = vL * r * (4 * (vd + vd) + 2 * (vd + vF))
    proj_lin = vd * d + d * d
    proj_lora = r * ((vd + d) + (d + d))
    return dict(
        base=2.0 * S * lin,
        lora=2.0 * S * lora,
        attn=L * 2.0 * S * S * d,
        lm_head=2.0 * R * d * V,
        clip=2.0 * Sv * (clip_lin + clip_lora) + vL * 4.0 * Sv * Sv * vd + 2.0 * P * 3 * cfg.v_patch ** 2 * vd,
        proj=2.0 * P * (proj_lin + proj_lora),
    )


def flops_per_pair(cfg: EngineConfig, Q: int, R: int) -> dict:
    """One preference pair = policy fwd x2 + reference fwd x2 + policy bwd x2 (LoRA-only training: dgrad through the
    frozen base = 1x base-linear FLOPs, LoRA dgrad+wgrad = 2x, attention 2.5x, lm_head dgrad 1x, projector 2x, CLIP 0)."""
    f = forward_flops_per_sequence(cfg, Q, R)
    fwd = sum(f.values())
    bwd = f["base"] + 2 * f["lora"] + 2.5 * f["attn"] + f["lm_head"] + 2 * f["proj"]
    return dict(fwd_per_seq=fwd, bwd_per_seq=bwd, per_pair=4 * fwd + 2 * bwd, parts=f)


def executed_attention_flops_per_sample(cfg: EngineConfig, Q: int, R: int, K: int = 2, packed: bool = True) -> dict:
    """Attention FLOPs the engine's LLaMA kernels execute for ONE sample (K responses), counted at element level under the
    causal + segment mask (tile padding not counted): per visible (query, key) pair and head, 2 GEMM units forward
    (QK^T, PV) and 7 backward in the two-kernel design (dQ kernel: S, dP, dQ; dK/dV kernel: S, dP, dV, dK), each
    2*head_dim FLOPs. packed = shared-prefix rows [prefix | resp_0 .. resp_{K-1}]; otherwise K stacked rows."""
    S0 = Q - 1 + cfg.n_patches
    if packed:
        pairs = S0 * (S0 + 1) / 2 + K * (R * S0 + R * (R + 1) / 2)
    else:
        S = S0 + R
        pairs = K * S * (S + 1) / 2
    unit = 2.0 * pairs * cfg.d_model * cfg.n_layers      # one GEMM unit over every head: 2 * pairs * head_dim * n_heads
    return dict(fwd=2 * unit, bwd=7 * unit, visible_pairs=pairs)
