"""Host-side mirror of DPOTrainer's loss / step logic (opadpo/dpo_models/dpo_trainer.py, rl_trainer.py) on the engine.

* `FusedDPO.pair(...)`           <- DPOTrainer.dpo_loss + `.mean()` + anchor term + reward stats (one kernel per pair)
* `FusedDPO.compute_policy_loss` <- DPOTrainer.compute_policy_loss (:475-802): Pair1..Pair4, AncPO, stats dict
* `DPOStepper`                   <- DPOTrainer.rollout (:214-427, reference-adapter pass) + RLTrainer.step_with_rollouts
                                    (:138-179: backward, grad accumulation, clip, AdamW) for one micro-batch list
"""
from __future__ import annotations

import os

import ctypes
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from ._lib import DpoPairDesc
from .dist import OverlappedGradAllReduce, allreduce_grad_bucket, make_grad_allreduce

_DIV = {"reverse_kl": 0, "js_divergence": 1, "alpha_divergence": 2}


@dataclass
class DPOArgs:
    """The loss-relevant subset of TrainingArguments (opadpo/opadpo_train.py:125-458; effective defaults of
    run/train_opa_dpo.sh:57-77)."""
    beta: float = 0.1
    label_smoothing: float = 0.0
    reference_free: bool = False
    f_divergence_type: str = "reverse_kl"
    alpha_divergence_coef: float = 1.0
    loss_type: str = "sigmoid"
    standard_pair_coef: float = 1.0
    AI_pair_coef: float = 1.0
    CoPO: bool = False
    CoPO_coef: float = 0.2
    AncPO: bool = False
    mDPO_anchor: bool = True
    Anchor_value: float = 0.0
    Anchor_coef: float = 1.0
    detailed_report: bool = False
    response_score: bool = False
    response_image_relation: bool = False
    temperature: float = 1.0
    response_len: int = 896
    max_grad_norm: float = 1.0
    learning_rate: float = 1e-6
    weight_decay: float = 0.0
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    # loop shape (opadpo_train.py:410-437; effective defaults of run/train_opa_dpo.sh:39-44)
    CoPO_method: str = "random"            # random | blockwise | attention
    CoPO_mask_ratio: float = 0.3
    rollout_accumulation_steps: int = 8
    step_per_device_batch_size: int = 2
    gradient_accumulation_steps: int = 4
    noptepochs: int = 1
    bf16: bool = True
    # outer loop (opadpo_train.py:125-458; run/train_opa_dpo.sh:39-53)
    total_epochs: int = 4
    rollout_batch_size: int = 64
    save_steps: int = 75
    save_steps_extra_list: tuple = ()
    eval_steps: Optional[int] = None
    max_step: int = 500
    warmup_steps: int = 5
    output_dir: Optional[str] = None


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


class _PairFn(torch.autograd.Function):
    """Forward launches the fused pair kernel, which also produces d(total)/d(pi_c), d(total)/d(pi_r) for
    total = loss_coef*mean(loss) + anchor_coef*mean(anchor); backward only scales them."""

    @staticmethod
    def forward(ctx, pi_c, pi_r, ref_c, ref_r, s_c, s_r, p: dict):
        dev = pi_c.device
        f = lambda t: None if t is None else t.detach().to(dev, torch.float32).contiguous()
        pi_c_, pi_r_, ref_c_, ref_r_, s_c_, s_r_ = map(f, (pi_c, pi_r, ref_c, ref_r, s_c, s_r))
        d_c = torch.zeros_like(pi_c_)
        d_r = torch.zeros_like(pi_r_)
        out = torch.empty(8, device=dev, dtype=torch.float32)
        d = DpoPairDesc()
        d.pi_c, d.pi_r, d.ref_c, d.ref_r = pi_c_.data_ptr(), pi_r_.data_ptr(), ref_c_.data_ptr(), ref_r_.data_ptr()
        d.s_c, d.s_r = _ptr(s_c_), _ptr(s_r_)
        d.d_c, d.d_r, d.out, d.n = d_c.data_ptr(), d_r.data_ptr(), out.data_ptr(), pi_c_.numel()
        d.beta, d.label_smoothing = p["beta"], p["label_smoothing"]
        d.reference_free, d.divergence, d.alpha_div = int(p["reference_free"]), p["divergence"], p["alpha_div"]
        d.loss_coef = p["loss_coef"]
        d.anchor_mode, d.anchor_sign = p["anchor_mode"], p["anchor_sign"]
        d.anchor_value, d.anchor_coef = p["anchor_value"], p["anchor_coef"]
        d.grad_scale = 1.0
        lib = _lib.load()
        _lib.check(lib.opadpo_op_dpo_pair(ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "opadpo_op_dpo_pair")
        ctx.save_for_backward(d_c, d_r)
        ctx.shape_c, ctx.shape_r = pi_c.shape, pi_r.shape
        ctx.mark_non_differentiable(out)
        total = out[0] * p["loss_coef"] + out[1] * p["anchor_coef"]
        return total, out

    @staticmethod
    def backward(ctx, g_total, _g_out):
        d_c, d_r = ctx.saved_tensors
        return (g_total * d_c).view(ctx.shape_c), (g_total * d_r).view(ctx.shape_r), None, None, None, None, None


class FusedDPO:
    def __init__(self, args: DPOArgs, pair_impl=None):
        """`pair_impl(pi_c, pi_r, ref_c, ref_r, s_c, s_r, params) -> (total, out[8])`: the one-pair operation; the default —
        and the only implementation this package has — is the fused CUDA kernel (`_PairFn`). The CPU tests pass a torch
        restatement of the kernel's contract here to check the composition below against the reference's
        compute_policy_loss run verbatim."""
        self._pair_impl = pair_impl if pair_impl is not None else _PairFn.apply
        if args.loss_type != "sigmoid":
            raise ValueError(f"Unknown loss type: {args.loss_type}.")   # dpo_trainer.py:468-471
        if args.f_divergence_type not in _DIV:
            raise ValueError(f"unknown f_divergence_type {args.f_divergence_type}")
        self.args = args

    def pair(self, pi_c, pi_r, ref_c, ref_r, s_c=None, s_r=None, loss_coef=1.0, anchor_sign=0.0) -> Tuple[Tensor, Tensor]:
        """-> (loss_coef*mean(dpo_loss) + Anchor_coef*mean(anchor term of this pair's chosen reward), out[8])."""
        a = self.args
        mode = 0
        if a.AncPO and anchor_sign != 0.0:
            mode = 1 if a.mDPO_anchor else 2
        p = dict(beta=a.beta, label_smoothing=a.label_smoothing, reference_free=a.reference_free,
                 divergence=_DIV[a.f_divergence_type], alpha_div=getattr(a, "alpha_divergence_coef", 1.0), loss_coef=float(loss_coef),
                 anchor_mode=mode, anchor_sign=float(anchor_sign), anchor_value=a.Anchor_value,
                 anchor_coef=a.Anchor_coef if mode else 0.0)
        return self._pair_impl(pi_c, pi_r, ref_c, ref_r, s_c, s_r, p)

    def compute_policy_loss(self, outputs: Dict[str, Tensor], rollouts: Dict[str, Tensor],
                            outputs_mask: Optional[Dict[str, Tensor]] = None) -> Tuple[Tensor, Dict[str, Tensor]]:
        """dpo_trainer.py:582-712 and the flat stats dict of :714-802 with every key the reference returns: `loss/*` (5),
        `policy/*` (12 reward means / gaps) and `logprobs/*` (15 masked mean / min / max of the raw policy log-probs).
        `outputs` / `outputs_mask` are AutoregressivePolicy.forward results."""
        a = self.args
        dev = outputs["standard_response_logprobs"].device
        g = lambda k: rollouts[k].to(dev, torch.float32)
        p_std, p_org, p_ai = (outputs[k + "_logprobs"] for k in
                              ("standard_response", "original_generate_response", "AI_pseudo_response"))
        r_std, r_org, r_ai = (g("ref_base_" + k + "_logprobs") for k in
                              ("standard_response", "original_generate_response", "AI_pseudo_response"))
        use_sc = a.detailed_report and a.response_score
        use_ir = a.detailed_report and a.response_image_relation
        sc_ai = g("AI_pseudo_response_scores") if use_sc else None
        sc_org = g("original_generate_response_scores") if use_sc else None
        ir_ai = g("AI_pseudo_response_image_relations") if use_ir else None
        t1, o1 = self.pair(p_std, p_org, r_std, r_org, loss_coef=a.standard_pair_coef, anchor_sign=+1.0)
        t2, o2 = self.pair(p_ai, p_org, r_ai, r_org, sc_ai, sc_org, loss_coef=a.AI_pair_coef, anchor_sign=-1.0)
        loss = t1 + t2
        zero = torch.zeros((), device=dev)
        stats = {"loss/stand_gen": o1[0], "loss/AI_gen": o2[0], "loss/stand_mask": zero, "loss/AI_mask": zero}
        anc = (o1[1] + o2[1]) if a.AncPO else zero
        stats.update({"policy/stand_gen_chosen_mean": o1[2], "policy/stand_gen_reject_mean": o1[3],
                      "policy/stand_gen_gap_mean": o1[2] - o1[3], "policy/AI_gen_chosen_mean": o2[2],
                      "policy/AI_gen_reject_mean": o2[3], "policy/AI_gen_gap_mean": o2[2] - o2[3]})
        if a.CoPO:
            pm_std = outputs_mask["mask_standard_response_logprobs"]
            pm_ai = outputs_mask["mask_AI_pseudo_response_logprobs"]
            rm_std, rm_ai = g("ref_mask_standard_response_logprobs"), g("ref_mask_AI_pseudo_response_logprobs")
            t3, o3 = self.pair(p_std, pm_std, r_std, rm_std, loss_coef=a.standard_pair_coef * a.CoPO_coef, anchor_sign=+1.0)
            t4, o4 = self.pair(p_ai, pm_ai, r_ai, rm_ai, ir_ai, ir_ai, loss_coef=a.AI_pair_coef * a.CoPO_coef,
                               anchor_sign=-1.0)
            loss = loss + t3 + t4
            stats["loss/stand_mask"], stats["loss/AI_mask"] = o3[0], o4[0]
            if a.AncPO:
                anc = anc + o3[1] + o4[1]
            # reward masks 3/4 are the CHOSEN masks for both sides (dpo_trainer.py:633-636): recompute reject means
            m3, m4 = (r_std != 0).float(), (r_ai != 0).float()
            rr3 = a.beta * (pm_std.detach() - (0.0 if a.reference_free else 1.0) * rm_std)
            rr4 = a.beta * (pm_ai.detach() - (0.0 if a.reference_free else 1.0) * rm_ai)
            rj3, rj4 = (rr3 * m3).sum() / m3.sum(), (rr4 * m4).sum() / m4.sum()
            stats.update({"policy/stand_mask_chosen_mean": o3[2], "policy/stand_mask_reject_mean": rj3,
                          "policy/stand_mask_gap_mean": o3[2] - rj3, "policy/AI_mask_chosen_mean": o4[2],
                          "policy/AI_mask_reject_mean": rj4, "policy/AI_mask_gap_mean": o4[2] - rj4})
            lpm_std, lpm_ai = pm_std.detach(), pm_ai.detach()
        else:
            # the reference reports the masked pairs' rewards (scalar zeros, dpo_trainer.py:694-700) under the same keys
            stats.update({f"policy/{p}_mask_{q}_mean": zero for p in ("stand", "AI") for q in ("chosen", "reject", "gap")})
            lpm_std = lpm_ai = torch.zeros_like(p_std, dtype=torch.float32)
            if a.AncPO:
                # without CoPO the reference adds the anchor on chosen_rewards3/4 == 0 (scalars): -logsigmoid(-a) - logsigmoid(a)
                # resp. 2*a^2 (dpo_trainer.py:700-708); constants w.r.t. the parameters
                av = torch.tensor(a.Anchor_value, device=dev)
                const = (-torch.nn.functional.logsigmoid(-av) - torch.nn.functional.logsigmoid(av)) if a.mDPO_anchor else 2 * av * av
                loss = loss + const * a.Anchor_coef
                anc = anc + const
        stats["loss/AncPO"] = anc
        stats.update(_logprob_stats(p_org.detach(), p_std.detach(), p_ai.detach(), lpm_std, lpm_ai))
        return loss, {k: v.detach() for k, v in stats.items()}


def _logprob_stats(lp_org: Tensor, lp_std: Tensor, lp_ai: Tensor, lpm_std: Tensor, lpm_ai: Tensor) -> Dict[str, Tensor]:
    """The `logprobs/*` block of compute_policy_loss's stats (dpo_trainer.py:714-751): masked mean over the batch, and the
    batch mean of every row's masked minimum / maximum, of the five raw policy log-prob arrays. A token counts when the
    policy's own log-prob there is non-zero (pads are exactly 0); the two masked-image arrays use the unmasked arrays'
    masks, as the reference does. All five arrays go through one batched reduction each instead of 15 separate ones."""
    L = torch.stack([lp_org, lp_std, lp_ai, lpm_std, lpm_ai]).float()          # [5, b, R]
    m_org, m_std, m_ai = lp_org != 0, lp_std != 0, lp_ai != 0
    M = torch.stack([m_org, m_std, m_ai, m_std, m_ai])
    mean = (L * M).sum((1, 2)) / M.sum((1, 2))
    lo = torch.where(M, L, L.new_tensor(1e9)).amin(2).mean(1)
    hi = torch.where(M, L, L.new_tensor(-1e9)).amax(2).mean(1)
    names = ("original", "standard", "AI", "standard_mask", "AI_mask")
    out = {}
    for i, n in enumerate(names):
        out[f"logprobs/{n}_logprobs"] = mean[i]
        out[f"logprobs/{n}_logprobs_min"] = lo[i]
        out[f"logprobs/{n}_logprobs_max"] = hi[i]
    return out


class DPOStepper:
    """One optimisation step of the OPA-DPO loop on the engine, in the north-star form (K responses per query,
    policy + frozen-reference adapters):  reference pass (DPOTrainer.rollout :333-335) -> policy pass -> fused loss ->
    LoRA backward -> [all-reduce] -> clip + AdamW (RLTrainer.step_with_rollouts :138-179)."""

    def __init__(self, policy, ref_policy, args: DPOArgs, world_size: int = 1, process_group=None):
        self.policy, self.ref_policy, self.args = policy, ref_policy, args
        self.loss = FusedDPO(args)
        self.model = policy.base_model
        self.world_size = world_size
        self.process_group = process_group
        self.lr = args.learning_rate
        # reference pass on a side stream next to the policy forward (step(); -1 % step time at 7B); OPADPO_CONCURRENT_REF=0: one stream
        self.concurrent_reference = os.environ.get("OPADPO_CONCURRENT_REF", "1") != "0"
        self._side = self._ref_ready = None
        self.overlap = None
        if world_size > 1 and self.model.flat_param is not None:
            self.overlap = make_grad_allreduce(self.model.engine, self.model.flat_param.grad, process_group)
            self.model.flat_param.grad = self.overlap.grad   # the symmetric bucket the engine now accumulates into (or the same tensor)

    @torch.no_grad()
    def reference_logprobs(self, batch: Dict[str, Tensor]) -> Dict[str, Tensor]:
        self.ref_policy.eval()
        out = self.ref_policy(**batch, temperature=self.args.temperature)
        return {"ref_base_" + k: v for k, v in out.items()}

    def pair_loss(self, batch: Dict[str, Tensor], ref: Dict[str, Tensor], chosen: str, rejected: str, grad_accum: int = 1,
                  last_micro_batch: bool = True, ref_ready: Optional[torch.cuda.Event] = None):
        self.policy.train()
        if self.overlap is not None and last_micro_batch:
            self.overlap.arm()   # this backward completes the bucket: all-reduce its slices as they become final
        out = self.policy(**batch, temperature=self.args.temperature, mode="policy")
        if ref_ready is not None:    # the reference pass ran on a side stream next to the policy forward
            torch.cuda.current_stream().wait_event(ref_ready)
        total, o = self.loss.pair(out[chosen + "_logprobs"], out[rejected + "_logprobs"],
                                  ref["ref_base_" + chosen + "_logprobs"], ref["ref_base_" + rejected + "_logprobs"])
        (total / grad_accum).backward()
        return total.detach(), o

    def optimizer_step(self) -> Tensor:
        a = self.args
        eng = self.model.engine
        # the LoRA-grad bucket is all-reduced (sum) once per optimizer step — as per-layer slices overlapped with the backward
        # when armed, else as one bucket-wide call; 1/world is applied inside the fused clip+AdamW kernel
        if self.overlap is not None:
            scale = self.overlap.finish()
        else:
            scale = allreduce_grad_bucket(self.model.flat_param.grad, self.process_group) if self.world_size > 1 else 1.0
        norm = eng.optimizer_step(self.lr, (a.adam_beta1, a.adam_beta2), a.adam_epsilon, a.weight_decay, a.max_grad_norm, scale)
        eng.zero_grad()
        return norm

    def step(self, batch: Dict[str, Tensor], chosen: str, rejected: str):
        if self.concurrent_reference and batch[chosen].is_cuda:
            # The frozen-adapter pass and the policy forward are independent until the loss: enqueue them on two streams. Their
            # persistent GEMM grids cannot share an SM (shared memory), so each launch's tail fills with the other's first
            # tiles, and one pass's HBM-bound kernels (norms, SwiGLU, embeddings) run under the other's tensor-bound ones.
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side, self._ref_ready = torch.cuda.Stream(), torch.cuda.Event()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                ref = self.reference_logprobs(batch)
                self._ref_ready.record(self._side)
            loss, o = self.pair_loss(batch, ref, chosen, rejected, ref_ready=self._ref_ready)
        else:
            ref = self.reference_logprobs(batch)
            loss, o = self.pair_loss(batch, ref, chosen, rejected)
        norm = self.optimizer_step()
        return loss, norm
