"""Optimizer / scheduler / checkpoint mirrors for the engine-owned LoRA bucket.

* `B200AdamW`       <- create_optimizer (utils/trainer_utils.py:9-38): AdamW(beta 0.9/0.999, eps 1e-8, wd) — the update itself
                      is the fused clip+AdamW kernel over the flat fp32 bucket (`opadpo_optimizer_step`); this class only
                      carries hyper-parameters, the step count and (de)serialises the moments like a torch optimizer.
* `cosine_schedule` <- create_scheduler (utils/trainer_utils.py:41-49) = HF get_scheduler("cosine", warmup, total).
* `save_checkpoint` / `load_checkpoint` <- DPOTrainer.save_model / resume_training (dpo_trainer.py:837-931):
      <output_dir>/checkpoint-{step}/adapter_model/lora_policy/{adapter_model.bin, adapter_config.json},
      optimizer.pt, scheduler.pt; `get_last_checkpoint` as utils/lora_utils.py:31-47.
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, Optional

import torch


class B200AdamW:
    def __init__(self, model, lr: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: float = 1.0):
        if model.flat_param is None:
            raise RuntimeError("B200AdamW: call model.make_trainable() first")
        self.model = model
        self.param_groups = [dict(lr=lr, initial_lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                  params=[model.flat_param])]
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self.last_grad_norm: Optional[torch.Tensor] = None

    def step(self, grad_scale: float = 1.0):
        g = self.param_groups[0]
        # slices the reference keeps frozen (the projector's LoRA in the DPO stage, B200LlavaModel.freeze_projector_lora): no
        # gradient -> outside the norm, moments stay zero, no Adam update; decoupled weight decay would still shrink them, so
        # with a non-zero decay their weights are put back afterwards
        frozen = [(o, c) for o, c in getattr(self.model, "frozen_grad_ranges", ())]
        keep = []
        for off, cnt in frozen:
            self.model.flat_param.grad[off:off + cnt].zero_()
            if g["weight_decay"]:
                keep.append((off, cnt, self.model.flat_param.data[off:off + cnt].clone()))
        # max_grad_norm None = no clipping (rl_trainer.py:164: `if self.args.max_grad_norm is not None`); the kernel skips the clip at 0
        self.last_grad_norm = self.model.engine.optimizer_step(g["lr"], g["betas"], g["eps"], g["weight_decay"],
                                                               0.0 if self.max_grad_norm is None else self.max_grad_norm, grad_scale)
        if keep:
            for off, cnt, w in keep:
                self.model.flat_param.data[off:off + cnt].copy_(w)
            self.model.engine.sync_shadow()
        self.step_count += 1
        return self.last_grad_norm

    def zero_grad(self, set_to_none: bool = True):
        self.model.engine.zero_grad()
        if hasattr(self.model, "release_slots"):
            self.model.release_slots()      # a forward left undifferentiated by a failed step must not pin its slot

    def state_dict(self) -> Dict:
        """torch-optimizer shaped: ONE flat parameter (the engine's LoRA bucket). Next to the moments it carries the fp32
        master copy of the trainable tensors: adapter_model.bin holds them rounded to bf16 (PEFT's format, what the reference
        stores), and at lr 1e-6 most of an update is below a bf16 ulp — without the master a resumed run would restart from
        the rounded values and diverge from an uninterrupted one."""
        v = self.model.engine.buffers()
        n = v["n_train"]
        return dict(state={0: dict(step=self.step_count, exp_avg=v["exp_avg"].detach().cpu().clone(),
                                   exp_avg_sq=v["exp_avg_sq"].detach().cpu().clone(),
                                   master=v["master"][:n].detach().cpu().clone())},
                    param_groups=[{k: val for k, val in self.param_groups[0].items() if k != "params"} | {"params": [0]}],
                    max_grad_norm=self.max_grad_norm)

    def load_state_dict(self, sd: Dict):
        """Call after the adapter has been loaded (load_checkpoint's order): the fp32 master stored here replaces the
        bf16-rounded values adapter_model.bin gave, then the bf16 shadow the GEMMs read is refreshed."""
        state = sd.get("state", {})
        v = self.model.engine.buffers()
        n = v["n_train"]
        if len(state) != 1 or 0 not in state or state[0]["exp_avg"].numel() != n:
            raise ValueError("optimizer.pt does not hold this engine's flat LoRA bucket (one parameter of "
                             f"{n} elements; got {len(state)} parameter entries). A file written by the reference trainer "
                             "keeps one state entry per LoRA tensor and cannot be resumed here; only adapter_model/ is "
                             "format-compatible (INTEGRATION.md).")
        st = state[0]
        v["exp_avg"].copy_(st["exp_avg"].to(v["exp_avg"].device))
        v["exp_avg_sq"].copy_(st["exp_avg_sq"].to(v["exp_avg_sq"].device))
        if "master" in st:
            v["master"][:n].copy_(st["master"].to(v["master"].device))
            self.model.engine.sync_shadow()
        self.step_count = int(st["step"])
        self.model.engine.set_option("adam_step", self.step_count)
        for k, val in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = tuple(val) if k == "betas" else val
        self.max_grad_norm = sd.get("max_grad_norm", self.max_grad_norm)


class cosine_schedule:
    """HF get_scheduler(name, optimizer, num_warmup_steps, num_training_steps) [UPSTREAM transformers.optimization] for the
    schedules the reference's `--lr_scheduler_type` can name without extra arguments — "cosine" (the recipe's,
    run/train_opa_dpo.sh:47-53; the default here, hence the class name), "linear", "constant", "constant_with_warmup" —
    stepped once per outer step (rl_trainer.py:188-189)."""

    KINDS = ("cosine", "linear", "constant", "constant_with_warmup")

    def __init__(self, optimizer: B200AdamW, num_warmup_steps: int, num_training_steps: int, kind: str = "cosine"):
        kind = getattr(kind, "value", kind)        # transformers.SchedulerType members compare by their string value
        if kind not in self.KINDS:
            raise ValueError(f"lr_scheduler_type {kind!r} is not one of {self.KINDS}")
        self.opt, self.warmup, self.total, self.kind = optimizer, num_warmup_steps, num_training_steps, kind
        self.last_epoch = 0
        self._apply()

    def factor(self, step: int) -> float:
        if self.kind == "constant":
            return 1.0
        if step < self.warmup:
            return float(step) / float(max(1, self.warmup))
        if self.kind == "constant_with_warmup":
            return 1.0
        if self.kind == "linear":
            return max(0.0, float(self.total - step) / float(max(1, self.total - self.warmup)))
        progress = float(step - self.warmup) / float(max(1, self.total - self.warmup))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 0.5 * 2.0 * progress)))

    def _apply(self):
        for g in self.opt.param_groups:
            g["lr"] = g["initial_lr"] * self.factor(self.last_epoch)

    def step(self):
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return [g["lr"] for g in self.opt.param_groups]

    def state_dict(self):
        return dict(last_epoch=self.last_epoch, warmup=self.warmup, total=self.total, kind=self.kind)

    def load_state_dict(self, sd):
        """Accepts this class's own state and a torch LambdaLR state dict (what the reference's scheduler.pt holds,
        dpo_trainer.py:884-886: `last_epoch` is all that matters there — the lambda itself is not serialised, so warm-up
        and total keep the constructor's values, exactly as on the reference's resume)."""
        self.last_epoch = int(sd["last_epoch"])
        self.warmup = sd.get("warmup", self.warmup)
        self.total = sd.get("total", self.total)
        self.kind = sd.get("kind", self.kind)
        self._apply()


def warmup_steps_of(args, num_training_steps: int) -> int:
    """TrainingArguments.get_warmup_steps [UPSTREAM transformers]: `warmup_steps` when positive, else
    ceil(num_training_steps * warmup_ratio); the args object's own method is used when it has one."""
    if callable(getattr(args, "get_warmup_steps", None)):
        return int(args.get_warmup_steps(num_training_steps))
    ws = int(getattr(args, "warmup_steps", 0) or 0)
    return ws if ws > 0 else math.ceil(num_training_steps * float(getattr(args, "warmup_ratio", 0.0) or 0.0))


def get_last_checkpoint(folder: str) -> Optional[str]:
    """utils/lora_utils.py:31-47: highest checkpoint-N, None when a `completed` marker exists or nothing is there."""
    if not os.path.isdir(folder):
        return None
    if os.path.exists(os.path.join(folder, "completed")):
        return None
    best, best_step = None, -1
    for name in os.listdir(folder):
        m = re.fullmatch(r"checkpoint-(\d+)", name)
        if m and os.path.isdir(os.path.join(folder, name)) and int(m.group(1)) > best_step:
            best, best_step = os.path.join(folder, name), int(m.group(1))
    return best


def last_checkpoint_and_completion(checkpoint_dir: str):
    """The reference's own return contract (utils/lora_utils.py:31-47): (path of the highest checkpoint-N or None, whether a
    `completed` marker exists). checkpoint-0 counts as "no checkpoint", as there. One difference, on purpose: a
    `checkpoint-final` directory next to the numbered ones is skipped here, where the reference's int() raises."""
    if not os.path.isdir(checkpoint_dir):
        return None, False
    if os.path.exists(os.path.join(checkpoint_dir, "completed")):
        return None, True
    best = get_last_checkpoint(checkpoint_dir)
    if best is None or best.endswith("checkpoint-0"):
        return None, False
    return best, False


def save_checkpoint_dir(model, optimizer: B200AdamW, scheduler: Optional[cosine_schedule], ckpt: str,
                        is_main_process: bool = True, for_best_ckpt: bool = False):
    """DPOTrainer.save_model (dpo_trainer.py:837-899) into the directory it is given: rank 0 writes the policy adapter in
    PEFT layout under <ckpt>/adapter_model/lora_policy, removes stray pytorch_model*.bin files there, and — unless
    `for_best_ckpt` — writes optimizer.pt and scheduler.pt and deletes the optimizer.pt of every sibling directory."""
    if not is_main_process:
        return None
    ckpt = ckpt.rstrip("/")
    os.makedirs(ckpt, exist_ok=True)
    model.save_adapter(os.path.join(ckpt, "adapter_model"), "lora_policy")
    for name in os.listdir(ckpt):
        if name.startswith("pytorch_model") and name.endswith(".bin"):
            os.remove(os.path.join(ckpt, name))
    if for_best_ckpt:
        return ckpt
    torch.save(optimizer.state_dict(), os.path.join(ckpt, "optimizer.pt"))
    if scheduler is not None:
        torch.save(scheduler.state_dict(), os.path.join(ckpt, "scheduler.pt"))
    parent = os.path.dirname(ckpt) or "."
    for name in os.listdir(parent):
        other = os.path.join(parent, name, "optimizer.pt")
        if os.path.join(parent, name) != ckpt and os.path.isfile(other):
            os.remove(other)
    return ckpt


def save_checkpoint(model, optimizer: B200AdamW, scheduler: Optional[cosine_schedule], output_dir: str, step,
                    is_main_process: bool = True):
    """save_checkpoint_dir into <output_dir>/checkpoint-{step} (the directory RLTrainer.train names, rl_trainer.py:250-257)."""
    return save_checkpoint_dir(model, optimizer, scheduler, os.path.join(output_dir, f"checkpoint-{step}"), is_main_process)


def load_checkpoint(model, optimizer: B200AdamW, scheduler: Optional[cosine_schedule], ckpt: str) -> int:
    """DPOTrainer.resume_training (dpo_trainer.py:901-931): adapter from <ckpt>/adapter_model/lora_policy, optimizer and
    scheduler state, step counter parsed from the directory name."""
    state = torch.load(os.path.join(ckpt, "adapter_model", "lora_policy", "adapter_model.bin"), map_location="cpu")
    model.load_adapter_state("lora_policy", state)
    opt_path = os.path.join(ckpt, "optimizer.pt")
    if os.path.exists(opt_path):
        optimizer.load_state_dict(torch.load(opt_path, map_location="cpu", weights_only=False))
    sch_path = os.path.join(ckpt, "scheduler.pt")
    if scheduler is not None and os.path.exists(sch_path):
        scheduler.load_state_dict(torch.load(sch_path, map_location="cpu", weights_only=False))
    m = re.search(r"checkpoint-(\d+)", os.path.basename(ckpt.rstrip("/")))
    return int(m.group(1)) if m else 0
