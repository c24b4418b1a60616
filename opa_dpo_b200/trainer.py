"""Host-side mirror of the reference's outer training loop on the engine.

`B200DPOTrainer` keeps the shape of `DPOTrainer` / `RLTrainer` (opadpo/dpo_models/dpo_trainer.py:214-427, 475-802;
rl_trainer.py:138-195) — `rollout(queries_batches)`, `compute_policy_loss(rollouts_batch)`,
`step_with_rollouts(rollouts)`, `step(train_dataloader, step_idx)` — with what the row asks for:

* the reference-adapter pass runs on the device inside the step and its outputs STAY in HBM: no `.cpu()` parking of the
  rollouts (dpo_trainer.py:421-424), no per-batch `gc.collect()` / `empty_cache()` (:227-228, rl_trainer.py:153-154), no
  re-upload in compute_policy_loss (:489-560);
* the discarded policy forward of rollout() (:336-338) is dropped — its outputs are never read;
* with `merge_frozen` the frozen reference adapter runs on merged weights (no LoRA GEMMs at all);
* `noptepochs > 1` reuses the same rollouts, as the reference does.
Sharing one GEMM launch between the reference and the policy pass was evaluated and left out: their activations differ
from the first LoRA layer on, so no FLOPs are shared, and at M >= 13k rows the GEMMs are tensor-bound — streaming the
base weights once instead of twice saves < 1 %.

The shuffled mini-batch order comes from torch's own DataLoader over row indices, so a seeded run visits the rollouts in
the reference's order (rl_trainer.py:325-345). Gradient accumulation follows accelerate 0.21's counter (a boundary every
`gradient_accumulation_steps` micro-batches, the loss divided by it) [UPSTREAM accelerate]; the LR scheduler steps once
per outer step (rl_trainer.py:188-189).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Iterable, List, Optional, Tuple

import torch
from torch import Tensor
from torch.utils.data import DataLoader

from . import collate
from .dist import OverlappedGradAllReduce, allreduce_grad_bucket, make_grad_allreduce
from .dpo import DPOArgs, FusedDPO

FIRST_STEP_IDX = 1      # rl_trainer.py:46


def train_loop(trainer, dataset_len: int, next_batch_source: Callable[[], Iterable], resume_training_ckpt: Optional[str] = None):
    """RLTrainer.train (rl_trainer.py:217-279) as a free function over a trainer exposing `args` (total_epochs,
    rollout_batch_size, rollout_accumulation_steps, save_steps, save_steps_extra_list, eval_steps, output_dir), `max_step`,
    `create_optimizer_and_scheduler(n)`, `resume_training(ckpt) -> step`, `step(loader, step_idx)`, `save_model(dir)`,
    `evaluate(step_idx)`. Same step numbering (first step is 1), same save cadence (at the START of a step whose index is a
    multiple of save_steps or listed, never at the resumed step itself), same skipping of already-trained steps (their
    batches are consumed so the data order is unchanged), same `max_step` cut."""
    a = trainer.args
    total_steps = dataset_len * a.total_epochs // a.rollout_batch_size
    trainer.create_optimizer_and_scheduler(min(total_steps, trainer.max_step))
    skipping_steps = trainer.resume_training(resume_training_ckpt) if resume_training_ckpt is not None else 0
    loader = next_batch_source()
    log_history = []
    for step_idx in range(FIRST_STEP_IDX, total_steps + FIRST_STEP_IDX):
        if step_idx < skipping_steps:
            for _ in range(a.rollout_accumulation_steps):
                next(loader)
            continue
        if step_idx >= trainer.max_step:
            break
        if step_idx % a.save_steps == 0 or step_idx in a.save_steps_extra_list:
            if step_idx > skipping_steps:
                trainer.save_model(os.path.join(a.output_dir, f"checkpoint-{step_idx}"))
        if a.eval_steps is not None and step_idx % a.eval_steps == 0:
            trainer.evaluate(step_idx)
        log_history.append(trainer.step(loader, step_idx))
    return log_history


RESPONSES = ("standard_response", "original_generate_response", "AI_pseudo_response")
_OPTIONAL = ("original_generate_response_scores", "AI_pseudo_response_scores", "original_generate_response_image_relations",
             "AI_pseudo_response_image_relations")


class B200DPOTrainer:
    def __init__(self, args: DPOArgs, policy, ref_policy, optimizer, lr_scheduler=None, world_size: int = 1,
                 process_group=None, mask_ops=None, loss=None, device=None):
        """mask_ops (`mask_images`, `mask_percentage_per_row`), loss (a FusedDPO) and device default to the CUDA
        implementations — the only ones this package has; the CPU tests pass restatements to compare `rollout()` with the
        reference's DPOTrainer.rollout run verbatim."""
        self.args, self.policy, self.ref_policy = args, policy, ref_policy
        self.optimizer, self.lr_scheduler = optimizer, lr_scheduler
        self.loss = loss if loss is not None else FusedDPO(args)
        self.mask_ops = mask_ops if mask_ops is not None else collate
        self.model = policy.base_model
        self.device = device if device is not None else self.model.engine.device
        self.world_size, self.process_group = world_size, process_group
        self._micro_step = 0
        self.log_fn: Optional[Callable] = None        # log_fn(stats, step=) <- accelerator.log (dpo_trainer.py:832-833)
        self.last_train_stats: Dict[str, Tensor] = {}
        self.max_step = getattr(args, "max_step", B200DPOTrainer.max_step)
        self.overlap = None
        if world_size > 1 and self.model.flat_param is not None:
            self.overlap = make_grad_allreduce(self.model.engine, self.model.flat_param.grad, process_group)
            self.model.flat_param.grad = self.overlap.grad   # the symmetric bucket the engine now accumulates into (or the same tensor)

    # ------------------------------------------------------------------ rollout (dpo_trainer.py:214-427)
    @torch.no_grad()
    def rollout(self, queries_data: Iterable[Dict[str, Tensor]]) -> Dict[str, Tensor]:
        a = self.args
        self.policy.eval()
        self.ref_policy.eval()
        out: List[Dict[str, Tensor]] = []
        dev = self.device
        for batch in queries_data:
            images = batch["images"].to(dev)
            if a.bf16:
                images = images.to(torch.bfloat16)          # :283-286 (the CoPO mean is taken over bf16 pixels)
            rb = {"images": images, "queries": batch["queries"].to(dev),
                  "queries_attn_masks": batch["queries_attention_mask"].to(dev)}
            for n in RESPONSES:
                rb[n] = batch[n].to(dev)
                rb[n + "_attention_mask"] = batch[n + "_attention_mask"].to(dev)
            fwd = {k: rb[k] for k in ("images", "queries", "queries_attn_masks") + RESPONSES}
            new_fwd = None
            if a.CoPO:
                sub = {"standard_response": rb["standard_response"], "AI_pseudo_response": rb["AI_pseudo_response"]}
                if a.CoPO_method in ("random", "blockwise"):
                    rb["masked_images"] = self.mask_ops.mask_images(images, a.CoPO_mask_ratio, a.CoPO_method)
                    new_fwd = dict(images=rb["masked_images"], queries=rb["queries"], queries_attn_masks=rb["queries_attn_masks"], **sub)
                elif a.CoPO_method == "attention":
                    assert images.size(-1) % 14 == 0
                    holes = self.mask_ops.mask_percentage_per_row(images.size(0), (images.size(-1) // 14) ** 2, a.CoPO_mask_ratio, dev)
                    rb["masked_query_attn_masks"] = torch.cat([holes.to(rb["queries_attn_masks"].dtype), rb["queries_attn_masks"]], 1)
                    new_fwd = dict(images=images, queries=rb["queries"], queries_attn_masks=rb["masked_query_attn_masks"], **sub)
                else:
                    raise NotImplementedError
            ref = self.ref_policy(**fwd, temperature=a.temperature)
            rb.update({"ref_base_" + k: v for k, v in ref.items()})
            if new_fwd is not None:
                ref_m = self.ref_policy(**new_fwd, temperature=a.temperature)
                rb.update({"ref_mask_" + k: v for k, v in ref_m.items()})
            if a.detailed_report and (a.response_score or a.response_image_relation):
                for k in _OPTIONAL:
                    rb[k] = batch[k].to(dev)
            out.append(rb)
        return {k: torch.cat([b[k] for b in out]) for k in out[0]}

    # ------------------------------------------------------------------ loss (dpo_trainer.py:475-802)
    def compute_policy_loss(self, rb: Dict[str, Tensor]) -> Tuple[Tensor, Dict[str, Tensor]]:
        a = self.args
        self.policy.train()
        resp = {n: rb[n] for n in RESPONSES}
        outputs = self.policy(images=rb["images"], queries=rb["queries"], queries_attn_masks=rb["queries_attn_masks"],
                              temperature=a.temperature, mode="policy", **resp)
        outputs_new = None
        if a.CoPO:
            new = {"mask_standard_response": rb["standard_response"], "mask_AI_pseudo_response": rb["AI_pseudo_response"]}
            if a.CoPO_method in ("random", "blockwise"):
                outputs_new = self.policy(images=rb["masked_images"], queries=rb["queries"],
                                          queries_attn_masks=rb["queries_attn_masks"], temperature=a.temperature,
                                          mode="policy", **new)
            else:
                outputs_new = self.policy(images=rb["images"], queries=rb["queries"],
                                          queries_attn_masks=rb["masked_query_attn_masks"], temperature=a.temperature,
                                          mode="policy", **new)
        return self.loss.compute_policy_loss(outputs, rb, outputs_new)

    # ------------------------------------------------------------------ optimisation (rl_trainer.py:138-179)
    def step_with_rollouts(self, rollouts: Dict[str, Tensor]) -> Dict[str, Tensor]:
        a = self.args
        n = next(iter(rollouts.values())).size(0)
        gas = max(1, a.gradient_accumulation_steps)
        stats_list: List[Dict[str, Tensor]] = []
        for _epoch in range(a.noptepochs):
            # the same sampler class the reference's rollouts DataLoader uses -> the same shuffled order under a seed
            for idx in DataLoader(range(n), batch_size=a.step_per_device_batch_size, shuffle=True, drop_last=True):
                idx = idx.to(self.device)
                rb = {k: v.index_select(0, idx) for k, v in rollouts.items()}
                self._micro_step += 1
                boundary = self._micro_step % gas == 0
                if self.overlap is not None and boundary and not a.CoPO:
                    self.overlap.arm()       # one backward completes the bucket: all-reduce its slices as they become final
                loss, stats = self.compute_policy_loss(rb)
                (loss / gas).backward()
                if boundary:
                    stats["loss/grad_norm"] = self._optimizer_step()
                    stats_list.append(stats)
        if not stats_list:
            return {}
        return {k: torch.stack([s[k] for s in stats_list]) for k in stats_list[0]}

    def _optimizer_step(self) -> Tensor:
        if self.overlap is not None:
            scale = self.overlap.finish()      # slices already in flight, or one bucket-wide all-reduce when not armed
        elif self.world_size > 1:
            scale = allreduce_grad_bucket(self.model.flat_param.grad, self.process_group)
        else:
            scale = 1.0
        norm = self.optimizer.step(scale).reshape(())      # clip_grad_norm_ + AdamW, fused (rl_trainer.py:164-175)
        self.optimizer.zero_grad()
        # the reference logs `loss/grad_norm` = _compute_grad_norm() AFTER clip_grad_norm_ (rl_trainer.py:164-171), i.e. the
        # clipped gradients' norm; the fused step returns the norm before clipping (kept as optimizer.last_grad_norm)
        mg = getattr(self.optimizer, "max_grad_norm", None)
        if mg is not None and mg > 0:
            norm = norm * torch.clamp(mg / (norm + 1e-6), max=1.0)
        return norm

    # ------------------------------------------------------------------ outer loop (rl_trainer.py:197-279)
    max_step: int = 300                       # TrainingArguments.max_step default (opadpo_train.py:152)

    def create_optimizer_and_scheduler(self, num_training_steps: int):
        """create_optimizer / create_scheduler (utils/trainer_utils.py:9-49) unless they were handed in: AdamW over the LoRA
        bucket with the args' learning_rate / adam_beta1,2 / adam_epsilon / weight_decay (every LoRA tensor is a "decay
        parameter" there: no biases, no norms) and the global-norm clip of rl_trainer.py:164-168 fused into its step; HF's
        `lr_scheduler_type` schedule (cosine in the recipe) with `args.get_warmup_steps(n)` warm-up steps."""
        a = self.args
        if self.optimizer is None:
            from .optim import B200AdamW
            self.optimizer = B200AdamW(self.model, lr=getattr(a, "learning_rate", 1e-6),
                                       betas=(getattr(a, "adam_beta1", 0.9), getattr(a, "adam_beta2", 0.999)),
                                       eps=getattr(a, "adam_epsilon", 1e-8), weight_decay=getattr(a, "weight_decay", 0.0),
                                       max_grad_norm=getattr(a, "max_grad_norm", 1.0))
        if self.lr_scheduler is None:
            from .optim import cosine_schedule, warmup_steps_of
            self.lr_scheduler = cosine_schedule(self.optimizer, warmup_steps_of(a, num_training_steps), num_training_steps,
                                                kind=getattr(a, "lr_scheduler_type", "cosine"))
        return self.optimizer, self.lr_scheduler

    def save_model(self, output_dir: Optional[str] = None, give_rw_access=True, check_corrupted=True, for_best_ckpt=False):
        """DPOTrainer.save_model (dpo_trainer.py:837-899): <output_dir>/adapter_model/lora_policy + optimizer.pt + scheduler.pt
        (rank 0 only; `output_dir` defaults to args.output_dir; any directory name, not only checkpoint-N)."""
        from .optim import save_checkpoint_dir
        output_dir = self.args.output_dir if output_dir is None else output_dir
        return save_checkpoint_dir(self.model, self.optimizer, self.lr_scheduler, output_dir,
                                   is_main_process=int(os.environ.get("RANK", 0)) == 0, for_best_ckpt=for_best_ckpt)

    def resume_training(self, checkpoint_dir: str) -> int:
        from .optim import load_checkpoint
        return load_checkpoint(self.model, self.optimizer, self.lr_scheduler, checkpoint_dir)

    def evaluate(self, step_idx: int):
        return None                            # the reference's DPOTrainer does not implement it either

    def train(self, dataset_len: int, train_dataloader_factory: Callable[[], Iterable], resume_training_ckpt: Optional[str] = None):
        return train_loop(self, dataset_len, train_dataloader_factory, resume_training_ckpt)

    def step(self, train_dataloader, step_idx: int = 0) -> Dict[str, float]:
        """RLTrainer.step (rl_trainer.py:181-195): rollout -> optimisation epochs -> scheduler -> the logged stats dict of
        record_step_stats. The per-optimizer-step tensors of step_with_rollouts stay readable as `last_train_stats`."""
        batches = [next(train_dataloader) for _ in range(self.args.rollout_accumulation_steps)]
        rollouts = self.rollout(batches)
        self.last_train_stats = self.step_with_rollouts(rollouts)
        if self.lr_scheduler is not None:
            self.lr_scheduler.step()
        return self.record_step_stats(train_stats=self.last_train_stats, rollouts=rollouts, step_idx=step_idx)

    def record_step_stats(self, train_stats: Dict[str, Tensor], rollouts: Dict[str, Tensor], step_idx: int, **_kwargs) -> Dict[str, float]:
        """DPOTrainer.record_step_stats (dpo_trainer.py:804-835): `objective/lr`, `objective/<key>` = batch mean of the
        per-row mean entropy over non-zero positions for every `*entropies*` rollout key, `dpo/<key>` (`logprobs/<key>`
        kept as is) = mean over the optimizer steps of the outer step; python floats; any '/' behind the first one turned
        into '-'. Rank 0 hands the dict to `log_fn(stats, step=step_idx)` (the accelerator.log stand-in) when one is set.
        All tensor values are read back with ONE device-to-host copy instead of one `.item()` per key."""
        stats: Dict[str, object] = {"objective/lr": self.optimizer.param_groups[0]["lr"]}
        for key in rollouts:
            if "entropies" in key:
                ent = rollouts[key]
                m = ent != 0.0
                stats[f"objective/{key}"] = ((ent * m).sum(1, keepdim=True) / m.sum(1, keepdim=True)).mean()
        for k, v in train_stats.items():
            stats[k if "logprobs/" in k else f"dpo/{k}"] = v.mean(dim=0)
        tkeys = [k for k, v in stats.items() if torch.is_tensor(v)]
        if tkeys:
            flat = torch.stack([stats[k].reshape(()).float() for k in tkeys]).tolist()
            stats.update(zip(tkeys, flat))
        renamed = {}
        for key, value in stats.items():
            cut = key.find("/") + 1
            renamed[key[:cut] + key[cut:].replace("/", "-") if "/" in key else key] = value
        if self.log_fn is not None and int(os.environ.get("RANK", 0)) == 0:
            self.log_fn(renamed, step=step_idx)
        return renamed


def train_opa_dpo(args, tokenizer, dataset_len: int, train_dataloader_factory: Callable[[], Iterable], accelerator=None,
                  model_factory=None, trainer_kwargs: Optional[dict] = None) -> "B200DPOTrainer":
    """The model / training half of the reference's entry point `opadpo_train.train` (opadpo/opadpo_train.py:583-595,
    727-797) — what is left of it once the hydra config, tokenizer and data module exist, which stay the reference's:

    1. `get_last_checkpoint(args.output_dir)`, falling back to `args.resume_dir` (:583-595);
    2. "building the LoRA model from scratch" when `base_model_name == policy_model_name_or_path` (:727-763): a freshly
       initialised adapter (`lora_r`, `lora_alpha`) is written once to `<output_dir>/init_model/{adapter_model.bin,
       adapter_config.json}` and becomes the policy checkpoint both adapters start from;
    3. `make_models(..., resume_from_checkpoint=checkpoint if args.resume_from_training)` (:769-778);
    4. `DPOTrainer(...).train(resume_training_ckpt=...)` (:781-793);
    5. rank 0 saves `<output_dir>/checkpoint-final`, then a barrier (:794-797).
    `dataset_len` / `train_dataloader_factory` stand for the data module (`len(train_dataset)` and the infinite loader of
    collated batches). Returns the trainer."""
    import json
    import torch.distributed as dist
    from . import checkpoint
    from .optim import last_checkpoint_and_completion
    from .policy import B200LlavaModel, make_models
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    ckpt, _completed = last_checkpoint_and_completion(args.output_dir)
    if ckpt is None and getattr(args, "resume_dir", None) is not None:
        ckpt, _ = last_checkpoint_and_completion(args.resume_dir)
    base_model = None
    if args.base_model_name == args.policy_model_name_or_path:
        init_dir = os.path.join(args.output_dir, "init_model")
        have = all(os.path.exists(os.path.join(init_dir, f)) for f in ("adapter_model.bin", "adapter_config.json"))
        if not have:
            cfg = checkpoint.engine_config_from_hf(args.base_model_name, lora_r=int(args.lora_r), lora_alpha=float(args.lora_alpha))
            factory = model_factory if model_factory is not None else B200LlavaModel
            base_model = factory(cfg, device=getattr(accelerator, "device", None), base_model_name_or_path=args.base_model_name)
            checkpoint.load_llava_base(base_model.engine, args.base_model_name)
            base_model.add_adapter("lora_policy")
            if rank == 0:
                os.makedirs(init_dir, exist_ok=True)
                torch.save(base_model.adapter_state_dict("lora_policy"), os.path.join(init_dir, "adapter_model.bin"))
                with open(os.path.join(init_dir, "adapter_config.json"), "w") as f:
                    json.dump(dict(base_model.peft_config["lora_policy"], lora_dropout=getattr(args, "lora_dropout", 0.0),
                                   bias=getattr(args, "lora_bias", "none")), f, indent=2, sort_keys=True)
            if dist.is_available() and dist.is_initialized():
                dist.barrier()
        args.policy_model_name_or_path = init_dir
    resume = ckpt if getattr(args, "resume_from_training", True) else None
    models = make_models(base_model, tokenizer, args, accelerator, resume_from_checkpoint=resume, model_factory=model_factory)
    trainer = B200DPOTrainer(args, models["policy"], models["ref_policy"], optimizer=None, world_size=world,
                             **(trainer_kwargs or {}))
    trainer.train(dataset_len, train_dataloader_factory, resume_training_ckpt=resume)
    if rank == 0:
        trainer.save_model(os.path.join(args.output_dir, "checkpoint-final"))
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    return trainer
