"""Outer training loop (RLTrainer.train, rl_trainer.py:217-279): step numbering, save cadence, resume
skipping, max_step cut. The reference's own `RLTrainer.train` is run VERBATIM here (imported from /root/reference with the
package stubs of oracle/make_golden.py) over a recording subclass; where /root/reference is absent (the GPU box) the same
event logs come from tests/golden/train_loop_events.json; when the reference is present the committed file is also checked
against a live run (OPADPO_REGEN_GOLDEN=1 rewrites it)."""
import json
import os
import sys
from types import SimpleNamespace

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "train_loop_events.json")

CASES = {
    "plain": dict(n=40, epochs=2, rbs=8, acc=2, save=3, extra=[], evals=None, max_step=100, resume=None),
    "max_step_cut": dict(n=64, epochs=4, rbs=8, acc=1, save=5, extra=[7], evals=4, max_step=12, resume=None),
    "resume": dict(n=48, epochs=2, rbs=8, acc=3, save=2, extra=[5], evals=None, max_step=100, resume="out/checkpoint-6"),
    "resume_on_save_step": dict(n=48, epochs=1, rbs=8, acc=1, save=2, extra=[], evals=None, max_step=100, resume="out/checkpoint-4"),
    "empty": dict(n=5, epochs=1, rbs=8, acc=1, save=1, extra=[], evals=None, max_step=10, resume=None),
}


def _args(c):
    return SimpleNamespace(total_epochs=c["epochs"], rollout_batch_size=c["rbs"], rollout_accumulation_steps=c["acc"],
                           save_steps=c["save"], save_steps_extra_list=c["extra"], eval_steps=c["evals"], output_dir="out")


class _Recorder:
    """The calls train() makes, in order."""

    def __init__(self, c):
        self.events, self.c, self.batches = [], c, 0

    def loader(self):
        def gen():
            while True:
                self.batches += 1
                yield self.batches
        return gen()


def _reference_events(c):
    import re
    from oracle.make_golden import install_stubs
    install_stubs()
    from opadpo.dpo_models import rl_trainer as R
    rec = _Recorder(c)

    class Sub(R.RLTrainer):
        def rollout(self, q): ...
        def compute_loss(self, r): ...
        def record_step_stats(self, *a, **k): ...
        def create_optimizer_and_scheduler(self, n):
            rec.events.append(["sched", n])
        def resume_training(self, d):
            rec.events.append(["resume", d])
            return int(re.search(r"checkpoint-(\d+)", d).group(1))
        def get_train_dataloader(self):
            return rec.loader()
        def step(self, loader, step_idx):
            got = [next(loader) for _ in range(self.args.rollout_accumulation_steps)]
            rec.events.append(["step", step_idx, got])
            return {"step": step_idx}
        def save_model(self, d=None):
            rec.events.append(["save", d])
        def evaluate(self, step_idx, unwrapped_policy=None):
            rec.events.append(["eval", step_idx])

    t = object.__new__(Sub)
    t.args, t.max_step, t.train_dataset = _args(c), c["max_step"], [0] * c["n"]
    t.accelerator = SimpleNamespace(is_main_process=False)
    t.log_history = []
    hist = t.train(resume_training_ckpt=c["resume"])
    return rec.events + [["history", [h["step"] for h in hist]]]


def _mirror_events(c):
    import re
    from opa_dpo_b200.trainer import train_loop
    rec = _Recorder(c)

    class T:
        args, max_step = _args(c), c["max_step"]
        def create_optimizer_and_scheduler(self, n):
            rec.events.append(["sched", n])
        def resume_training(self, d):
            rec.events.append(["resume", d])
            return int(re.search(r"checkpoint-(\d+)", d).group(1))
        def step(self, loader, step_idx):
            got = [next(loader) for _ in range(self.args.rollout_accumulation_steps)]
            rec.events.append(["step", step_idx, got])
            return {"step": step_idx}
        def save_model(self, d):
            rec.events.append(["save", d])
        def evaluate(self, step_idx):
            rec.events.append(["eval", step_idx])

    hist = train_loop(T(), c["n"], rec.loader, c["resume"])
    return rec.events + [["history", [h["step"] for h in hist]]]


def test_train_loop_matches_reference_rl_trainer():
    with open(GOLD) as f:
        gold = json.load(f)
    if os.path.isdir("/root/reference"):       # the committed golden must be what the reference does today
        live = json.loads(json.dumps({name: _reference_events(c) for name, c in CASES.items()}))
        if live != gold and os.environ.get("OPADPO_REGEN_GOLDEN") == "1":
            with open(GOLD, "w") as f:
                json.dump(live, f, indent=1, sort_keys=True)
            gold = live
        assert live == gold, "tests/golden/train_loop_events.json is stale (OPADPO_REGEN_GOLDEN=1 rewrites it)"
    assert set(gold) == set(CASES)
    for name, c in CASES.items():
        assert _mirror_events(c) == gold[name], name
    # what the cases pin: a checkpoint is written at the START of steps 3, 6, ...; never at the step a run resumed from;
    # skipped steps still consume their batches; max_step cuts before the step with that index runs
    assert ["save", "out/checkpoint-3"] in gold["plain"] and gold["plain"][-1][1][-1] == 10
    assert ["save", "out/checkpoint-4"] not in gold["resume_on_save_step"]
    assert gold["max_step_cut"][-1][1][-1] == 11
    assert gold["resume"][2][:2] == ["step", 6] and gold["resume"][2][2] == [16, 17, 18]
    assert gold["empty"] == [["sched", 0], ["history", []]]
