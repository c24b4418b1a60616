"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/opadpo_b200.h declares; the product path fails loudly without a GPU (no CPU fallback); host-side mirrors keep
the reference's calling conventions. No compute calls are made here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from opa_dpo_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "opadpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(opadpo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from opa_dpo_b200 import _lib
    declared = _header_symbols()
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(raw, s)]
    assert not missing, f"declared in include/opadpo_b200.h but not exported: {missing}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, (set(declared) ^ set(_lib.EXPORTED_SYMBOLS))
    assert lib.opadpo_abi_version() == _lib.EXPECTED_ABI


def test_struct_layouts_match_header_field_order():
    """ctypes mirrors must list the same fields in the same order as the C structs."""
    from opa_dpo_b200 import _lib
    src = open(os.path.join(ROOT, "include", "opadpo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", re.sub(r"\[.*?\]", "", part.strip()).replace("*", " "))[0])
        return names

    for struct, cls in (("opadpo_gemm_desc", _lib.GemmDesc), ("opadpo_attn_desc", _lib.AttnDesc),
                        ("opadpo_dpo_pair_desc", _lib.DpoPairDesc), ("opadpo_config", _lib.Config),
                        ("opadpo_sampler", _lib.Sampler), ("opadpo_allreduce_desc", _lib.AllReduceDesc)):
        assert fields(struct) == [f[0] for f in cls._fields_], struct
    assert ctypes.sizeof(_lib.Sampler) == 24 + 4 + 4 * 8 + 4      # uint64 alignment pads the tail
    # rank, world | grad, stage_local, stage_mc | 8 + 8 peer addresses | count | epoch, blocks | err_flag | timeout_ms, fp32_in_place
    assert ctypes.sizeof(_lib.AllReduceDesc) == 8 + 3 * 8 + 16 * 8 + 8 + 8 + 8 + 8
    assert _lib.MAX_RANKS == 8 and _lib.AR_MAX_BLOCKS == 256 and "#define OPADPO_MAX_RANKS 8" in src and \
        "#define OPADPO_AR_MAX_BLOCKS 256" in src


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu(lib):
    from opa_dpo_b200 import Engine, EngineConfig, OpadpoError, _lib
    with pytest.raises(OpadpoError):
        Engine(EngineConfig())
    cfg = EngineConfig().to_c()
    h = ctypes.c_void_p()
    rc = lib.opadpo_engine_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc < 0 and b"no CUDA device" in lib.opadpo_last_error()
    assert lib.opadpo_forward_logprobs(None, 0, None, 1, None, None, None, 1, 1, None, 1, 1.0, None, None, -1, None) < 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opa_dpo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f"{f} imports oracle/"


def test_policy_mirror_discovers_responses_like_the_reference():
    """rl_models.py:91-92 key filter, response_len check and the PEFT key naming of saved adapters."""
    from opa_dpo_b200.engine import EngineConfig, linear_modules
    keys = ["standard_response", "AI_pseudo_response_scores", "original_generate_response", "x_response_mask",
            "AI_pseudo_response_image_relations", "mask_AI_pseudo_response", "mode"]
    got = [k for k in keys if "response" in k and "_mask" not in k and "scores" not in k and "image_relations" not in k]
    assert got == ["standard_response", "original_generate_response", "mask_AI_pseudo_response"]
    mods = linear_modules(EngineConfig())
    assert len(mods) == 32 * 7 + 2 + 23 * 6
    assert "model.layers.31.mlp.down_proj" in mods and mods["model.layers.0.mlp.gate_proj"] == (11008, 4096)
    assert not any("lm_head" in m for m in mods)
    n_llm = sum(256 * (o + i) for m, (o, i) in mods.items() if "vision_model" not in m)
    assert n_llm == 643_039_232   # 639.6 M (LLM) + 3.4 M (projector): SURVEY.md §8e


def test_flop_model_matches_baseline_table():
    from opa_dpo_b200.engine import EngineConfig
    from opa_dpo_b200.flops import flops_per_pair
    assert abs(flops_per_pair(EngineConfig(), 128, 512)["per_pair"] / 1e12 - 113.3) < 0.1
    assert abs(flops_per_pair(EngineConfig(), 128, 128)["per_pair"] / 1e12 - 77.1) < 0.1
    assert abs(flops_per_pair(EngineConfig.llava_13b(), 128, 1024)["per_pair"] / 1e12 - 305.9) < 0.1


def test_cosine_schedule_matches_hf_and_last_checkpoint(tmp_path):
    """create_scheduler (utils/trainer_utils.py:41-49) == HF cosine with warm-up; get_last_checkpoint (lora_utils.py:31-47)."""
    from transformers.optimization import get_cosine_schedule_with_warmup
    from opa_dpo_b200.optim import cosine_schedule, get_last_checkpoint

    class FakeOpt:
        param_groups = [dict(lr=1e-6, initial_lr=1e-6)]

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-6)
    hf = get_cosine_schedule_with_warmup(opt, 5, 120)
    mine = cosine_schedule(FakeOpt(), 5, 120)
    for _ in range(120):
        assert abs(hf.get_last_lr()[0] - mine.get_last_lr()[0]) < 1e-18
        opt.step(); hf.step(); mine.step()
    assert get_last_checkpoint(str(tmp_path)) is None
    for n in (5, 40, 100):
        (tmp_path / f"checkpoint-{n}").mkdir()
    (tmp_path / "checkpoint-final").mkdir()
    assert get_last_checkpoint(str(tmp_path)).endswith("checkpoint-100")
    (tmp_path / "completed").write_text("")
    assert get_last_checkpoint(str(tmp_path)) is None


def test_scheduler_types_and_warmup_match_hf_get_scheduler():
    """create_scheduler (utils/trainer_utils.py:41-49) = transformers.get_scheduler(args.lr_scheduler_type, optimizer,
    args.get_warmup_steps(n), n): every schedule the flag can name without extra arguments, and the warm-up rule."""
    from types import SimpleNamespace
    from transformers import SchedulerType
    from transformers.optimization import get_scheduler
    from opa_dpo_b200.optim import cosine_schedule, warmup_steps_of
    for kind in ("cosine", "linear", "constant", "constant_with_warmup"):
        for warm, total in ((5, 60), (0, 7), (9, 9)):
            fake = SimpleNamespace(param_groups=[dict(lr=3e-5, initial_lr=3e-5)])
            opt = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=3e-5)
            hf = get_scheduler(kind, optimizer=opt, num_warmup_steps=warm, num_training_steps=total)
            mine = cosine_schedule(fake, warm, total, kind=SchedulerType(kind))
            for _ in range(total + 3):
                assert abs(hf.get_last_lr()[0] - mine.get_last_lr()[0]) < 1e-18, (kind, warm, total)
                opt.step(); hf.step(); mine.step()
            again = cosine_schedule(SimpleNamespace(param_groups=[dict(lr=3e-5, initial_lr=3e-5)]), 1, 2)
            again.load_state_dict(mine.state_dict())
            assert again.get_last_lr() == mine.get_last_lr() and again.kind == kind
    with pytest.raises(ValueError):
        cosine_schedule(SimpleNamespace(param_groups=[dict(lr=1.0, initial_lr=1.0)]), 0, 10, kind="polynomial")
    # TrainingArguments.get_warmup_steps: warmup_steps wins when positive, else ceil(n * warmup_ratio)
    # (the rule of the pinned transformers 4.34.1, training_args.py get_warmup_steps; constructing TrainingArguments needs accelerate)
    for ws, ratio, n, want in ((5, 0.0, 100, 5), (0, 0.03, 100, 3), (0, 0.03, 101, 4), (7, 0.5, 10, 7), (0, 0.0, 10, 0)):
        assert warmup_steps_of(SimpleNamespace(warmup_steps=ws, warmup_ratio=ratio), n) == want, (ws, ratio, n)
    assert warmup_steps_of(SimpleNamespace(get_warmup_steps=lambda n: 11), 100) == 11


def test_importing_the_whole_package_does_not_load_the_oracle():
    """oracle/ is test infrastructure: no module of the package imports it (statically), and importing the whole package
    does not pull it in."""
    import ast
    import glob
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in glob.glob(os.path.join(root, "opa_dpo_b200", "*.py")):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
    code = ("import sys, importlib, pkgutil, opa_dpo_b200\n"
            "for m in pkgutil.iter_modules(opa_dpo_b200.__path__):\n"
            "    importlib.import_module('opa_dpo_b200.' + m.name)\n"
            "assert not any(k == 'oracle' or k.startswith('oracle.') for k in sys.modules), 'oracle imported'\n")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root)


def test_collator_abi_argument_errors_without_gpu():
    """Error convention of the collator-stage entry points (include/opadpo_b200.h): bad arguments return OPADPO_EINVAL (-1)
    with a message, before anything touches CUDA — so this runs on a machine without a GPU."""
    import ctypes
    from opa_dpo_b200 import _lib
    lib = _lib.load()
    lib.opadpo_last_error.restype = ctypes.c_char_p
    seg = _lib.Segments(None, None, None, None, None, None, 2)
    assert lib.opadpo_op_collate_responses(ctypes.byref(seg), 8, 0, 2, 1, None, None, None, None, None, None) == -1
    assert b"collate_responses" in lib.opadpo_last_error()
    assert lib.opadpo_op_collate_queries(None, None, 1, 8, 0, 30861, -200, None, None, None) == -1
    assert lib.opadpo_op_mask_images(None, None, 1, 3, 28, 28, None, 0, 1, 0, None, None) == -1
    buf = (ctypes.c_float * 4)()
    assert lib.opadpo_op_mask_images(buf, buf, 1, 3, 28, 30, None, 0, 14, 0, None, None) == -1      # block must divide H and W
    assert b"block" in lib.opadpo_last_error()
    assert lib.opadpo_op_attention_holes(None, None, 1, 16, 0, None) == -1
    assert lib.opadpo_op_preprocess_images(None, None, None, None, 1, 336, None, None, None, 1, None, None) == -1
    assert lib.opadpo_op_truncate_after_eos(None, None, 1, 8, 2, 0, None, 0, None) == -1
    ids = (ctypes.c_int64 * 8)()
    assert lib.opadpo_op_truncate_after_eos(ids, ids, 1, 8, 2, 0, None, 3, None) == -1               # n_extra > 0 without the array
    assert lib.opadpo_backward_ent(None, 0, None, None, None) == -1 and b"null engine" in lib.opadpo_last_error()


def test_reference_arm_full_depth_step_is_isolated(monkeypatch):
    """bench.py --impl reference: the full-depth configs[0] CPU step runs in a child process behind a host-memory check and a
    time box, so a small or slow host loses that figure, never the reference line."""
    import bench
    avail = bench._available_host_gb()
    assert avail is None or avail > 0
    monkeypatch.setattr(bench, "_available_host_gb", lambda: 8.0)
    out = bench.full_depth_config0_isolated(4, 5.0)
    assert "skipped" in out and "8 GB available" in out["skipped"]
    monkeypatch.setattr(bench, "_available_host_gb", lambda: None)
    calls = {}

    class P:
        returncode, stdout, stderr = 0, 'noise\n{"seconds_per_step": 2.0, "pairs_per_s_at_R512_by_flops": 0.5}\n', ""

    def fake_run(cmd, **kw):
        calls["cmd"], calls["kw"] = cmd, kw
        return P()
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    assert bench.full_depth_config0_isolated(4, 7.0)["pairs_per_s_at_R512_by_flops"] == 0.5
    assert calls["cmd"][-2:] == ["--full-depth-child", "4"] and calls["kw"]["timeout"] == 7.0

    def killed(cmd, **kw):
        class Q:
            returncode, stdout, stderr = -9, "", "Killed"
        return Q()
    monkeypatch.setattr(bench.subprocess, "run", killed)
    assert "child exited -9" in bench.full_depth_config0_isolated(4, 7.0)["failed"]

    def slow(cmd, **kw):
        raise bench.subprocess.TimeoutExpired(cmd, kw["timeout"])
    monkeypatch.setattr(bench.subprocess, "run", slow)
    assert "time box" in bench.full_depth_config0_isolated(4, 7.0)["failed"]
