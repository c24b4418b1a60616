"""Reading the reference's on-disk layouts into the engine ("Checkpoint layout", read side).

* base weights: HF-sharded `pytorch_model-*.bin` + `pytorch_model.bin.index.json` (what `loading_vision_tower_parameter`
  walks, opadpo/opadpo_train.py:539-557), a single `pytorch_model.bin`, or the safetensors equivalents;
* CLIP tower: the directory named by `config.json["image_checkpoint"]["image_model"]`
  (base_operations/modify_base_model_config.py:2-8), HF `CLIPVisionModel` names (`vision_model.*`); tensors of the LLaVA
  checkpoint whose names contain `vision_tower` / `mm_projector` are loaded afterwards and win, as in the reference
  (opadpo_train.py:705-711);
* adapters: `<path>/[adapter_model/][lora_default/]{adapter_config.json, adapter_model.bin}` (qlora_model.py:57-64,
  318-327) — the directory resolution only; the tensors go through `B200LlavaModel.load_adapter`.
Host-side I/O only: tensors are streamed shard by shard (one shard in host memory at a time) into anything that exposes
`load_tensor(name, tensor)` — the Engine, or a recorder in the CPU tests.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Tuple

import torch

from .engine import EngineConfig

_INDEX_FILES = ("pytorch_model.bin.index.json", "model.safetensors.index.json")
_SINGLE_FILES = ("pytorch_model.bin", "model.safetensors")


def resolve_adapter_dir(checkpoint_dir: str) -> str:
    """qlora_model.py:57-64 / 318-327: descend into adapter_model/ and then lora_default/ when they exist."""
    for sub in ("adapter_model", "lora_default"):
        nxt = os.path.join(checkpoint_dir, sub)
        if os.path.exists(nxt):
            checkpoint_dir = nxt
    return checkpoint_dir


def shard_files(model_dir: str) -> List[Tuple[str, Optional[List[str]]]]:
    """[(file, tensor names it holds or None = all)] in first-appearance order of the index's weight_map."""
    for idx in _INDEX_FILES:
        p = os.path.join(model_dir, idx)
        if os.path.exists(p):
            with open(p) as f:
                weight_map: Dict[str, str] = json.load(f).get("weight_map", {})
            files: Dict[str, List[str]] = {}
            for name, fn in weight_map.items():
                files.setdefault(fn, []).append(name)
            return [(fn, names) for fn, names in files.items()]
    for fn in _SINGLE_FILES:
        if os.path.exists(os.path.join(model_dir, fn)):
            return [(fn, None)]
    raise FileNotFoundError(f"no checkpoint index or weight file under {model_dir}")


def _load_file(path: str) -> Dict[str, torch.Tensor]:
    if path.endswith(".safetensors"):
        try:
            from safetensors.torch import load_file
        except ImportError as e:      # pragma: no cover - depends on the image
            raise RuntimeError(f"{path}: safetensors is not installed") from e
        return load_file(path, device="cpu")
    return torch.load(path, map_location="cpu", weights_only=True)


def iter_checkpoint_tensors(model_dir: str, want: Optional[Callable[[str], bool]] = None) -> Iterator[Tuple[str, torch.Tensor]]:
    """(name, tensor) of every tensor the index lists (or the file holds), one shard in memory at a time."""
    for fn, names in shard_files(model_dir):
        if names is not None and want is not None and not any(want(n) for n in names):
            continue
        sd = _load_file(os.path.join(model_dir, fn))
        for name in (names if names is not None else list(sd.keys())):
            if want is not None and not want(name):
                continue
            if name not in sd:
                raise KeyError(f"{fn} does not hold {name} although the index says so")
            yield name, sd[name]
        del sd


def clip_dir_from_config(base_dir: str) -> Optional[str]:
    """config.json["image_checkpoint"]["image_model"] (modify_base_model_config.py:2-8), else LLaVA's mm_vision_tower when it
    is a local directory."""
    with open(os.path.join(base_dir, "config.json")) as f:
        cfg = json.load(f)
    cand = (cfg.get("image_checkpoint") or {}).get("image_model") or cfg.get("mm_vision_tower")
    if cand and os.path.isdir(cand):
        return cand
    if cand and os.path.isdir(os.path.join(base_dir, cand)):
        return os.path.join(base_dir, cand)
    return None


def engine_config_from_hf(base_dir: str, clip_dir: Optional[str] = None, lora_r: int = 256, lora_alpha: float = 512.0) -> EngineConfig:
    """LLaVA `config.json` (+ the CLIP tower's) -> EngineConfig. The executed CLIP depth follows mm_vision_select_layer of
    the BASE checkpoint's config (-2 for LLaVA-1.5), not the training flag."""
    with open(os.path.join(base_dir, "config.json")) as f:
        c = json.load(f)
    kw = dict(d_model=c["hidden_size"], n_layers=c["num_hidden_layers"], n_heads=c["num_attention_heads"],
              d_ffn=c["intermediate_size"], vocab=c["vocab_size"], rms_eps=c.get("rms_norm_eps", 1e-5),
              rope_theta=c.get("rope_theta", 10000.0), max_positions=c.get("max_position_embeddings", 4096),
              lora_r=lora_r, lora_alpha=lora_alpha)
    if c.get("num_key_value_heads", c["num_attention_heads"]) != c["num_attention_heads"]:
        raise ValueError("grouped-query attention is not part of the LLaVA-1.5 path")
    clip_dir = clip_dir or clip_dir_from_config(base_dir)
    if clip_dir is not None:
        with open(os.path.join(clip_dir, "config.json")) as f:
            v = json.load(f)
        v = v.get("vision_config", v)
        select = c.get("mm_vision_select_layer", -2)
        n_hidden = v["num_hidden_layers"]
        executed = n_hidden + select + 1 if select < 0 else select        # hidden_states[select]: embeddings are index 0
        if not 0 < executed <= n_hidden:
            raise ValueError(f"mm_vision_select_layer {select} out of range for a {n_hidden}-layer tower")
        kw.update(v_d=v["hidden_size"], v_layers=executed, v_heads=v["num_attention_heads"], v_ffn=v["intermediate_size"],
                  v_image=v["image_size"], v_patch=v["patch_size"], v_eps=v.get("layer_norm_eps", 1e-5))
    return EngineConfig(**kw)


def load_llava_base(sink, base_dir: str, clip_dir: Optional[str] = None, log: Optional[Callable[[str], None]] = None,
                    required: Optional[Iterable[str]] = None) -> Dict[str, int]:
    """Stream a LLaVA-1.5 base checkpoint into `sink.load_tensor(name, tensor)`: first the CLIP tower from its own directory,
    then every tensor of the LLaVA shards (its vision_tower / mm_projector entries overriding the tower's, as
    loading_vision_tower_parameter does). Only names the engine does not hold (OPADPO_ENOENT: a text tower, buffers it has
    no use for) are counted as skipped; a shape or dtype mismatch on a tensor it does hold — a resized vocabulary, another
    hidden size or patch size — raises, as from_pretrained does. Afterwards every name in `required` (default: the sink's
    `required_tensor_names()`, i.e. everything the engine executes) must have been loaded, or the missing list is raised.
    Returns {"loaded", "skipped"}."""
    from ._lib import OpadpoError, OpadpoUnknownName
    stats = {"loaded": 0, "skipped": 0}
    seen = set()

    def put(name: str, t: torch.Tensor):
        try:
            sink.load_tensor(name, t)
            stats["loaded"] += 1
            seen.add(name)
        except OpadpoUnknownName:
            stats["skipped"] += 1
            if log:
                log(f"skipped {name} {tuple(t.shape)}")

    clip_dir = clip_dir or clip_dir_from_config(base_dir)
    if clip_dir is not None:
        for name, t in iter_checkpoint_tensors(clip_dir, lambda n: n.startswith("vision_model.")):
            put("model.vision_tower.vision_tower." + name, t)
    for name, t in iter_checkpoint_tensors(base_dir):
        put(name, t)
    if required is None and hasattr(sink, "required_tensor_names"):
        required = sink.required_tensor_names()
    missing = [n for n in (required or []) if n not in seen]
    if missing:
        raise OpadpoError(f"base checkpoint {base_dir} does not provide {len(missing)} tensors the engine executes: "
                          + ", ".join(missing[:8]) + (" ..." if len(missing) > 8 else ""))
    return stats
