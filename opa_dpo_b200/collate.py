"""Host-side mirror of the reference's collator stage on the GPU kernels.

* `B200DataCollator`        <- utils/data_utils_dpo.py:101-287 `DataCollatorForCausalLM`: same constructor fields, same
                               batch keys / dtypes, same fallback when a report cannot be assembled. Tokenisation, JSON
                               parsing and the sentence re-alignment stay on the host (strings); padding, BOS / 29871
                               stripping, concatenation, EOS insertion, the per-token score / image-relation rows and
                               the attention masks are one kernel launch per response stream.
* `preprocess_images`       <- DPO_Dataset.__getitem__ image path (:311-335): expand2square + CLIP preprocess of decoded
                               uint8 RGB images, bit-identical to Pillow + the pinned HF image processor.
* `mask_images`             <- mask_single_image (opadpo/dpo_models/dpo_trainer.py:83-109; opa_trainer.py:66-72)
* `mask_percentage_per_row` <- dpo_trainer.py:119-125
The random permutations of the two CoPO functions are drawn with torch.randperm on the CPU generator in the reference's
order, so a seeded run masks the same pixels. All arithmetic happens in libopadpo_b200.so; no CPU fallback exists.
"""
from __future__ import annotations

import ctypes
import json
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import OpadpoError, Segments

SCORE_WEIGHT = {1: 2.5, 2: 2.0, 3: 1.5, 4: 1.0}                       # score_mapping, data_utils_dpo.py:32-37
RELATION_WEIGHT = {"image_recognition_error": 3.0, "correct": 1.0, "language_comprehension_error": 1.0}   # :38-42
IMAGE_PLACEHOLDER_ID = 30861                                          # :121
IMAGE_TOKEN_INDEX = -200
SEG_KEEP_BOS, SEG_STRIP_SPACE, SEG_VERBATIM = 1, 2, 4
ADD_EOS, PAD_EOS, TRUNCATE, TOKEN_MASK = 1, 2, 4, 8
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
_DESCRIPTION_KEYS = ("image_description", "image description")


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device(device) -> torch.device:
    if not torch.cuda.is_available():
        raise OpadpoError("opa_dpo_b200.collate needs a CUDA device: the collator kernels have no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


# ---------------------------------------------------------------------------------------------------- host: strings
def realign_copied(original: str, pieces: List[str]) -> List[str]:
    """Behaviour of complete_copied_content (data_utils_dpo.py:56-71): walk through `original`, and let every non-empty
    piece absorb the text between the previous match and its own end; the tail of `original` goes to the last piece. If a
    piece is not found the pieces are returned untouched."""
    out: List[str] = []
    rest = original
    for piece in pieces:
        core = piece.strip()
        if not core:
            out.append("")
            continue
        at = rest.find(core)
        if at < 0:
            return pieces
        cut = at + len(core)
        out.append(rest[:cut].strip(" "))
        rest = rest[cut:]
    tail = rest.strip()
    if out and tail:
        out[-1] += tail
    return out


def _entry(d: dict, spaced: str, snake: str, default):
    return d.get(spaced, d.get(snake, default))


class SegmentPack:
    """Flat arrays of one response stream of a batch: the input format of opadpo_op_collate_responses."""

    def __init__(self):
        self.tokens: List[int] = []
        self.seg_offsets: List[int] = [0]
        self.score: List[float] = []
        self.relation: List[float] = []
        self.flags: List[int] = []
        self.row_offsets: List[int] = [0]
        self.row_present: List[int] = []          # entries that tokenised to more than BOS, per row
        self.float_weights = False                # some present entry carries float weights (see report_segments)

    def add(self, raw: Sequence[int], score: float, relation: float, flags: int, integer_weights: bool = False):
        if len(raw) > 1 and not integer_weights:
            self.float_weights = True
        self.tokens.extend(int(t) for t in raw)
        self.seg_offsets.append(len(self.tokens))
        self.score.append(score)
        self.relation.append(relation)
        self.flags.append(flags)
        self._present += 1 if len(raw) > 1 else 0

    def begin_row(self):
        self._present = 0

    def end_row(self):
        self.row_offsets.append(len(self.flags))
        self.row_present.append(self._present)

    def arrays(self) -> Dict[str, np.ndarray]:
        return dict(tokens=np.asarray(self.tokens, np.int32), seg_offsets=np.asarray(self.seg_offsets, np.int64),
                    seg_score=np.asarray(self.score, np.float32), seg_relation=np.asarray(self.relation, np.float32),
                    seg_flags=np.asarray(self.flags, np.uint8), row_offsets=np.asarray(self.row_offsets, np.int64))


def plain_pack(token_rows: Sequence[Sequence[int]]) -> SegmentPack:
    """One segment per row holding the tokenizer output as is (the plain-tokenisation branch)."""
    p = SegmentPack()
    for raw in token_rows:
        p.begin_row()
        p.add(raw, 0.0, 0.0, SEG_KEEP_BOS | SEG_VERBATIM)
        p.end_row()
    return p


def report_segments(tokenize, reports: List[dict], originals: List[str]) -> Tuple[SegmentPack, SegmentPack]:
    """data_utils_dpo.py:141-199 up to the torch.cat: per report key the rewritten entry feeds the AI-pseudo stream and
    (unless the key is 'Added') the copied entry feeds the original-generation stream, with the sentence's score /
    error-type weights. `tokenize(text) -> list of ids including BOS`. Raises whatever the reference would raise on a
    malformed report (the caller turns that into the fallback)."""
    orig, ai = SegmentPack(), SegmentPack()
    for rep, original in zip(reports, originals):
        rep = {k: v for k, v in rep.items() if k not in _DESCRIPTION_KEYS}
        keys = list(rep.keys())
        copied = [(_entry(rep[k], "copied content", "copied_content", "") or "") for k in keys]
        for k, fixed in zip(keys, realign_copied(original, copied)):
            if fixed:
                if "copied content" in rep[k]:
                    rep[k]["copied content"] = fixed
                elif "copied_content" in rep[k]:
                    rep[k]["copied_content"] = fixed
        orig.begin_row()
        ai.begin_row()
        for count, k in enumerate(keys):
            e = rep[k]
            keep = SEG_KEEP_BOS if count == 0 else 0
            rewritten = tokenize(_entry(e, "rewritten content", "rewritten_content", ""))
            if k != "Added":
                cop = tokenize(_entry(e, "copied content", "copied_content", ""))
                ws = SCORE_WEIGHT.get(e.get("score", 4), 1.0)
                wr = RELATION_WEIGHT.get(_entry(e, "error type", "error_type", "correct"), 1.0)
                orig.add(cop, ws, wr, keep | SEG_STRIP_SPACE)
                ai.add(rewritten, ws, wr, keep)
            else:
                # 'Added' entries get torch.ones_like(ids) — INTEGER ones, not ones * float (:193-194); when every entry
                # that reaches a stream of the batch is of this kind, torch.cat / stack never promote and the reference's
                # score and image-relation rows come out int64
                ai.add(rewritten, 1.0, 1.0, keep, integer_weights=True)
        orig.end_row()
        ai.end_row()
    return orig, ai


# ---------------------------------------------------------------------------------------------------- device ops
def collate_responses(pack: SegmentPack, R: int, pad_id: int, eos_id: int, mode: int, device=None, weights: bool = True):
    """-> ids int64 [b,R], mask uint8 [b,R], scores f32 [b,R] | None, relations f32 [b,R] | None, overflow int32 [b]."""
    dev = _device(device)
    lib = _lib.load()
    a = {k: torch.from_numpy(v).to(dev) for k, v in pack.arrays().items()}
    b = len(pack.row_offsets) - 1
    ids = torch.empty((b, R), device=dev, dtype=torch.int64)
    mask = torch.empty((b, R), device=dev, dtype=torch.uint8)
    sc = torch.empty((b, R), device=dev, dtype=torch.float32) if weights else None
    rel = torch.empty((b, R), device=dev, dtype=torch.float32) if weights else None
    ov = torch.empty((b,), device=dev, dtype=torch.int32)
    s = Segments(a["tokens"].data_ptr(), a["seg_offsets"].data_ptr(), a["seg_score"].data_ptr(), a["seg_relation"].data_ptr(),
                 a["seg_flags"].data_ptr(), a["row_offsets"].data_ptr(), b)
    _lib.check(lib.opadpo_op_collate_responses(ctypes.byref(s), R, pad_id, eos_id, mode, ids.data_ptr(), mask.data_ptr(),
                                               None if sc is None else sc.data_ptr(), None if rel is None else rel.data_ptr(),
                                               ov.data_ptr(), _stream()), "opadpo_op_collate_responses")
    return ids, mask, sc, rel, ov


def collate_queries(token_rows: Sequence[Sequence[int]], Q: int, pad_id: int, device=None) -> Tuple[Tensor, Tensor]:
    dev = _device(device)
    lib = _lib.load()
    flat = torch.tensor([int(t) for r in token_rows for t in r] or [0], dtype=torch.int32).to(dev)
    off = torch.tensor(np.cumsum([0] + [len(r) for r in token_rows]), dtype=torch.int64).to(dev)
    b = len(token_rows)
    ids = torch.empty((b, Q), device=dev, dtype=torch.int64)
    mask = torch.empty((b, Q), device=dev, dtype=torch.uint8)
    _lib.check(lib.opadpo_op_collate_queries(flat.data_ptr(), off.data_ptr(), b, Q, pad_id, IMAGE_PLACEHOLDER_ID,
                                             IMAGE_TOKEN_INDEX, ids.data_ptr(), mask.data_ptr(), _stream()),
               "opadpo_op_collate_queries")
    return ids, mask


def preprocess_images(images: Sequence, size: int = 336, pad_to_square: bool = True, image_mean=CLIP_MEAN, image_std=CLIP_STD,
                      device=None) -> Tensor:
    """Decoded uint8 RGB images [H_i, W_i, 3] (numpy arrays or tensors, any sizes) -> float32 [n, 3, size, size] on the
    device, as `expand2square` + `processor.preprocess(...)['pixel_values']` give (data_utils_dpo.py:311-335)."""
    dev = _device(device)
    lib = _lib.load()
    arrs = [np.ascontiguousarray(i.cpu().numpy() if isinstance(i, Tensor) else i, dtype=np.uint8) for i in images]
    for a in arrs:
        if a.ndim != 3 or a.shape[2] != 3:
            raise OpadpoError(f"preprocess_images: expected [H, W, 3] uint8, got {a.shape}")
    n = len(arrs)
    offsets = np.cumsum([0] + [a.size for a in arrs]).astype(np.int64)
    host = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs]))
    pixels = host.pin_memory().to(dev, non_blocking=True)
    hs = np.asarray([a.shape[0] for a in arrs], np.int32)
    ws = np.asarray([a.shape[1] for a in arrs], np.int32)
    mean, std = np.asarray(image_mean, np.float32), np.asarray(image_std, np.float32)
    bg = np.asarray([int(x * 255) for x in image_mean], np.uint8)      # data_utils_dpo.py:328
    out = torch.empty((n, 3, size, size), device=dev, dtype=torch.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(lib.opadpo_op_preprocess_images(pixels.data_ptr(), p(offsets), p(hs), p(ws), n, size, p(mean), p(std), p(bg),
                                               int(pad_to_square), out.data_ptr(), _stream()), "opadpo_op_preprocess_images")
    return out


def mask_images(images: Tensor, mask_percentage: float, mask_method: str = "random", block: int = 14,
                indices: Optional[Tensor] = None, return_means: bool = False):
    """Batch form of `torch.stack([mask_single_image(images[i].unsqueeze(0), ratio, method) for i ...]).squeeze(1)`
    (dpo_trainer.py:294-296): images [n,C,H,W] float32 or bfloat16 on the device; returns the same dtype. One
    torch.randperm per image, in image order, on the CPU generator — the reference's random stream."""
    if mask_method not in ("random", "blockwise"):
        raise NotImplementedError(mask_method)
    dev = images.device
    if dev.type != "cuda":
        raise OpadpoError("mask_images: images must be on the CUDA device")
    lib = _lib.load()
    n, C, H, W = images.shape
    blk = 1 if mask_method == "random" else block
    total = H * W if blk == 1 else (H // blk) * (W // blk)
    m = int(total * mask_percentage)
    if indices is None:
        indices = torch.stack([torch.randperm(total)[:m] for _ in range(n)]) if n else torch.zeros((0, m), dtype=torch.int64)
    idx = indices.to(dev, torch.int64).contiguous()
    bf16 = images.dtype == torch.bfloat16
    src = images.to(torch.float32).contiguous()
    out = torch.empty_like(src)
    means = torch.empty((n,), device=dev, dtype=torch.float32)
    _lib.check(lib.opadpo_op_mask_images(src.data_ptr(), out.data_ptr(), n, C, H, W, idx.data_ptr(), idx.shape[1], blk, int(bf16),
                                         means.data_ptr(), _stream()), "opadpo_op_mask_images")
    out = out.to(images.dtype)
    return (out, means) if return_means else out


def mask_percentage_per_row(n_rows: int, n_cols: int, percentage: float, device=None, indices: Optional[Tensor] = None) -> Tensor:
    """dpo_trainer.py:119-125 on an all-True [n_rows, n_cols] matrix -> bool tensor on the device."""
    dev = _device(device)
    lib = _lib.load()
    m = int(n_cols * percentage)
    if indices is None:
        indices = torch.stack([torch.randperm(n_cols)[:m] for _ in range(n_rows)]) if n_rows else torch.zeros((0, m), dtype=torch.int64)
    idx = indices.to(dev, torch.int64).contiguous()
    out = torch.empty((n_rows, n_cols), device=dev, dtype=torch.uint8)
    _lib.check(lib.opadpo_op_attention_holes(out.data_ptr(), idx.data_ptr(), n_rows, n_cols, idx.shape[1], _stream()),
               "opadpo_op_attention_holes")
    return out.bool()


# ---------------------------------------------------------------------------------------------------- the collator
class B200DataCollator:
    """DataCollatorForCausalLM (data_utils_dpo.py:101-287) with the array work on the GPU. Instances are the dicts
    DPO_Dataset.__getitem__ returns; 'images' may be the preprocessed float tensor (as in the reference) or the decoded
    uint8 [H,W,3] image, which is then preprocessed on the device too. The batch lives on the device."""

    def __init__(self, tokenizer, query_len: int = 128, response_len: int = 896, detailed_report: bool = False, device=None,
                 image_size: int = 336, pad_to_square: bool = True, ops=None):
        """`ops`: the array operations (`collate_responses`, `collate_queries`, `preprocess_images`); the default — and the
        only implementation this package has — are the CUDA kernels of this module. The CPU tests pass the oracle here to
        check the host half (report parsing, fallback rules, batch keys / dtypes) against the reference collator."""
        self.tokenizer, self.query_len, self.response_len = tokenizer, query_len, response_len
        self.detailed_report, self.device = detailed_report, device
        self.image_size, self.pad_to_square = image_size, pad_to_square
        self.ops = ops if ops is not None else _DEVICE_OPS

    def _tokenize(self, text: str) -> List[int]:
        ids = self.tokenizer(text, return_tensors="pt")["input_ids"]
        return [int(t) for t in ids[0]]

    def _plain(self, texts: List[str]):
        t = self.tokenizer
        ids, mask, *_ = self.ops.collate_responses(plain_pack([self._tokenize(x) for x in texts]), self.response_len, t.pad_token_id,
                                                   t.eos_token_id, ADD_EOS | TRUNCATE | TOKEN_MASK, self.device, weights=False)
        return ids, mask.long()     # the tokenizer's attention mask after add_eos walked it too: 1..1, eos_id, 0..0 (:76-79)

    def __call__(self, instances: List[Dict]) -> Dict[str, Tensor]:
        t, R = self.tokenizer, self.response_len
        reports = [json.loads(i["AI_json_report"]) for i in instances] if self.detailed_report else None
        q_ids, q_mask = self.ops.collate_queries([self._tokenize(i["queries"]) for i in instances], self.query_len,
                                                 t.pad_token_id, self.device)
        std_ids, std_mask = self._plain([i["standard_response"] for i in instances])
        batch = {"queries": q_ids, "queries_attention_mask": q_mask.long(), "standard_response": std_ids,
                 "standard_response_attention_mask": std_mask}
        detailed = None
        if self.detailed_report:
            try:
                orig, ai = report_segments(self._tokenize, reports, [i["original_generate_response"] for i in instances])
                if min(orig.row_present + ai.row_present) == 0:
                    raise RuntimeError("a response stream without any entry (torch.cat of an empty list in the reference)")
                o = self.ops.collate_responses(orig, R, t.pad_token_id, t.eos_token_id, ADD_EOS, self.device)
                a = self.ops.collate_responses(ai, R, t.pad_token_id, t.eos_token_id, ADD_EOS | PAD_EOS, self.device)
                if bool(torch.cat([o[4], a[4]]).any().item()):
                    raise RuntimeError("a report is longer than response_len (pad_and_stack raises in the reference)")
                detailed = (o, a)
            except Exception as e:      # the reference prints and falls back to plain tokenisation with zero scores (:259-278)
                print(e)
        if detailed is not None:
            o, a = detailed
            if not ai.float_weights:           # only 'Added' entries in the whole batch: the reference's rows are int64
                a = (a[0], a[1], a[2].long(), a[3].long(), a[4])
            batch.update({"original_generate_response": o[0], "original_generate_response_attention_mask": o[1].bool(),
                          "AI_pseudo_response": a[0], "AI_pseudo_response_attention_mask": a[1].bool(),
                          "original_generate_response_scores": o[2], "AI_pseudo_response_scores": a[2],
                          "original_generate_response_image_relations": o[3], "AI_pseudo_response_image_relations": a[3]})
        else:
            o_ids, o_mask = self._plain([i["original_generate_response"] for i in instances])
            a_ids, a_mask = self._plain([i["AI_pseudo_response"] for i in instances])
            batch.update({"original_generate_response": o_ids, "AI_pseudo_response": a_ids})
            if self.detailed_report:
                z = torch.zeros_like(o_ids)
                batch.update({"original_generate_response_scores": z, "AI_pseudo_response_scores": torch.zeros_like(a_ids),
                              "original_generate_response_image_relations": z.clone(),
                              "AI_pseudo_response_image_relations": torch.zeros_like(a_ids)})
            batch["original_generate_response_attention_mask"] = o_mask
            batch["AI_pseudo_response_attention_mask"] = a_mask
        images = [i["images"] for i in instances]
        if all(isinstance(x, (np.ndarray, Tensor)) and x.dtype in (np.uint8, torch.uint8) for x in images):
            batch["images"] = self.ops.preprocess_images(images, self.image_size, self.pad_to_square, device=self.device)
        elif all(x is not None and x.shape == images[0].shape for x in images):
            batch["images"] = torch.stack([x if isinstance(x, Tensor) else torch.as_tensor(x) for x in images]).to(q_ids.device)
        else:
            batch["images"] = images
        return batch


class _DeviceOps:
    """The array operations B200DataCollator runs: this module's CUDA-kernel wrappers."""
    collate_responses = staticmethod(collate_responses)
    collate_queries = staticmethod(collate_queries)
    preprocess_images = staticmethod(preprocess_images)


_DEVICE_OPS = _DeviceOps()



class B200DPODataset:
    """DPO_Dataset (utils/data_utils_dpo.py:289-343) with the image work left to the device: same items, same query
    template, same keys — but 'images' is the DECODED uint8 RGB array [H, W, 3] (image file, `image_id`, or base64
    `image_bytes`, in the reference's order of precedence) instead of the CPU-preprocessed tensor; B200DataCollator then
    runs expand2square + CLIP preprocess for the whole batch on the GPU (`preprocess_images`), bit-identical to what the
    reference's processor produces. JPEG / PNG decoding itself stays with Pillow on the host."""

    query_templete_prefix = ("<s> A chat between a curious user and an artificial intelligence assistant. The assistant gives "
                             "helpful, detailed, and polite answers to the user's questions. USER: ")
    query_templete_suffix = " ASSISTANT: "

    def __init__(self, dataset, image_dir: Optional[str] = None):
        self.dataset = dataset
        self.image_dir = image_dir

    def __len__(self):
        return len(self.dataset)

    def _decode(self, item) -> np.ndarray:
        import base64
        import io
        import os
        from PIL import Image
        folder = self.image_dir if self.image_dir is not None else os.environ.get("IMAGE_DIR", "")
        if "images" in item:
            img = Image.open(os.path.join(folder, item["images"]))
        elif "image_id" in item:
            img = Image.open(os.path.join(folder, item["image_id"]))
        elif "image_bytes" in item:
            img = Image.open(io.BytesIO(base64.b64decode(item["image_bytes"].encode("utf-8"))))
        else:
            raise ValueError("No image found in the dataset")
        return np.asarray(img.convert("RGB"), dtype=np.uint8)

    def __getitem__(self, idx):
        item = self.dataset[idx]
        return {"queries": self.query_templete_prefix + item["queries"].replace("<image>", "图 ") + self.query_templete_suffix,
                "images": self._decode(item),
                "standard_response": item["standard_response"],
                "original_generate_response": item["original_generate_response"],
                "AI_pseudo_response": item["AI_pseudo_response"],
                "AI_json_report": item["AI_json_report"]}


def make_dpo_data_module(tokenizer, args, data_args, device=None, ops=None) -> Dict:
    """make_dpo_data_module (utils/data_utils_dpo.py:353-370), same signature and keys: the HF `datasets` directory at
    `data_args.data_path` (`load_from_disk`; columns of base_operations/make_opadpo_dataset.py:111-128), wrapped by the dataset
    mirror, and the collator with the reference's hard-coded lengths (query 128, response 896, :363-365) and
    `args.detailed_report`. Whether images are padded to a square follows `data_args.image_aspect_ratio == "pad"` (:320)."""
    from datasets import load_from_disk
    dataset = load_from_disk(data_args.data_path)
    pad = getattr(data_args, "image_aspect_ratio", "pad") == "pad"
    collator = B200DataCollator(tokenizer=tokenizer, query_len=128, response_len=896, detailed_report=args.detailed_report,
                                device=device, pad_to_square=pad, ops=ops)
    return dict(train_dataset=B200DPODataset(dataset), eval_dataset=None, data_collator=collator)
