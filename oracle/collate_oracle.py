"""CPU oracle of the collator stage — TEST INFRASTRUCTURE ONLY.

A numpy restatement of what the reference does between the tokenizer / image decoder and the policy forward. Only
tests/, __graft_entry__.smoke() and the cpu_baseline legs of bench.py / tools/bench_*.py may import it; the product
(opa_dpo_b200/) never does.
Pinned by tests/golden/collate_golden.npz, which oracle/make_golden_collate.py produced by running the reference's own
code here (DataCollatorForCausalLM.__call__ over a deterministic stub tokenizer, DPO_Dataset.__getitem__ over
HuggingFace's PIL-backed CLIP image processor, dpo_trainer.mask_single_image / mask_percentage_per_row).

Reference lines restated (paths under /root/reference):
  * response assembly from the GPT-4V report   utils/data_utils_dpo.py:139-235   (detailed_report branch)
  * plain tokenisation branch / fallback       utils/data_utils_dpo.py:120-137, 259-278
  * pad_and_stack / add_eos / pad_eos          utils/data_utils_dpo.py:44-99
  * queries: left padding, 30861 -> -200       utils/data_utils_dpo.py:118-121
  * expand2square + CLIP preprocess            utils/data_utils_dpo.py:311-333
  * CoPO masking                               opadpo/dpo_models/dpo_trainer.py:83-125 (also opa_models/opa_trainer.py:126-141)
Third-party arithmetic that is not in /root/reference and is restated from its published algorithm:
  * Pillow's ImagingResample (bicubic, 8 bits per channel, two passes with 22-bit fixed-point coefficients) [UPSTREAM
    Pillow; environment.yaml pins pillow 10.x, the container has 12.2 — the 8bpc resampler is unchanged between them];
    checked bit-for-bit against PIL.Image.resize in tests/test_collate_cpu.py.
  * transformers 4.34.1 CLIPImageProcessor.preprocess [UPSTREAM]: PIL resize(shortest_edge) -> center_crop ->
    rescale (uint8 * (1/255) in float64, cast to float32) -> normalize ((x - mean) / std in float32).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

SCORE_MAPPING = {1: 2.5, 2: 2.0, 3: 1.5, 4: 1.0}                                   # data_utils_dpo.py:32-37
ERROR_TYPE_MAPPING = {"image_recognition_error": 3.0, "correct": 1.0, "language_comprehension_error": 1.0}   # :38-42
SPACE_TOKEN = 29871          # the '' piece the LLaMA tokenizer puts in front of a segment (data_utils_dpo.py:186)
IMAGE_PLACEHOLDER_ID = 30861  # id of the placeholder glyph the dataset puts where <image> was (:121, :341)
IMAGE_TOKEN_INDEX = -200
FLAG_KEEP_BOS, FLAG_STRIP_SPACE, FLAG_VERBATIM = 1, 2, 4
COLLATE_ADD_EOS, COLLATE_PAD_EOS, COLLATE_TRUNCATE, COLLATE_TOKEN_MASK = 1, 2, 4, 8
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


# ------------------------------------------------------------------------------------------------ token side
class Overflow(Exception):
    """pad_and_stack with a row longer than max_length: torch.full gets a negative size and raises; the reference's
    collator catches it and falls back to plain tokenisation with all-zero scores (data_utils_dpo.py:259-278)."""


def effective_segment(raw: Sequence[int], flags: int) -> Optional[np.ndarray]:
    """One `copied content` / `rewritten content` entry as the reference slices it (data_utils_dpo.py:166-186):
    raw = tokenizer(text)['input_ids'][0]; absent when it is only BOS; BOS is dropped unless this is the first key of the
    report; a copied segment then loses a leading 29871."""
    raw = np.asarray(raw, dtype=np.int64)
    if flags & FLAG_VERBATIM:       # plain tokenisation (:120-137): whatever the tokenizer returned, a lone BOS included
        return raw
    if raw.size <= 1:
        return None
    t = raw if (flags & FLAG_KEEP_BOS) else raw[1:]
    if (flags & FLAG_STRIP_SPACE) and t[0] == SPACE_TOKEN:
        t = t[1:]
    return t


def collate_responses(rows: List[List[Tuple[Sequence[int], float, float, int]]], R: int, pad_id: int, eos_id: int,
                      mode: int):
    """rows[i] = [(raw token ids, score weight, image-relation weight, flags), ...] in report-key order.
    -> ids int64 [b,R], attention mask uint8 [b,R], scores f32 [b,R], relations f32 [b,R], overflow bool [b]
    (data_utils_dpo.py:200-235). mode: COLLATE_ADD_EOS (add_eos :73-87), COLLATE_PAD_EOS (pad_eos :89-99),
    COLLATE_TRUNCATE (tokenizer truncation of the plain branch instead of the overflow exception), COLLATE_TOKEN_MASK."""
    b = len(rows)
    ids = np.full((b, R), pad_id, np.int64)
    sc = np.zeros((b, R), np.float32)
    rel = np.zeros((b, R), np.float32)
    overflow = np.zeros((b,), bool)
    tok_mask = np.zeros((b, R), np.uint8)
    for i, segs in enumerate(rows):
        toks, s, r = [], [], []
        for raw, ws, wr, fl in segs:
            t = effective_segment(raw, fl)
            if t is None:
                continue
            toks.append(t)
            s.append(np.full(t.shape, ws, np.float32))
            r.append(np.full(t.shape, wr, np.float32))
        t = np.concatenate(toks) if toks else np.zeros((0,), np.int64)
        s = np.concatenate(s) if s else np.zeros((0,), np.float32)
        r = np.concatenate(r) if r else np.zeros((0,), np.float32)
        if t.size > R:
            if not (mode & COLLATE_TRUNCATE):
                overflow[i] = True
            t, s, r = t[:R], s[:R], r[:R]
        ids[i, :t.size], sc[i, :t.size], rel[i, :t.size] = t, s, r
        tok_mask[i, :t.size] = 1
        if (mode & COLLATE_ADD_EOS) and t.size < R:
            # add_eos walks EVERY tensor of the BatchEncoding, the attention mask included (data_utils_dpo.py:76-79): the
            # mask's first 0 is overwritten with eos_id too, so the plain branch's mask reads 1..1, 2, 0..0
            tok_mask[i, t.size] = eos_id
        if mode & COLLATE_ADD_EOS:                      # first pad position becomes EOS
            pads = np.nonzero(ids[i] == pad_id)[0]
            if pads.size:
                ids[i, pads[0]] = eos_id
        if mode & COLLATE_PAD_EOS:                      # the EOS position inherits the previous token's weight
            e = np.nonzero(ids[i] == eos_id)[0]
            if e.size:
                for arr in (sc, rel):
                    prev = arr[i, e[0] - 1]              # e[0] == 0 wraps to the last position, as in the reference
                    arr[i, e[0]] = prev if prev != 0 else 1.0
    # the plain branch hands out the TOKENIZER's attention mask (made before add_eos); the report branch ids != pad (:253-255)
    return ids, (tok_mask if mode & COLLATE_TOKEN_MASK else (ids != pad_id).astype(np.uint8)), sc, rel, overflow


def collate_queries(rows: List[Sequence[int]], Q: int, pad_id: int):
    """tokenizer(queries, padding='max_length', truncation=True, max_length=Q) with padding_side='left', then
    input_ids[input_ids == 30861] = IMAGE_TOKEN_INDEX (data_utils_dpo.py:118-121)."""
    b = len(rows)
    ids = np.full((b, Q), pad_id, np.int64)
    mask = np.zeros((b, Q), bool)
    for i, t in enumerate(rows):
        t = np.asarray(t, np.int64)[:Q]
        if t.size:
            ids[i, Q - t.size:] = t
            mask[i, Q - t.size:] = True
    ids[ids == IMAGE_PLACEHOLDER_ID] = IMAGE_TOKEN_INDEX
    return ids, mask


# ------------------------------------------------------------------------------------------------ image side
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bicubic filter (support 2) over the whole input."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.int32)
    bounds = np.zeros((out_size, 2), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255)


def pil_resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C]; horizontal pass, uint8 rounding, vertical pass (ImagingResample)."""
    H, W, C = img.shape
    x = img.astype(np.int64)
    if W != out_w:
        kk, b = resample_coeffs(W, out_w)
        t = np.zeros((H, out_w, C), np.int64)
        for xx in range(out_w):
            x0, n = b[xx]
            t[:, xx, :] = (1 << (PRECISION_BITS - 1)) + (x[:, x0:x0 + n, :] * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1)
        x = _clip8(t)
    if H != out_h:
        kk, b = resample_coeffs(H, out_h)
        t = np.zeros((out_h, x.shape[1], C), np.int64)
        for yy in range(out_h):
            y0, n = b[yy]
            t[yy] = (1 << (PRECISION_BITS - 1)) + (x[y0:y0 + n] * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0)
        x = _clip8(t)
    return x.astype(np.uint8)


def expand2square(img: np.ndarray, background: Sequence[int]) -> np.ndarray:
    """data_utils_dpo.py:312-326: paste the image centred on a square of the background colour."""
    H, W, C = img.shape
    if W == H:
        return img
    L = max(H, W)
    out = np.empty((L, L, C), np.uint8)
    out[:] = np.asarray(background, np.uint8)
    if W > H:
        y0 = (W - H) // 2
        out[y0:y0 + H] = img
    else:
        x0 = (H - W) // 2
        out[:, x0:x0 + W] = img
    return out


def clip_preprocess(img: np.ndarray, size: int = 336, pad_to_square: bool = True, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """uint8 RGB [H,W,3] -> float32 [3,size,size] as DPO_Dataset.__getitem__ produces it (data_utils_dpo.py:311-335)."""
    if pad_to_square:
        img = expand2square(img, tuple(int(x * 255) for x in mean))
    H, W, _ = img.shape
    short, long_ = (W, H) if W <= H else (H, W)
    new_short, new_long = size, int(size * long_ / short)
    new_w, new_h = (new_short, new_long) if W <= H else (new_long, new_short)
    r = pil_resize_bicubic(img, new_w, new_h)
    top, left = (new_h - size) // 2, (new_w - size) // 2
    r = r[top:top + size, left:left + size]
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)                   # rescale
    x = (x - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)        # normalize, float32
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def image_mean(image: np.ndarray, bf16_pixels: bool = False) -> np.float32:
    """`image.mean()` of dpo_trainer.py:85. Floating-point: torch sums in float32 in an order that is not part of its
    contract, so this value (float64 accumulation) agrees with the reference's to a few float32 ulps of mean(|x|), not
    bit-for-bit; for bf16 images both round to the same bf16 value except at a rounding tie."""
    img = image.astype(np.float32)
    if bf16_pixels:
        img = _round_bf16(img)
    m = np.float32(img.astype(np.float64).mean())
    if bf16_pixels:
        m = _round_bf16(np.asarray([m], np.float32))[0]
    return m


def mask_single_image(image: np.ndarray, indices: np.ndarray, method: str = "random", block: int = 14,
                      bf16_pixels: bool = False, mean_value: Optional[float] = None) -> np.ndarray:
    """dpo_trainer.py:83-109 with the permutation prefix given (`indices` = torch.randperm(...)[:n_mask]): float32
    [C,H,W]; every channel of the selected pixels (or 14x14 blocks) is set to the mean over the whole image.
    bf16_pixels: the trainer casts the images to bf16 before masking (dpo_trainer.py:283-286), so the mean is taken over
    the bf16-rounded pixels (fp32 accumulation) and rounded to bf16 itself. `mean_value` overrides the fill value (tests
    pass the reference's own mean so that everything else can be compared bit-for-bit)."""
    img = image.astype(np.float32).copy()
    if bf16_pixels:
        img = _round_bf16(img)
    C, H, W = img.shape
    m = image_mean(image, bf16_pixels) if mean_value is None else np.float32(mean_value)
    if method == "random":
        flat = img.reshape(C, -1)
        flat[:, indices] = m
    elif method == "blockwise":
        wb = W // block
        for idx in np.asarray(indices):
            h, w = int(idx) // wb, int(idx) % wb
            img[:, h * block:(h + 1) * block, w * block:(w + 1) * block] = m
    else:
        raise NotImplementedError(method)
    return img


def attention_holes(n_cols: int, indices: np.ndarray) -> np.ndarray:
    """mask_percentage_per_row (dpo_trainer.py:119-125): all-True [n, n_cols] with indices[i] set False in row i."""
    n = indices.shape[0]
    m = np.ones((n, n_cols), bool)
    for i in range(n):
        m[i, indices[i]] = False
    return m


def _round_bf16(x: np.ndarray) -> np.ndarray:
    """round-to-nearest-even float32 -> bf16 -> float32."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


# ------------------------------------------------------------------------------------------------ stub tokenizer
def word_id(w: str) -> int:
    """Deterministic word -> id map of the stub tokenizer; "SPACE" / "IMG" give the two ids the collator treats specially."""
    import hashlib
    if w == "SPACE":
        return SPACE_TOKEN
    if w == "IMG":
        return IMAGE_PLACEHOLDER_ID
    h = int(hashlib.md5(w.encode()).hexdigest()[:8], 16)
    v = 3 + h % 29000
    return v + 1 if v in (SPACE_TOKEN, IMAGE_PLACEHOLDER_ID) else v


class StubTokenizer:
    """The slice of the HF tokenizer interface the collator uses."""
    pad_token_id, eos_token_id, bos_token_id = 0, 2, 1

    def __init__(self):
        self.padding_side = "right"
        self.log = []          # (text, ids) of every single-string call, in call order

    def encode_one(self, text: str):
        return [self.bos_token_id] + [word_id(w) for w in text.split()]

    def __call__(self, text, padding=False, truncation=False, max_length=None, return_tensors=None):
        import torch
        from transformers import BatchEncoding
        if isinstance(text, str):
            ids = self.encode_one(text)
            self.log.append((text, list(ids)))
            return BatchEncoding({"input_ids": torch.tensor([ids], dtype=torch.long),
                                  "attention_mask": torch.ones(1, len(ids), dtype=torch.long)})
        rows, masks = [], []
        for t in text:
            ids = self.encode_one(t)
            if truncation and max_length is not None:
                ids = ids[:max_length]
            m = [1] * len(ids)
            if padding == "max_length":
                n = max_length - len(ids)
                if self.padding_side == "left":
                    ids, m = [self.pad_token_id] * n + ids, [0] * n + m
                else:
                    ids, m = ids + [self.pad_token_id] * n, m + [0] * n
            rows.append(ids)
            masks.append(m)
        return BatchEncoding({"input_ids": torch.tensor(rows, dtype=torch.long),
                              "attention_mask": torch.tensor(masks, dtype=torch.long)})
