"""Collator stage on the GPU: the kernels, called through the C ABI, against the golden vectors the
reference's own collator / dataset / masking code produced and against the CPU oracle on ragged random inputs.
Integer / byte work is compared bit-for-bit; the only floating-point quantity, the CoPO fill value (an image mean), is
allowed the float32 summation-order tolerance stated in the test."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLD = os.path.join(ROOT, "tests", "golden", "collate_golden.npz")


@pytest.fixture(scope="module")
def G():
    return np.load(GOLD)


def _pack_from(G, prefix):
    from opa_dpo_b200.collate import SegmentPack
    p = SegmentPack()
    p.tokens = G[prefix + "tokens"].tolist()
    p.seg_offsets = G[prefix + "seg_offsets"].tolist()
    p.score = G[prefix + "seg_score"].tolist()
    p.relation = G[prefix + "seg_relation"].tolist()
    p.flags = G[prefix + "seg_flags"].tolist()
    p.row_offsets = G[prefix + "row_offsets"].tolist()
    return p


def _rows_to_pack(rows):
    from opa_dpo_b200.collate import SegmentPack
    p = SegmentPack()
    for segs in rows:
        p.begin_row()
        for raw, ws, wr, f in segs:
            p.add(raw, ws, wr, f)
        p.end_row()
    return p


def test_collate_responses_kernel_vs_reference(G):
    from opa_dpo_b200 import collate as K
    Q, R = map(int, G["A_QR"])
    ids, mask, sc, rel, ov = K.collate_responses(_pack_from(G, "A_orig_"), R, 0, 2, K.ADD_EOS)
    torch.cuda.synchronize()
    assert not ov.any()
    assert np.array_equal(ids.cpu().numpy(), G["A_out_original_generate_response"])
    assert np.array_equal(mask.bool().cpu().numpy(), G["A_out_original_generate_response_attention_mask"])
    assert np.array_equal(sc.cpu().numpy(), G["A_out_original_generate_response_scores"])
    assert np.array_equal(rel.cpu().numpy(), G["A_out_original_generate_response_image_relations"])
    ids, mask, sc, rel, ov = K.collate_responses(_pack_from(G, "A_ai_"), R, 0, 2, K.ADD_EOS | K.PAD_EOS)
    assert np.array_equal(ids.cpu().numpy(), G["A_out_AI_pseudo_response"])
    assert np.array_equal(mask.bool().cpu().numpy(), G["A_out_AI_pseudo_response_attention_mask"])
    assert np.array_equal(sc.cpu().numpy(), G["A_out_AI_pseudo_response_scores"])
    assert np.array_equal(rel.cpu().numpy(), G["A_out_AI_pseudo_response_image_relations"])


def test_collate_responses_kernel_random_ragged_vs_oracle():
    """Empty rows, BOS-only entries, leading-29871 entries, pad / EOS ids inside the text, R from 1 to the repo's 896,
    every mode combination; overflow flags included."""
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    from _collate_util import random_rows
    rs = np.random.RandomState(3)
    for trial in range(24):
        R = int(rs.choice([1, 2, 7, 33, 128, 896]))
        rows = random_rows(rs, int(rs.randint(1, 18)), 8, 40 if R < 896 else 160, R)
        pack = _rows_to_pack(rows)
        for mode in (0, 1, 3, 5, 7, 8, 9, 13):
            want = C.collate_responses(rows, R, 0, 2, mode)
            got = K.collate_responses(pack, R, 0, 2, mode)
            for a, w, name in zip(got, want, ("ids", "mask", "scores", "relations", "overflow")):
                assert np.array_equal(a.cpu().numpy().astype(w.dtype), w), (trial, R, mode, name)


def test_collate_queries_kernel_vs_reference(G):
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    Q, R = map(int, G["A_QR"])
    ids, mask = K.collate_queries(json.loads(str(G["A_query_tokens"])), Q, 0)
    assert np.array_equal(ids.cpu().numpy(), G["A_out_queries"])
    assert np.array_equal(mask.cpu().numpy(), G["A_out_queries_attention_mask"].astype(np.uint8))
    toks = json.loads(str(G["C_tokens"]))["queries"] + [[]]
    ids, mask = K.collate_queries(toks, Q, 0)
    want = C.collate_queries(toks, Q, 0)
    assert np.array_equal(ids.cpu().numpy(), want[0]) and np.array_equal(mask.bool().cpu().numpy(), want[1])
    rs = np.random.RandomState(1)
    rows = [rs.randint(3, 32000, rs.randint(0, 300)).tolist() for _ in range(37)]
    for r in rows[::3]:
        if r:
            r[rs.randint(0, len(r))] = C.IMAGE_PLACEHOLDER_ID
    ids, mask = K.collate_queries(rows, 128, 0)
    want = C.collate_queries(rows, 128, 0)
    assert np.array_equal(ids.cpu().numpy(), want[0]) and np.array_equal(mask.bool().cpu().numpy(), want[1])


def test_data_collator_mirror_vs_reference_batches(G):
    """B200DataCollator end to end (host strings + kernels) on the instances the reference collator was run on:
    detailed report, the overflow fallback (zero int64 scores), and detailed_report off with truncation."""
    from oracle import collate_oracle as C
    from opa_dpo_b200.collate import B200DataCollator
    Q, R = map(int, G["A_QR"])
    dummy = torch.zeros(3, 4, 4)

    def run(instances_key, detailed):
        inst = json.loads(str(G[instances_key]))
        for i in inst:
            i["images"] = dummy
        return B200DataCollator(C.StubTokenizer(), Q, R, detailed_report=detailed)(inst)

    batch = run("A_instances", True)
    for k in ("queries", "queries_attention_mask", "standard_response", "original_generate_response",
              "original_generate_response_attention_mask", "AI_pseudo_response", "AI_pseudo_response_attention_mask",
              "original_generate_response_scores", "AI_pseudo_response_scores",
              "original_generate_response_image_relations", "AI_pseudo_response_image_relations"):
        got, want = batch[k].cpu().numpy(), G["A_out_" + k]
        assert got.dtype == want.dtype and np.array_equal(got, want), k
    assert batch["images"].shape == (3, 3, 4, 4)
    batch = run("B_instances", True)            # one report overflows R: whole batch falls back
    for k in ("original_generate_response", "AI_pseudo_response", "original_generate_response_scores",
              "AI_pseudo_response_scores", "AI_pseudo_response_image_relations", "standard_response"):
        got, want = batch[k].cpu().numpy(), G["B_out_" + k]
        assert got.dtype == want.dtype and np.array_equal(got, want), k
    inst = json.loads(str(G["A_instances"]))[:2]   # C instances are not stored; rebuild the plain branch from its tokens
    toks = json.loads(str(G["C_tokens"]))

    class FixedTok(C.StubTokenizer):
        def __init__(self, table):
            super().__init__()
            self.table = table

        def encode_one(self, text):
            return self.table[text]

    table = {}
    for j in range(2):
        for k in ("queries", "standard_response", "original_generate_response", "AI_pseudo_response"):
            inst[j][k] = f"{k}#{j}"
            table[f"{k}#{j}"] = toks[k][j]
        inst[j]["images"] = dummy
    batch = B200DataCollator(FixedTok(table), Q, R, detailed_report=False)(inst)
    for k in ("queries", "queries_attention_mask", "standard_response", "standard_response_attention_mask",
              "original_generate_response", "AI_pseudo_response"):
        got, want = batch[k].cpu().numpy(), G["C_out_" + k]
        assert got.dtype == want.dtype and np.array_equal(got, want), k
    assert "AI_pseudo_response_scores" not in batch


def test_preprocess_images_kernel_vs_reference(G):
    from opa_dpo_b200 import collate as K
    i = 0
    while f"D{i}_in" in G:
        S, pad = map(int, G[f"D{i}_cfg"])
        got = K.preprocess_images([G[f"D{i}_in"]], S, bool(pad))[0].cpu().numpy()
        assert np.array_equal(got, G[f"D{i}_out"]), f"case {i}"
        i += 1
    big = json.loads(str(G["D_big"]))
    imgs, keys = [], []
    for key in big:
        H, W, seed = map(int, key.split("x"))
        imgs.append(np.random.RandomState(seed).randint(0, 256, (H, W, 3)).astype(np.uint8))
        keys.append(key)
    out = K.preprocess_images(imgs, 336, True).cpu().numpy()       # one ragged batch at the real 336 size
    for j, key in enumerate(keys):
        assert hashlib.sha256(out[j].tobytes()).hexdigest() == big[key]["sha256"], key


def test_preprocess_images_kernel_ragged_batch_vs_oracle():
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    rs = np.random.RandomState(11)
    shapes = [(1, 1), (2, 9), (9, 2), (57, 31), (31, 57), (64, 64), (130, 77), (400, 123), (700, 1100)]
    imgs = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in shapes]
    for S, pad in ((32, True), (32, False), (336, True), (224, False)):
        got = K.preprocess_images(imgs, S, pad).cpu().numpy()
        for j, im in enumerate(imgs):
            assert np.array_equal(got[j], C.clip_preprocess(im, S, pad)), (shapes[j], S, pad)


def test_copo_masks_kernel_vs_reference(G):
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    from _collate_util import _mask_cases, _mask_input
    for name, cfg in _mask_cases(G):
        img = torch.from_numpy(_mask_input(G, name, cfg)).cuda()
        x = img.to(torch.bfloat16) if cfg["bf16"] else img
        idx = torch.from_numpy(G[name + "_idx"])[None]
        out, means = K.mask_images(x[None], cfg["ratio"], cfg["method"], indices=idx, return_means=True)
        m_ref, m = float(G[name + "_mean"]), float(means[0])
        # the fill value is a float mean: summation-order tolerance 2^-20 * mean|x| (bf16 images: the same bf16 value)
        assert abs(m - m_ref) <= (0.0 if cfg["bf16"] else 2.0 ** -20 * float(img.abs().mean())), (name, m, m_ref)
        want = C.mask_single_image(img.cpu().numpy(), G[name + "_idx"], cfg["method"], bf16_pixels=cfg["bf16"], mean_value=m)
        assert out.dtype == x.dtype and np.array_equal(out[0].float().cpu().numpy(), want), name
        if name + "_out" in G and m == m_ref:
            assert np.array_equal(out[0].float().cpu().numpy(), G[name + "_out"]), name
    holes = K.mask_percentage_per_row(3, 576, 0.3, indices=torch.from_numpy(G["E_attn_idx"]))
    assert np.array_equal(holes.cpu().numpy(), G["E_attn_out"])
    # seeded run draws the reference's permutations: same pixels as mask_single_image under the same seed
    torch.manual_seed(99)
    cfg = json.loads(str(G["E0_cfg"]))
    img = torch.from_numpy(_mask_input(G, "E0", cfg)).cuda()
    out = K.mask_images(img[None], cfg["ratio"], cfg["method"])
    changed = (out[0] != img).any(0).flatten().nonzero()[:, 0].cpu().numpy()
    assert set(changed.tolist()) <= set(G["E0_idx"].tolist()) and len(changed) >= len(G["E0_idx"]) - 2


def test_collator_full_size_properties():
    """BASELINE-size batch (16 rows x R=896, Q=128, 16 images 480x640 -> 336): size-independent properties — every row
    ends in exactly one inserted EOS after its tokens, pads after it, masks consistent, weights zero at pads, image
    values inside the normalised [0,255] range and idempotent on an already 336x336 image."""
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    from _collate_util import random_rows
    rs = np.random.RandomState(5)
    R = 896
    rows = []
    for i in range(16):
        segs = []
        for j in range(12):
            raw = [1] + rs.randint(3, 32000, rs.randint(1, 60)).tolist()
            segs.append((raw, 2.0, 3.0, K.SEG_KEEP_BOS if j == 0 else 0))
        rows.append(segs)
    ids, mask, sc, rel, ov = K.collate_responses(_rows_to_pack(rows), R, 0, 2, K.ADD_EOS | K.PAD_EOS)
    ids, mask, sc, ov = ids.cpu(), mask.cpu(), sc.cpu(), ov.cpu()
    assert not ov.any()
    for i in range(16):
        n = sum(len(r[0]) - (0 if j == 0 else 1) for j, r in enumerate(rows[i]))
        assert ids[i, n] == 2 and (ids[i, n + 1:] == 0).all() and (ids[i, :n] != 0).all()
        assert mask[i, :n + 1].all() and not mask[i, n + 1:].any()
        assert (sc[i, :n + 1] == 2.0).all() and (sc[i, n + 1:] == 0).all()
    imgs = [rs.randint(0, 256, (480, 640, 3)).astype(np.uint8) for _ in range(16)]
    out = K.preprocess_images(imgs, 336, True)
    lo = (0.0 - np.asarray(C.CLIP_MEAN)) / np.asarray(C.CLIP_STD)
    hi = (1.0 - np.asarray(C.CLIP_MEAN)) / np.asarray(C.CLIP_STD)
    for c in range(3):
        assert out[:, c].min().item() >= lo[c] - 1e-5 and out[:, c].max().item() <= hi[c] + 1e-5
    assert np.array_equal(out[3].cpu().numpy(), C.clip_preprocess(imgs[3], 336, True))
    sq = rs.randint(0, 256, (336, 336, 3)).astype(np.uint8)
    same = K.preprocess_images([sq], 336, True)[0].cpu().numpy()
    plain = ((sq.astype(np.float64) * (1 / 255)).astype(np.float32) - np.asarray(C.CLIP_MEAN, np.float32)) / np.asarray(C.CLIP_STD, np.float32)
    assert np.array_equal(same, plain.transpose(2, 0, 1))


def test_preprocess_images_generic_kernels_vs_oracle():
    """The fallback kernels: a canvas row wider than shared memory (W*3 > 48 KB -> per-tap bounds-checked horizontal pass)
    and an output size whose row is not a multiple of 4 bytes (-> scalar vertical pass)."""
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    rs = np.random.RandomState(13)
    wide = rs.randint(0, 256, (3, 16400, 3)).astype(np.uint8)
    small = rs.randint(0, 256, (45, 61, 3)).astype(np.uint8)
    for S in (30, 32):
        got = K.preprocess_images([wide, small], S, False).cpu().numpy()
        assert np.array_equal(got[0], C.clip_preprocess(wide, S, False)), ("wide", S)
        assert np.array_equal(got[1], C.clip_preprocess(small, S, False)), ("small", S)
    got = K.preprocess_images([small], 30, True).cpu().numpy()
    assert np.array_equal(got[0], C.clip_preprocess(small, 30, True))


def test_collate_responses_many_segments_fallback_vs_oracle():
    """Rows with more entries than the kernel indexes in shared memory (512) take the per-position walk; rows with many
    empty / BOS-only entries exercise the boundary handling of the binary search."""
    from oracle import collate_oracle as C
    from opa_dpo_b200 import collate as K
    rs = np.random.RandomState(21)
    rows = []
    for nseg in (600, 512, 513, 40):
        segs = []
        for j in range(nseg):
            n = int(rs.choice([0, 0, 1, 2, 3]))
            raw = [1] + rs.randint(3, 32000, n).tolist()
            if n and rs.rand() < 0.3:
                raw[1] = C.SPACE_TOKEN
            segs.append((raw, float(rs.choice([1.0, 2.5])), float(rs.choice([1.0, 3.0])),
                         (C.FLAG_KEEP_BOS if j == 0 else 0) | (C.FLAG_STRIP_SPACE if rs.rand() < 0.5 else 0)))
        rows.append(segs)
    pack = _rows_to_pack(rows)
    for R in (64, 896, 2000):
        for mode in (3, 9):
            want = C.collate_responses(rows, R, 0, 2, mode)
            got = K.collate_responses(pack, R, 0, 2, mode)
            for a, w, name in zip(got, want, ("ids", "mask", "scores", "relations", "overflow")):
                assert np.array_equal(a.cpu().numpy().astype(w.dtype), w), (R, mode, name)


def test_truncate_after_eos_kernel_vs_reference():
    """Rollout post-processing on the device against generator.truncate_after_eos_with_padding's own outputs (golden) and
    the oracle on random rows at the repo's R = 896."""
    from oracle import llava_oracle as O
    from opa_dpo_b200.generation import truncate_after_eos_with_padding as trunc
    gold = np.load(os.path.join(ROOT, "tests", "golden", "opadpo_tiny.npz"))
    comp = torch.from_numpy(gold["trunc.in"]).long()
    assert np.array_equal(trunc(comp.cuda(), 2, 0, additional_tokens=[1577, 29973]).cpu().numpy(), gold["trunc.out"])
    assert np.array_equal(trunc(comp.cuda(), 2, 0).cpu().numpy(), gold["trunc.out_noextra"])
    rs = np.random.RandomState(2)
    for R in (1, 5, 64, 896):
        x = torch.from_numpy(rs.choice([2, 7, 8, 9, 1577, 29973, 11], size=(33, R), p=[.01, .33, .33, .3, .01, .01, .01])).long()
        for extra in (None, [1577, 29973], [29973, 1577], [11, 11, 8]):
            want = O.truncate_after_eos_with_padding(x, 2, 0, extra)
            got = trunc(x.cuda(), 2, 0, extra)
            assert got.dtype == torch.int64 and torch.equal(got.cpu(), want), (R, extra)
    quirk = torch.tensor([[5, 2, 6, 1577, 7, 7]]).cuda()
    assert trunc(quirk, 2, 0, [1577, 29973]).tolist() == [[5, 2, 6, 1577, 0, 0]]
