"""Collator stage, CPU side: the numpy oracle against the golden vectors that the REFERENCE's own
collator / dataset / masking code produced (tests/golden/collate_golden.npz, oracle/make_golden_collate.py), against
Pillow itself where it is installed, and the kernels' shared host/device core (csrc/collate_core.h, compiled here with
g++ into a test harness) against the oracle. Integer / byte work: everything is compared bit-for-bit."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import collate_oracle as C  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "collate_golden.npz")


@pytest.fixture(scope="module")
def G():
    return np.load(GOLD)


def unpack_rows(G, prefix):
    """flat segment arrays -> the oracle's list-of-lists form."""
    tok, so = G[prefix + "tokens"], G[prefix + "seg_offsets"]
    sc, rel, fl, ro = G[prefix + "seg_score"], G[prefix + "seg_relation"], G[prefix + "seg_flags"], G[prefix + "row_offsets"]
    rows = []
    for i in range(len(ro) - 1):
        rows.append([(tok[so[j]:so[j + 1]].tolist(), float(sc[j]), float(rel[j]), int(fl[j])) for j in range(ro[i], ro[i + 1])])
    return rows


def test_oracle_detailed_report_vs_reference(G):
    Q, R = G["A_QR"]
    ids, mask, sc, rel, ov = C.collate_responses(unpack_rows(G, "A_orig_"), R, 0, 2, C.COLLATE_ADD_EOS)
    assert not ov.any()
    assert np.array_equal(ids, G["A_out_original_generate_response"])
    assert np.array_equal(mask, G["A_out_original_generate_response_attention_mask"])
    assert np.array_equal(sc, G["A_out_original_generate_response_scores"])
    assert np.array_equal(rel, G["A_out_original_generate_response_image_relations"])
    ids, mask, sc, rel, ov = C.collate_responses(unpack_rows(G, "A_ai_"), R, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
    assert np.array_equal(ids, G["A_out_AI_pseudo_response"])
    assert np.array_equal(mask, G["A_out_AI_pseudo_response_attention_mask"])
    assert np.array_equal(sc, G["A_out_AI_pseudo_response_scores"])
    assert np.array_equal(rel, G["A_out_AI_pseudo_response_image_relations"])
    # standard response and queries: plain tokenisation
    std = [[(t, 0.0, 0.0, C.FLAG_KEEP_BOS)] for t in json.loads(str(G["A_standard_tokens"]))]
    ids, *_ = C.collate_responses(std, R, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_TRUNCATE)
    assert np.array_equal(ids, G["A_out_standard_response"])
    qi, qm = C.collate_queries(json.loads(str(G["A_query_tokens"])), Q, 0)
    assert np.array_equal(qi, G["A_out_queries"]) and np.array_equal(qm, G["A_out_queries_attention_mask"].astype(bool))
    assert (qi == C.IMAGE_TOKEN_INDEX).sum(1).tolist() == [1, 1, 0]   # the third query is truncated before its image slot


def test_oracle_overflow_fallback_and_plain_branch_vs_reference(G):
    Q, R = G["A_QR"]
    plain = json.loads(str(G["B_plain_tokens"]))
    for k in ("standard_response", "original_generate_response", "AI_pseudo_response"):
        rows = [[(t, 0.0, 0.0, C.FLAG_KEEP_BOS)] for t in plain[k]]
        ids, _, sc, rel, _ = C.collate_responses(rows, R, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_TRUNCATE)
        assert np.array_equal(ids, G["B_out_" + k]), k
        assert not sc.any() and not rel.any()
    assert not G["B_out_AI_pseudo_response_scores"].any() and not G["B_out_original_generate_response_scores"].any()
    toks = json.loads(str(G["C_tokens"]))
    for k in ("standard_response", "original_generate_response", "AI_pseudo_response"):
        rows = [[(t, 0.0, 0.0, C.FLAG_KEEP_BOS)] for t in toks[k]]
        ids, mask, *_ = C.collate_responses(rows, R, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_TRUNCATE)
        assert np.array_equal(ids, G["C_out_" + k]), k
    assert (G["C_out_standard_response"][1] != 2).all(), "a truncated row carries no EOS"
    qi, qm = C.collate_queries(toks["queries"], Q, 0)
    assert np.array_equal(qi, G["C_out_queries"]) and np.array_equal(qm, G["C_out_queries_attention_mask"].astype(bool))
    # the overflow itself is reported, not silently truncated, when truncation is off
    long_row = [[(list(range(1, R + 10)), 1.0, 1.0, C.FLAG_KEEP_BOS)]]
    assert C.collate_responses(long_row, R, 0, 2, C.COLLATE_ADD_EOS)[4].tolist() == [True]


def test_oracle_pad_eos_wraps_when_eos_is_first():
    """pad_eos reads row_score[index - 1]: with the EOS at position 0 that is the LAST position (data_utils_dpo.py:94-96)."""
    ids, _, sc, rel, _ = C.collate_responses([[]], 8, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
    assert ids.tolist() == [[2, 0, 0, 0, 0, 0, 0, 0]] and sc[0, 0] == 1.0 and rel[0, 0] == 1.0
    full = [[(list(range(10, 18)), 2.5, 3.0, C.FLAG_KEEP_BOS)]]     # exactly R tokens: no pad, no EOS, nothing inherits
    ids, _, sc, _, _ = C.collate_responses(full, 8, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
    assert ids.tolist() == [list(range(10, 18))] and (sc == 2.5).all()


def test_oracle_image_preprocess_vs_reference(G):
    i = 0
    while f"D{i}_in" in G:
        S, pad = G[f"D{i}_cfg"]
        got = C.clip_preprocess(G[f"D{i}_in"], int(S), bool(pad))
        assert got.dtype == np.float32 and np.array_equal(got, G[f"D{i}_out"]), f"case {i}"
        i += 1
    assert i >= 6


def test_oracle_image_preprocess_full_size_hashes(G):
    for key, ref in json.loads(str(G["D_big"])).items():
        H, W, seed = map(int, key.split("x"))
        img = np.random.RandomState(seed).randint(0, 256, (H, W, 3)).astype(np.uint8)
        out = C.clip_preprocess(img, 336, True)
        assert hashlib.sha256(out.tobytes()).hexdigest() == ref["sha256"], key
        assert [float(v) for v in out[:, ::97, ::89].ravel()[:24]] == ref["probe"]


def test_oracle_resize_vs_pillow():
    Image = pytest.importorskip("PIL.Image")
    rs = np.random.RandomState(0)
    for H, W, oh, ow in [(480, 640, 336, 336), (200, 200, 336, 336), (333, 517, 336, 521), (37, 53, 56, 56), (700, 90, 64, 336)]:
        img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        ref = np.array(Image.fromarray(img).resize((ow, oh), resample=Image.BICUBIC))
        assert np.array_equal(C.pil_resize_bicubic(img, ow, oh), ref), (H, W, oh, ow)


def _mask_cases(G):
    for name in ("E0", "E1", "E2", "E3", "E4"):
        cfg = json.loads(str(G[name + "_cfg"]))
        yield name, cfg


def _mask_input(G, name, cfg):
    if name + "_in" in G:
        return G[name + "_in"]
    import torch
    g = torch.Generator().manual_seed(cfg["seed"])
    return torch.randn(tuple(cfg["shape"]), generator=g).numpy()


def test_oracle_copo_masks_vs_reference(G):
    for name, cfg in _mask_cases(G):
        img = _mask_input(G, name, cfg)
        # the fill value is a float mean: summation-order tolerance 2^-20 * mean|x| (bf16 images: the same bf16 value);
        # everything else — which pixels change, the untouched pixels — is bit-exact given the reference's fill value
        m_ref, m = float(G[name + "_mean"]), float(C.image_mean(img, cfg["bf16"]))
        assert abs(m - m_ref) <= (0.0 if cfg["bf16"] else 2.0 ** -20 * float(np.abs(img).mean())), (name, m, m_ref)
        out = C.mask_single_image(img, G[name + "_idx"], cfg["method"], bf16_pixels=cfg["bf16"], mean_value=m_ref)
        if name + "_out" in G:
            assert np.array_equal(out, G[name + "_out"]), name
        else:
            assert hashlib.sha256(out.tobytes()).hexdigest() == str(G[name + "_sha"]), name
    assert np.array_equal(C.attention_holes(576, G["E_attn_idx"]), G["E_attn_out"])
