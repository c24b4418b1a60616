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
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import collate_oracle as C  # noqa: E402
from _collate_util import _mask_cases, _mask_input, random_rows, unpack_rows  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "collate_golden.npz")


@pytest.fixture(scope="module")
def G():
    return np.load(GOLD)


def test_oracle_detailed_report_vs_reference(G):
    Q, R = G["A_QR"]
    ids, mask, sc, rel, ov = C.collate_responses(unpack_rows(G, "A_orig_"), R, 0, 2, C.COLLATE_ADD_EOS)
    assert not ov.any()
    assert np.array_equal(ids, G["A_out_original_generate_response"])
    assert np.array_equal(mask.astype(bool), G["A_out_original_generate_response_attention_mask"])
    assert np.array_equal(sc, G["A_out_original_generate_response_scores"])
    assert np.array_equal(rel, G["A_out_original_generate_response_image_relations"])
    ids, mask, sc, rel, ov = C.collate_responses(unpack_rows(G, "A_ai_"), R, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
    assert np.array_equal(ids, G["A_out_AI_pseudo_response"])
    assert np.array_equal(mask.astype(bool), G["A_out_AI_pseudo_response_attention_mask"])
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


# ------------------------------------------------------------------------------------------------ kernel core on the host
@pytest.fixture(scope="module")
def harness():
    """g++ build of tests/host_harness/collate_host.cpp: the kernels' per-element core, compiled for the CPU."""
    import ctypes
    out_dir = os.path.join(ROOT, "build", "host_harness")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcollate_host.so")
    src = os.path.join(ROOT, "tests", "host_harness", "collate_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _p(a):
    import ctypes
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _harness_responses(lib, G, prefix, R, mode, with_weights=True):
    arrs = {k: np.ascontiguousarray(G[prefix + k]) for k in ("tokens", "seg_offsets", "seg_score", "seg_relation", "seg_flags", "row_offsets")}
    b = len(arrs["row_offsets"]) - 1
    ids = np.full((b, R), -7, np.int64)
    mask = np.full((b, R), 9, np.uint8)
    sc = np.full((b, R), np.nan, np.float32) if with_weights else None
    rel = np.full((b, R), np.nan, np.float32) if with_weights else None
    ov = np.full((b,), -1, np.int32)
    lib.host_collate_responses(_p(arrs["tokens"]), _p(arrs["seg_offsets"]), _p(arrs["seg_score"]), _p(arrs["seg_relation"]),
                               _p(arrs["seg_flags"]), _p(arrs["row_offsets"]), b, int(R), 0, 2, mode, _p(ids), _p(mask), _p(sc),
                               _p(rel), _p(ov))
    return ids, mask, sc, rel, ov


def test_kernel_core_responses_vs_reference(harness, G):
    Q, R = G["A_QR"]
    ids, mask, sc, rel, ov = _harness_responses(harness, G, "A_orig_", R, C.COLLATE_ADD_EOS)
    assert not ov.any()
    assert np.array_equal(ids, G["A_out_original_generate_response"])
    assert np.array_equal(mask.astype(bool), G["A_out_original_generate_response_attention_mask"])
    assert np.array_equal(sc, G["A_out_original_generate_response_scores"])
    assert np.array_equal(rel, G["A_out_original_generate_response_image_relations"])
    ids, mask, sc, rel, ov = _harness_responses(harness, G, "A_ai_", R, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
    assert np.array_equal(ids, G["A_out_AI_pseudo_response"])
    assert np.array_equal(sc, G["A_out_AI_pseudo_response_scores"])
    assert np.array_equal(rel, G["A_out_AI_pseudo_response_image_relations"])
    # a tighter R: the same rows overflow, flagged and truncated exactly like the oracle
    rows = unpack_rows(G, "A_ai_")
    for R2 in (5, 17, 30):
        want = C.collate_responses(rows, R2, 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
        got = _harness_responses(harness, G, "A_ai_", R2, C.COLLATE_ADD_EOS | C.COLLATE_PAD_EOS)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3]) and np.array_equal(got[4].astype(bool), want[4])


def _pack(rows):
    toks, so, sc, rel, fl, ro = [], [0], [], [], [], [0]
    for segs in rows:
        for raw, ws, wr, f in segs:
            toks.extend(raw)
            so.append(len(toks))
            sc.append(ws)
            rel.append(wr)
            fl.append(f)
        ro.append(len(sc))
    return {"tokens": np.asarray(toks + [0], np.int32), "seg_offsets": np.asarray(so, np.int64), "seg_score": np.asarray(sc + [0], np.float32),
            "seg_relation": np.asarray(rel + [0], np.float32), "seg_flags": np.asarray(fl + [0], np.uint8), "row_offsets": np.asarray(ro, np.int64)}


def test_kernel_core_responses_random_ragged_vs_oracle(harness):
    rs = np.random.RandomState(3)
    for trial in range(40):
        R = int(rs.choice([1, 2, 7, 33, 128]))
        rows = random_rows(rs, int(rs.randint(1, 6)), 6, 12, R)
        for mode in (0, 1, 3, 5, 7, 8, 9, 13):
            want = C.collate_responses(rows, R, 0, 2, mode)
            got = _harness_responses(harness, _pack(rows), "", R, mode)
            for a, w, name in zip(got, want, ("ids", "mask", "scores", "relations", "overflow")):
                assert np.array_equal(np.asarray(a).astype(w.dtype), w), (trial, mode, name)


def test_kernel_core_queries_vs_reference(harness, G):
    Q, R = G["A_QR"]
    for key, want_ids, want_mask in (("A_query_tokens", "A_out_queries", "A_out_queries_attention_mask"),):
        toks = json.loads(str(G[key]))
        flat = np.asarray([t for r in toks for t in r], np.int32)
        off = np.asarray(np.cumsum([0] + [len(r) for r in toks]), np.int64)
        ids = np.zeros((len(toks), Q), np.int64)
        mask = np.zeros((len(toks), Q), np.uint8)
        harness.host_collate_queries(_p(flat), _p(off), len(toks), int(Q), 0, C.IMAGE_PLACEHOLDER_ID, C.IMAGE_TOKEN_INDEX, _p(ids), _p(mask))
        assert np.array_equal(ids, G[want_ids]) and np.array_equal(mask, G[want_mask].astype(np.uint8))
    toks = json.loads(str(G["C_tokens"]))["queries"] + [[]]          # + an empty query: all pad
    flat = np.asarray([t for r in toks for t in r], np.int32)
    off = np.asarray(np.cumsum([0] + [len(r) for r in toks]), np.int64)
    ids = np.zeros((len(toks), Q), np.int64)
    mask = np.zeros((len(toks), Q), np.uint8)
    harness.host_collate_queries(_p(flat), _p(off), len(toks), int(Q), 0, C.IMAGE_PLACEHOLDER_ID, C.IMAGE_TOKEN_INDEX, _p(ids), _p(mask))
    want = C.collate_queries(toks, Q, 0)
    assert np.array_equal(ids, want[0]) and np.array_equal(mask.astype(bool), want[1])
    assert np.array_equal(ids[:2], G["C_out_queries"])


def _harness_preprocess(lib, imgs, S, pad, fast=1):
    import ctypes
    flat = np.concatenate([i.reshape(-1) for i in imgs]).astype(np.uint8)
    off = np.asarray(np.cumsum([0] + [i.size for i in imgs])[:-1], np.int64)
    hs = np.asarray([i.shape[0] for i in imgs], np.int32)
    ws = np.asarray([i.shape[1] for i in imgs], np.int32)
    mean, std = np.asarray(C.CLIP_MEAN, np.float32), np.asarray(C.CLIP_STD, np.float32)
    bg = np.asarray([int(x * 255) for x in C.CLIP_MEAN], np.uint8)
    out = np.full((len(imgs), 3, S, S), np.nan, np.float32)
    lib.host_preprocess_images(_p(flat), _p(off), _p(hs), _p(ws), len(imgs), S, _p(mean), _p(std), _p(bg), int(pad), int(fast), _p(out))
    return out


def test_kernel_core_image_preprocess_vs_reference(harness, G):
    i = 0
    while f"D{i}_in" in G:
        S, pad = G[f"D{i}_cfg"]
        for fast in (0, 1, 2):
            got = _harness_preprocess(harness, [G[f"D{i}_in"]], int(S), bool(pad), fast)[0]
            assert np.array_equal(got, G[f"D{i}_out"]), f"case {i} fast={fast}"
        i += 1
    for key, ref in json.loads(str(G["D_big"])).items():
        H, W, seed = map(int, key.split("x"))
        img = np.random.RandomState(seed).randint(0, 256, (H, W, 3)).astype(np.uint8)
        for fast in (0, 1):
            out = _harness_preprocess(harness, [img], 336, True, fast)[0]
            assert hashlib.sha256(out.tobytes()).hexdigest() == ref["sha256"], (key, fast)


def test_kernel_core_image_preprocess_ragged_batch_vs_oracle(harness):
    rs = np.random.RandomState(11)
    shapes = [(1, 1), (2, 9), (9, 2), (57, 31), (31, 57), (64, 64), (130, 77), (400, 123)]
    imgs = [rs.randint(0, 256, (h, w, 3)).astype(np.uint8) for h, w in shapes]
    for S, pad, fast in ((32, True, 1), (32, False, 1), (48, True, 0), (30, True, 1), (31, False, 1), (30, False, 0), (32, True, 2)):
        got = _harness_preprocess(harness, imgs, S, pad, fast)     # S = 30, 31: 3*S not a multiple of 4 -> no quad path
        for j, im in enumerate(imgs):
            assert np.array_equal(got[j], C.clip_preprocess(im, S, pad)), (shapes[j], S, pad, fast)


def test_kernel_core_bf16_rounding(harness):
    import ctypes
    harness.host_round_bf16.restype = ctypes.c_float
    harness.host_round_bf16.argtypes = [ctypes.c_float]
    import torch
    x = torch.randn(2000) * 3
    want = x.to(torch.bfloat16).float()
    got = torch.tensor([harness.host_round_bf16(float(v)) for v in x])
    assert torch.equal(got, want)
    assert np.array_equal(C._round_bf16(x.numpy()), want.numpy())


# ------------------------------------------------------------------------------------------------ host mirror (strings)
def test_host_mirror_realign_copied_vs_reference(G):
    from opa_dpo_b200.collate import realign_copied
    for original, pieces, want in json.loads(str(G["F_realign"])):
        assert realign_copied(original, list(pieces)) == want, (original, pieces)


def test_host_mirror_report_segments_vs_reference_tokenizer_log(G):
    """B200DataCollator's host half (JSON report -> flat segment arrays) against the arrays rebuilt from the tokenizer
    calls the REFERENCE collator made on the same instances."""
    from opa_dpo_b200.collate import report_segments, plain_pack
    tok = C.StubTokenizer()
    inst = json.loads(str(G["A_instances"]))
    reports = [json.loads(i["AI_json_report"]) for i in inst]
    orig, ai = report_segments(tok.encode_one, reports, [i["original_generate_response"] for i in inst])
    for name, pack in (("orig", orig), ("ai", ai)):
        for k, v in pack.arrays().items():
            assert np.array_equal(v, G[f"A_{name}_{k}"]), (name, k)
    assert min(orig.row_present + ai.row_present) > 0
    p = plain_pack([tok.encode_one(i["standard_response"]) for i in inst]).arrays()
    rows = [[(p["tokens"][p["seg_offsets"][j]:p["seg_offsets"][j + 1]].tolist(), 0.0, 0.0, int(p["seg_flags"][j]))] for j in range(len(inst))]
    ids, mask, *_ = C.collate_responses(rows, int(G["A_QR"][1]), 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_TRUNCATE | C.COLLATE_TOKEN_MASK)
    assert np.array_equal(ids, G["A_out_standard_response"])


def test_oracle_tokenizer_mask_mode_vs_reference(G):
    """Plain branch: the batch carries the TOKENIZER's attention mask (the EOS written over the first pad stays 0)."""
    toks = json.loads(str(G["C_tokens"]))
    rows = [[(t, 0.0, 0.0, C.FLAG_KEEP_BOS)] for t in toks["standard_response"]]
    _, mask, *_ = C.collate_responses(rows, int(G["A_QR"][1]), 0, 2, C.COLLATE_ADD_EOS | C.COLLATE_TRUNCATE | C.COLLATE_TOKEN_MASK)
    assert np.array_equal(mask, G["C_out_standard_response_attention_mask"])
    assert mask[0].max() == 2, "add_eos also overwrites the attention mask's first 0 with eos_id"


def test_collator_product_has_no_cpu_path():
    """Without a CUDA device the collator ops raise; they never route through the oracle or any CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from opa_dpo_b200 import collate as K
    from opa_dpo_b200._lib import OpadpoError
    with pytest.raises(OpadpoError):
        K.collate_queries([[1, 2, 3]], 8, 0)
    with pytest.raises(OpadpoError):
        K.preprocess_images([np.zeros((4, 4, 3), np.uint8)], 8)
    import inspect
    src = inspect.getsource(K)
    assert "oracle" not in src.replace("the oracle", "")


# ------------------------------------------------------------------------------------------------ rollout post-processing
def _trunc_ref(x, eos, pad, extra):
    from oracle import llava_oracle as O
    import torch
    return O.truncate_after_eos_with_padding(torch.from_numpy(x), eos, pad, extra).numpy()


def test_kernel_core_truncate_after_eos_vs_reference(harness):
    """The truncate kernel's decision logic (csrc/collate_core.h) against the reference's own outputs
    (tests/golden/opadpo_tiny.npz: generator.truncate_after_eos_with_padding run verbatim) and the oracle on random rows,
    incl. the override quirk: a later stop token wins even when it lies behind the EOS."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "opadpo_tiny.npz"))
    comp = np.ascontiguousarray(gold["trunc.in"]).astype(np.int64)

    def run(x, extra):
        out = np.full_like(x, -9)
        ex = np.asarray(extra or [0], np.int64)
        harness.host_truncate_after_eos(_p(x), _p(out), x.shape[0], x.shape[1], 2, 0, _p(ex), len(extra or []))
        return out
    assert np.array_equal(run(comp, [1577, 29973]), gold["trunc.out"])
    assert np.array_equal(run(comp, None), gold["trunc.out_noextra"])
    rs = np.random.RandomState(2)
    for trial in range(50):
        R = int(rs.choice([1, 5, 64, 300]))
        x = rs.choice([2, 7, 8, 9, 1577, 29973, 11], size=(int(rs.randint(1, 6)), R), p=[.04, .3, .3, .26, .03, .03, .04]).astype(np.int64)
        for extra in (None, [1577, 29973], [29973, 1577], [11, 11, 8]):
            assert np.array_equal(run(x, extra), _trunc_ref(x, 2, 0, extra)), (trial, extra)
    quirk = np.asarray([[5, 2, 6, 1577, 7, 7]], np.int64)            # EOS at 1, '?' at 3: the cut moves BEHIND the EOS
    assert run(quirk, [1577, 29973]).tolist() == [[5, 2, 6, 1577, 0, 0]]


def test_host_mirror_rollout_writer_vs_reference(G, tmp_path):
    """record_step_stats: byte-identical file to the reference's (online_generator.py:380-396, run verbatim for the golden)."""
    import base64
    from opa_dpo_b200.generation import record_step_stats
    rd = json.loads(str(G["G_rollout_input"]))
    rd["image_bytes"] = [base64.b64decode(x) for x in rd["image_bytes"]]
    path = record_step_stats(str(tmp_path), 7, rd)
    assert os.path.basename(path) == f"step7_rank{int(os.environ.get('RANK', 0))}.json" and os.path.basename(os.path.dirname(path)) == "rollouts"
    assert open(path).read() == str(G["G_rollout_json"])
    assert record_step_stats(None, 7, rd) is None


def test_host_mirror_realign_copied_property_vs_reference_function():
    """Where /root/reference is present (this container; not the GPU box) the host mirror of complete_copied_content is
    checked against the reference function itself on generated inputs (hypothesis)."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st
    from oracle.make_golden import install_stubs
    install_stubs()
    from utils import data_utils_dpo as D
    from opa_dpo_b200.collate import realign_copied
    alphabet = st.sampled_from(list("ab .,\n"))
    text = st.text(alphabet, max_size=24)

    @settings(max_examples=400, deadline=None)
    @given(original=text, cuts=st.lists(st.integers(0, 24), max_size=5), noise=st.lists(text, max_size=2), pad=st.booleans())
    def check(original, cuts, noise, pad):
        # pieces that are mostly substrings of the original (as GPT-4V's 'copied content' is), sometimes padded / foreign
        cuts = sorted(min(c, len(original)) for c in cuts)
        pieces = [original[a:b] for a, b in zip([0] + cuts, cuts + [len(original)])]
        if pad:
            pieces = [" " + p + " " for p in pieces]
        pieces += noise
        assert realign_copied(original, list(pieces)) == D.complete_copied_content(original, list(pieces))
    check()


# ------------------------------------------------------------------------------------------------ host mirror, whole __call__
class OracleOps:
    """The collator's array operations answered by the CPU oracle (torch CPU tensors), so that B200DataCollator.__call__ —
    report parsing, fallback rules, batch keys and dtypes — can run without a GPU. Test infrastructure: the product's
    only implementation of these operations are the CUDA kernels."""

    @staticmethod
    def _rows(pack):
        a = pack.arrays()
        return [[(a["tokens"][a["seg_offsets"][j]:a["seg_offsets"][j + 1]].tolist(), float(a["seg_score"][j]), float(a["seg_relation"][j]),
                  int(a["seg_flags"][j])) for j in range(a["row_offsets"][i], a["row_offsets"][i + 1])] for i in range(len(a["row_offsets"]) - 1)]

    def collate_responses(self, pack, R, pad_id, eos_id, mode, device=None, weights=True):
        import torch
        ids, mask, sc, rel, ov = C.collate_responses(self._rows(pack), R, pad_id, eos_id, mode)
        t = torch.from_numpy
        return t(ids), t(mask), (t(sc) if weights else None), (t(rel) if weights else None), t(ov.astype(np.int32))

    def collate_queries(self, rows, Q, pad_id, device=None):
        import torch
        ids, mask = C.collate_queries(rows, Q, pad_id)
        return torch.from_numpy(ids), torch.from_numpy(mask.astype(np.uint8))

    def preprocess_images(self, images, size=336, pad_to_square=True, device=None):
        import torch
        return torch.from_numpy(np.stack([C.clip_preprocess(np.asarray(i), size, pad_to_square) for i in images]))


def _same_batch(got, want, keys=None):
    import torch
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in (keys or want):
        g, w = got[k], want[k]
        assert isinstance(g, torch.Tensor) and g.dtype == w.dtype and torch.equal(g, w), k


def test_host_mirror_whole_collator_vs_reference_goldens(G):
    """B200DataCollator.__call__ (host half) over the oracle ops == the reference collator's batches: detailed report, the
    overflow fallback, plain branch."""
    import torch
    from opa_dpo_b200.collate import B200DataCollator
    Q, R = map(int, G["A_QR"])
    for key, detailed, prefix in (("A_instances", True, "A_out_"), ("B_instances", True, "B_out_")):
        inst = json.loads(str(G[key]))
        for i in inst:
            i["images"] = torch.zeros(3, 4, 4)
        batch = B200DataCollator(C.StubTokenizer(), Q, R, detailed_report=detailed, ops=OracleOps())(inst)
        names = [k[len(prefix):] for k in G.files if k.startswith(prefix)]
        assert names
        for k in names:
            assert batch[k].dtype == torch.from_numpy(G[prefix + k]).dtype and np.array_equal(batch[k].numpy(), G[prefix + k]), (key, k)


def test_host_mirror_whole_collator_property_vs_reference_collator():
    """Generated reports (hypothesis) through the reference's DataCollatorForCausalLM.__call__ run verbatim and through
    B200DataCollator.__call__ over the oracle ops: same keys, dtypes and values — including every way a report makes the
    reference fall back (malformed entries, a stream without entries, a row longer than response_len)."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    pytest.importorskip("hypothesis")
    import contextlib
    import io
    import torch
    from hypothesis import given, settings, strategies as st
    from oracle.make_golden import install_stubs
    install_stubs()
    from utils import data_utils_dpo as D
    from opa_dpo_b200.collate import B200DataCollator
    word = st.sampled_from(["a", "b", "cc", "SPACE", "d."])
    phrase = st.lists(word, max_size=6).map(" ".join)
    entry = st.one_of(
        st.fixed_dictionaries({}, optional={
            "copied content": phrase, "copied_content": phrase, "rewritten content": phrase, "rewritten_content": phrase,
            "score": st.sampled_from([1, 2, 3, 4, 5, "4"]), "error type": st.sampled_from(["correct", "image_recognition_error", "other"]),
            "error_type": st.sampled_from(["language_comprehension_error", "correct"])}),
        st.just("not a dict"))
    report = st.dictionaries(st.sampled_from(["Sentence 1", "Sentence 2", "Sentence 3", "Added", "image_description"]), entry, max_size=5)
    instance = st.fixed_dictionaries({"queries": phrase.map(lambda p: "SYS " + p + " IMG q"), "standard_response": phrase,
                                      "original_generate_response": phrase, "AI_pseudo_response": phrase, "report": report})

    @settings(max_examples=500, deadline=None)
    @given(insts=st.lists(instance, min_size=1, max_size=3), R=st.sampled_from([6, 12, 40]), detailed=st.booleans())
    def check(insts, R, detailed):
        batch_in = [{**{k: v for k, v in i.items() if k != "report"}, "AI_json_report": json.dumps(i["report"]),
                     "images": torch.zeros(3, 2, 2)} for i in insts]
        with contextlib.redirect_stdout(io.StringIO()):          # both print the exception when they fall back
            want = D.DataCollatorForCausalLM(tokenizer=C.StubTokenizer(), query_len=8, response_len=R, detailed_report=detailed)(
                [dict(i) for i in batch_in])
            got = B200DataCollator(C.StubTokenizer(), 8, R, detailed_report=detailed, ops=OracleOps())([dict(i) for i in batch_in])
        _same_batch(got, want)
    check()


def test_host_mirror_dataset_vs_reference_dataset(tmp_path, monkeypatch):
    """B200DPODataset.__getitem__ against the reference's DPO_Dataset.__getitem__ run verbatim: identical text fields
    (query template, the placeholder glyph for <image>), and the decoded uint8 image it hands to the GPU collator turns —
    through the oracle's expand2square + CLIP preprocess, which the kernels are tested against — into exactly the tensor
    the reference's processor returns. All three image sources (file via 'images' / 'image_id', base64 'image_bytes')."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    import base64
    import io
    from types import SimpleNamespace
    from PIL import Image
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
    from oracle.make_golden import install_stubs
    install_stubs()
    from utils import data_utils_dpo as D
    from opa_dpo_b200.collate import B200DPODataset
    rs = np.random.RandomState(0)
    imgs = [rs.randint(0, 256, s).astype(np.uint8) for s in ((40, 70, 3), (91, 64, 3), (56, 56, 3))]
    Image.fromarray(imgs[0]).save(tmp_path / "a.png")
    Image.fromarray(imgs[1]).save(tmp_path / "b.png")
    buf = io.BytesIO()
    Image.fromarray(imgs[2]).save(buf, format="PNG")
    text = {"standard_response": "s", "original_generate_response": "o", "AI_pseudo_response": "p", "AI_json_report": "{}"}
    items = [{"images": "a.png", "queries": "<image>\nwhat is this?", **text},
             {"image_id": "b.png", "queries": "describe <image> now", **text},
             {"image_bytes": base64.b64encode(buf.getvalue()).decode(), "queries": "no placeholder", **text}]
    monkeypatch.setenv("IMAGE_DIR", str(tmp_path))
    proc = CLIPImageProcessorPil(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56})
    ref = D.DPO_Dataset(SimpleNamespace(image_processor=proc, image_aspect_ratio="pad"), items)
    mine = B200DPODataset(items)
    assert len(mine) == len(ref) == 3
    for i in range(3):
        r, m = ref[i], mine[i]
        assert set(r) == set(m)
        for k in r:
            if k != "images":
                assert r[k] == m[k], k
        assert m["images"].dtype == np.uint8 and np.array_equal(m["images"], imgs[i])
        want = r["images"].numpy() if hasattr(r["images"], "numpy") else np.asarray(r["images"])
        assert np.array_equal(C.clip_preprocess(m["images"], 56, True), want.astype(np.float32)), i
    with pytest.raises(ValueError):
        B200DPODataset([{"queries": "q", **text}])[0]


def test_make_dpo_data_module_vs_reference(tmp_path, monkeypatch):
    """collate.make_dpo_data_module against the reference's make_dpo_data_module (data_utils_dpo.py:353-370) run verbatim on
    the same on-disk HF dataset: keys, dataset length and items' text fields, the collator's hard-coded 128 / 896 lengths and
    detailed_report flag."""
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference")
    import base64
    import io
    from types import SimpleNamespace
    datasets = pytest.importorskip("datasets")
    from PIL import Image
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
    from oracle.make_golden import install_stubs
    install_stubs()
    from utils import data_utils_dpo as D
    from opa_dpo_b200 import collate as K
    rs = np.random.RandomState(1)
    rows = []
    for i in range(3):
        buf = io.BytesIO()
        Image.fromarray(rs.randint(0, 256, (30 + i, 44, 3)).astype(np.uint8)).save(buf, format="PNG")
        rows.append({"queries": f"<image>\nquestion {i}?", "image_bytes": base64.b64encode(buf.getvalue()).decode(),
                     "standard_response": f"std {i}", "original_generate_response": f"orig {i}", "AI_pseudo_response": f"ai {i}",
                     "AI_json_report": "{}"})
    path = str(tmp_path / "opadpo_ds")
    datasets.Dataset.from_list(rows).save_to_disk(path)
    proc = CLIPImageProcessorPil(size={"shortest_edge": 56}, crop_size={"height": 56, "width": 56})
    data_args = SimpleNamespace(data_path=path, image_processor=proc, image_aspect_ratio="pad")
    args = SimpleNamespace(detailed_report=True)
    tok = SimpleNamespace(pad_token_id=0, eos_token_id=2)
    want = D.make_dpo_data_module(tok, args, data_args)
    got = K.make_dpo_data_module(tok, args, data_args, ops=object())
    assert list(got) == list(want) == ["train_dataset", "eval_dataset", "data_collator"] and got["eval_dataset"] is None
    assert len(got["train_dataset"]) == len(want["train_dataset"]) == 3
    for i in range(3):
        r, m = want["train_dataset"][i], got["train_dataset"][i]
        assert set(r) == set(m) and all(r[k] == m[k] for k in r if k != "images")
        assert np.array_equal(C.clip_preprocess(m["images"], 56, True), np.asarray(r["images"], dtype=np.float32))
    cw, cg = want["data_collator"], got["data_collator"]
    assert (cg.query_len, cg.response_len, cg.detailed_report) == (cw.query_len, cw.response_len, cw.detailed_report) == (128, 896, True)
    assert cg.pad_to_square is True and cg.tokenizer is tok
    assert K.make_dpo_data_module(tok, args, SimpleNamespace(data_path=path, image_aspect_ratio="square"), ops=object())["data_collator"].pad_to_square is False
