"""Generates tests/golden/collate_golden.npz by running the REFERENCE's own collator-stage code (row f-3).

Run in the build container only (needs /root/reference and Pillow; the GPU box never runs this):
    python oracle/make_golden_collate.py

Executed verbatim from /root/reference:
  * utils.data_utils_dpo.DataCollatorForCausalLM.__call__  (:101-287; detailed_report on and off, the overflow fallback)
    over a deterministic whitespace tokenizer that mimics the LLaMA tokenizer's interface (BOS first, pad 0, eos 2,
    padding_side, max_length truncation). Every single-text tokenizer call the collator makes is logged, so the golden
    holds the exact raw segments the GPU collator is given.
  * utils.data_utils_dpo.DPO_Dataset.__getitem__           (:289-343; base64 image -> PIL -> expand2square -> CLIP
    preprocess) over transformers' PIL-backed CLIP image processor (the lineage of the pinned 4.34.1 slow processor).
  * opadpo.dpo_models.dpo_trainer.mask_single_image / mask_percentage_per_row (:83-125) under torch.manual_seed.
"""
from __future__ import annotations

import base64
import hashlib
import io
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import install_stubs  # noqa: E402

from oracle.collate_oracle import StubTokenizer, word_id  # noqa: E402,F401


def words(rs, n, prefix="w"):
    return " ".join(f"{prefix}{rs.randint(0, 5000)}" for _ in range(n))


def make_instances(rs, lengths):
    """Instances whose reports exercise: score / error-type weights, defaults, an 'Added' key, an empty copied entry, a
    copied segment that begins with the 29871 piece, an empty rewritten entry."""
    out = []
    for li, L in enumerate(lengths):
        report, orig_parts, ai_parts = {}, [], []
        nkeys = 4 + li
        for k in range(nkeys):
            cop = words(rs, L)
            rew = words(rs, L + (k % 3))
            if k == 1:
                cop = "SPACE " + cop
            entry = {"copied content": cop, "rewritten content": rew}
            if k % 2 == 0:
                entry["score"] = 1 + (k + li) % 4
            if k % 3 == 0:
                entry["error type"] = ["image_recognition_error", "correct", "language_comprehension_error"][(k + li) % 3]
            if k == 2:
                entry = {"copied_content": cop, "rewritten_content": "", "score": 2, "error_type": "image_recognition_error"}
            if k == 3:
                entry["copied content"] = ""
            report[f"Sentence {k + 1}"] = entry
            if entry.get("copied content", entry.get("copied_content", "")):
                orig_parts.append(cop)
            if entry.get("rewritten content", entry.get("rewritten_content", "")):
                ai_parts.append(rew)
        report["Added"] = {"rewritten content": words(rs, 3, "add")}
        report["image_description"] = {"copied content": "never used", "rewritten content": "never used"}
        out.append({"queries": "SYS " + words(rs, 5 + 9 * li, "q") + " IMG " + words(rs, 3, "q"), "images": torch.zeros(1),
                    "standard_response": words(rs, 10 + 7 * li, "s"),
                    "original_generate_response": " ".join(orig_parts),
                    "AI_pseudo_response": " ".join(ai_parts), "AI_json_report": json.dumps(report)})
    return out


def segments_from_log(instances, log, tok):
    """Rebuild the per-row segment lists (raw ids, score weight, relation weight, flags) from the logged tokenizer calls:
    per report key the collator tokenizes the rewritten entry first, then (unless the key is 'Added') the copied one."""
    from oracle import collate_oracle as C
    it = iter(log)
    orig_rows, ai_rows = [], []
    for inst in instances:
        rep = {k: v for k, v in json.loads(inst["AI_json_report"]).items() if k not in ("image_description", "image description")}
        o, a = [], []
        for count, (key, e) in enumerate(rep.items()):
            keep = C.FLAG_KEEP_BOS if count == 0 else 0
            _, rew = next(it)
            if key != "Added":
                _, cop = next(it)
                ws = C.SCORE_MAPPING.get(e.get("score", 4), 1.0)
                wr = C.ERROR_TYPE_MAPPING.get(e.get("error type", e.get("error_type", "correct")), 1.0)
                o.append((cop, ws, wr, keep | C.FLAG_STRIP_SPACE))
                a.append((rew, ws, wr, keep))
            else:
                a.append((rew, 1.0, 1.0, keep))
        orig_rows.append(o)
        ai_rows.append(a)
    assert next(it, None) is None
    return orig_rows, ai_rows


def pack_rows(rows):
    """segment lists -> flat arrays (the device collator's input format)."""
    toks, seg_off, sc, rel, fl, row_off = [], [0], [], [], [], [0]
    for segs in rows:
        for raw, ws, wr, f in segs:
            toks.extend(raw)
            seg_off.append(len(toks))
            sc.append(ws)
            rel.append(wr)
            fl.append(f)
        row_off.append(len(sc))
    return dict(tokens=np.asarray(toks, np.int32), seg_offsets=np.asarray(seg_off, np.int64),
                seg_score=np.asarray(sc, np.float32), seg_relation=np.asarray(rel, np.float32),
                seg_flags=np.asarray(fl, np.uint8), row_offsets=np.asarray(row_off, np.int64))


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    install_stubs()
    from utils import data_utils_dpo as D
    from opadpo.dpo_models import dpo_trainer as T
    from PIL import Image
    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil

    G = {}
    rs = np.random.RandomState(7)
    Q, R = 24, 64

    # ---- (A) detailed report
    tok = StubTokenizer()
    col = D.DataCollatorForCausalLM(tokenizer=tok, query_len=Q, response_len=R, detailed_report=True)
    inst = make_instances(rs, [2, 3, 4])
    tok.log.clear()
    batch = col(inst)
    assert batch["AI_pseudo_response_scores"].dtype == torch.float32, "fell back: a report overflowed R"
    orig_rows, ai_rows = segments_from_log(inst, tok.log, tok)
    for name, rows in (("orig", orig_rows), ("ai", ai_rows)):
        for k, v in pack_rows(rows).items():
            G[f"A_{name}_{k}"] = v
    for k in ("queries", "queries_attention_mask", "standard_response", "original_generate_response",
              "original_generate_response_attention_mask", "AI_pseudo_response", "AI_pseudo_response_attention_mask",
              "original_generate_response_scores", "AI_pseudo_response_scores",
              "original_generate_response_image_relations", "AI_pseudo_response_image_relations"):
        G["A_out_" + k] = batch[k].numpy()
    G["A_instances"] = np.asarray(json.dumps([{k: v for k, v in i.items() if k != "images"} for i in inst]))
    G["A_query_tokens"] = np.asarray(json.dumps([tok.encode_one(i["queries"]) for i in inst]))
    G["A_standard_tokens"] = np.asarray(json.dumps([tok.encode_one(i["standard_response"]) for i in inst]))
    G["A_QR"] = np.asarray([Q, R])

    # ---- (B) a report that overflows R: the reference falls back to plain tokenisation with zero (int64) scores
    tok = StubTokenizer()
    col = D.DataCollatorForCausalLM(tokenizer=tok, query_len=Q, response_len=R, detailed_report=True)
    inst_b = make_instances(rs, [2, 20])
    batch = col(inst_b)
    assert batch["AI_pseudo_response_scores"].dtype == torch.int64
    G["B_instances"] = np.asarray(json.dumps([{k: v for k, v in i.items() if k != "images"} for i in inst_b]))
    for k in ("original_generate_response", "AI_pseudo_response", "original_generate_response_scores",
              "AI_pseudo_response_scores", "AI_pseudo_response_image_relations", "standard_response"):
        G["B_out_" + k] = batch[k].numpy()
    G["B_plain_tokens"] = np.asarray(json.dumps({k: [tok.encode_one(i[k]) for i in inst_b] for k in
                                                 ("standard_response", "original_generate_response", "AI_pseudo_response")}))

    # ---- (C) detailed_report off: plain tokenisation with truncation (one response longer than R)
    tok = StubTokenizer()
    col = D.DataCollatorForCausalLM(tokenizer=tok, query_len=Q, response_len=R, detailed_report=False)
    inst_c = make_instances(rs, [2, 3])
    inst_c[1]["standard_response"] = words(rs, R + 5, "long")
    inst_c[0]["queries"] = words(rs, Q + 7, "q") + " IMG tail"      # longer than Q: truncated on the right
    inst_c[1]["standard_response"] = words(rs, R + 5, "long")
    inst_c[0]["AI_pseudo_response"] = words(rs, R - 1, "exact")      # BOS + R-1 words fills the row: no pad, no EOS
    batch = col(inst_c)
    for k in ("queries", "queries_attention_mask", "standard_response", "standard_response_attention_mask",
              "original_generate_response", "AI_pseudo_response"):
        G["C_out_" + k] = batch[k].numpy()
    G["C_tokens"] = np.asarray(json.dumps({k: [tok.encode_one(i[k]) for i in inst_c] for k in
                                           ("queries", "standard_response", "original_generate_response", "AI_pseudo_response")}))

    # ---- (D) images: DPO_Dataset.__getitem__ (base64 -> PIL -> expand2square -> CLIP preprocess)
    def run_dataset(img_u8, size, aspect):
        buf = io.BytesIO()
        Image.fromarray(img_u8).save(buf, format="PNG")
        proc = CLIPImageProcessorPil(size={"shortest_edge": size}, crop_size={"height": size, "width": size})
        ds = D.DPO_Dataset(SimpleNamespace(image_processor=proc, image_aspect_ratio=aspect),
                           [{"image_bytes": base64.b64encode(buf.getvalue()).decode(), "queries": "q <image>",
                             "standard_response": "", "original_generate_response": "", "AI_pseudo_response": "",
                             "AI_json_report": "{}"}])
        x = ds[0]["images"]
        return (x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)).astype(np.float32)

    small = [(40, 70, 56, "pad"), (91, 64, 56, "pad"), (56, 56, 56, "pad"), (300, 200, 56, "pad"), (33, 47, 56, "square"),
             (120, 75, 56, "square")]
    for i, (H, W, S, aspect) in enumerate(small):
        img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        G[f"D{i}_in"] = img
        G[f"D{i}_out"] = run_dataset(img, S, aspect)
        G[f"D{i}_cfg"] = np.asarray([S, 1 if aspect == "pad" else 0])
    big = [(480, 640, 101), (640, 427, 102), (150, 200, 103), (336, 336, 104), (1000, 750, 105)]
    hashes = {}
    for H, W, seed in big:       # full-size cases: the input is regenerated from its seed, the output pinned by hash
        img = np.random.RandomState(seed).randint(0, 256, (H, W, 3)).astype(np.uint8)
        out = run_dataset(img, 336, "pad")
        hashes[f"{H}x{W}x{seed}"] = dict(sha256=sha(out), probe=[float(v) for v in out[:, ::97, ::89].ravel()[:24]])
    G["D_big"] = np.asarray(json.dumps(hashes))

    # ---- (E) CoPO masking
    for name, shape, ratio, method, dt in (("E0", (3, 28, 42), 0.3, "random", torch.float32),
                                           ("E1", (3, 28, 42), 0.5, "blockwise", torch.float32),
                                           ("E2", (3, 28, 42), 0.3, "random", torch.bfloat16),
                                           ("E3", (3, 336, 336), 0.3, "random", torch.float32),
                                           ("E4", (3, 336, 336), 0.3, "blockwise", torch.bfloat16)):
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        img = torch.randn(shape, generator=g)
        torch.manual_seed(99)
        out = T.mask_single_image(img.to(dt).unsqueeze(0), ratio, method)[0].float().numpy()
        torch.manual_seed(99)
        total = shape[1] * shape[2] if method == "random" else (shape[1] // 14) * (shape[2] // 14)
        idx = torch.randperm(total)[: int(total * ratio)].numpy()
        G[name + "_idx"] = idx
        probe = int(idx[0]) if method == "random" else (int(idx[0]) // (shape[2] // 14)) * 14 * shape[2] + (int(idx[0]) % (shape[2] // 14)) * 14
        G[name + "_mean"] = np.asarray(out.reshape(shape[0], -1)[0, probe], np.float32)   # the fill value the reference used
        G[name + "_cfg"] = np.asarray(json.dumps(dict(shape=shape, ratio=ratio, method=method, bf16=dt == torch.bfloat16,
                                                      seed=sum(map(ord, name)))))
        if shape[1] <= 64:
            G[name + "_in"] = img.numpy()
            G[name + "_out"] = out
        else:
            G[name + "_sha"] = np.asarray(sha(out))
    torch.manual_seed(5)
    m = T.mask_percentage_per_row(torch.ones(3, 576, dtype=torch.bool), 0.3)
    torch.manual_seed(5)
    G["E_attn_idx"] = np.stack([torch.randperm(576)[: int(576 * 0.3)].numpy() for _ in range(3)])
    G["E_attn_out"] = m.numpy()

    # ---- (F) complete_copied_content (data_utils_dpo.py:56-71) on its own
    cases = [("the cat sat on the mat. it purred.", ["the cat", " sat on the mat.", ""]),
             ("alpha beta gamma delta", ["beta", "delta"]),
             ("alpha beta gamma delta", ["beta", "missing piece", "delta"]),
             ("  leading and trailing  ", [" leading", "", "trailing "]),
             ("one two three", ["", ""]),
             ("one two three", []),
             ("repeat repeat repeat tail", ["repeat", "repeat", "repeat"]),
             ("a b c", ["a b c"]),
             ("x", ["", "x", ""])]
    G["F_realign"] = np.asarray(json.dumps([[o, p, D.complete_copied_content(o, list(p))] for o, p in cases]))

    # ---- (G) rollout record writer (online_generator.py:380-396), run verbatim
    import tempfile
    from opadpo.generator_models import online_generator as OG
    rd = {"query": ["<image>\nwhat is this?", "<image>\ndescribe"], "image_id": ["a.jpg", "b.jpg"],
          "standard_response": ["a cat", "a dog \u00e9"], "original_generate_response": ["cat", "dog"],
          "AI_generate_response": ["cat.", "dog."], "AI_pseudo_response": ["a cat.", "a dog."],
          "AI_json_report": [{"Sentence 1": {"score": 4}}, {"Added": {"rewritten content": "x"}}],
          "image_bytes": [b"\x00\x01\xffjpeg", b"\x89PNG"]}
    with tempfile.TemporaryDirectory() as tmp:
        OG.Online_Generator.record_step_stats(SimpleNamespace(args=SimpleNamespace(output_dir=tmp)), 7, dict(rd))
        with open(os.path.join(tmp, "rollouts", f"step7_rank{int(os.environ.get('RANK', 0))}.json")) as f:
            G["G_rollout_json"] = np.asarray(f.read())
    rd_store = dict(rd)
    rd_store["image_bytes"] = [base64.b64encode(x).decode() for x in rd["image_bytes"]]
    G["G_rollout_input"] = np.asarray(json.dumps(rd_store))

    out_path = os.path.join(ROOT, "tests", "golden", "collate_golden.npz")
    np.savez_compressed(out_path, **G)
    print(f"wrote {out_path}: {len(G)} arrays, {os.path.getsize(out_path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
