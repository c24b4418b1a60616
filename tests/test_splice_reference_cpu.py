"""Row a7: the multimodal splice. The LLaVA fork the reference trains on is not vendored, but its patch
(llava_setup/llava_modifications.patch) carries the new `prepare_inputs_labels_for_multimodal` and its helpers as pure
additions. This test lifts those functions out of the patch TEXT, executes them unmodified, and compares the oracle's
`splice` (which the engine's embed_splice kernel is tested against) with them: token rows with the projected patch rows at
the -200 slot, the attention mask with ones — or the CoPO image-token mask — on the image rows. Needs /root/reference."""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import llava_oracle as O  # noqa: E402

PATCH = "/root/reference/llava_setup/llava_modifications.patch"


def _added_block(lines, start_marker, stop_marker):
    """The run of '+' lines from the one containing start_marker up to (not including) the one containing stop_marker."""
    i = next(k for k, l in enumerate(lines) if start_marker in l and l.startswith("+"))
    out = []
    for l in lines[i:]:
        if stop_marker in l:
            break
        assert l.startswith("+"), f"not a pure addition: {l[:80]!r}"
        out.append(l[1:])
    return out


def _reference_arch():
    with open(PATCH) as f:
        lines = f.read().split("\n")
    first = next(k for k, l in enumerate(lines) if "def _get_slices_between_image_tokens" in l)
    assert lines[first - 1].strip() == "+    @staticmethod".strip() or "@staticmethod" in lines[first - 1]
    body = ["    @staticmethod"] + _added_block(lines[first:], "def _get_slices_between_image_tokens", "def prepare_inputs_labels_for_multimodal_RL")
    body += _added_block(lines, "    def prepare_inputs_labels_for_multimodal(", "def initialize_vision_tokenizer")
    src = "class RefArch:\n" + "\n".join(body) + "\n"
    from torch.nn.utils.rnn import pad_sequence
    from typing import Iterable
    ns = dict(torch=torch, Tensor=torch.Tensor, Iterable=Iterable, TYPE_CHECKING=False, IMAGE_TOKEN_INDEX=-200, IGNORE_INDEX=-100,
              pad_sequence=pad_sequence)
    exec(compile(src, "llava_modifications.patch[llava_arch.py additions]", "exec"), ns)
    return ns["RefArch"]


@pytest.mark.parametrize("copo_mask", [False, True])
def test_oracle_splice_vs_reference_patch_code(copo_mask):
    if not os.path.exists(PATCH):
        pytest.skip("needs /root/reference")
    RefArch = _reference_arch()
    cfg = O.OracleConfig()
    w = O.init_weights(cfg)
    P, d = cfg.n_patches, cfg.d_model
    g = torch.Generator().manual_seed(3)
    images, q, qm, resp = O.synthetic_batch(cfg, 3, 16, 9, ["a_response", "b_response"], left_pad=2)
    ids = torch.cat([torch.cat([q, resp["a_response"]], 1), torch.cat([q, resp["b_response"]], 1)], 0)      # K stacked rows
    mask = torch.cat([torch.cat([qm.bool(), resp["a_response"] != 0], 1), torch.cat([qm.bool(), resp["b_response"] != 0], 1)], 0)
    feats = torch.randn(ids.size(0), P, d, generator=g)
    img_mask = (torch.rand(ids.size(0), P, generator=g) > 0.3) if copo_mask else None

    emb = torch.nn.Embedding(cfg.vocab, d)
    emb.weight.data.copy_(w["model.embed_tokens.weight"])

    class Arch(RefArch):
        config = SimpleNamespace()

        def get_model(self):
            return SimpleNamespace(embed_tokens=emb, device=torch.device("cpu"))

        def _encode_image_batch(self, images):
            return images              # the test hands over projected features directly

    # the reference's convention for the CoPO mask: the image part is PREPENDED to the row's mask (rl_models.py:103-105)
    attn = mask if img_mask is None else torch.cat([img_mask, mask], 1)
    with torch.no_grad():
        ref_embeds, ref_labels, ref_mask = Arch().prepare_inputs_labels_for_multimodal(ids, attn, None, feats)
        got_embeds, got_mask = O.splice(cfg, w, ids, mask, feats, img_mask)
    assert ref_labels is None
    assert got_embeds.shape == ref_embeds.shape == (ids.size(0), ids.size(1) - 1 + P, d)
    assert torch.equal(got_embeds, ref_embeds)
    assert torch.equal(got_mask, ref_mask.bool())
    slot = (ids == -200).float().argmax(1)
    for n in range(ids.size(0)):
        assert torch.equal(got_embeds[n, slot[n]:slot[n] + P], feats[n])
        assert got_mask[n, slot[n]:slot[n] + P].all() or copo_mask


def test_oracle_projector_vs_reference_patch_builder():
    """mm_projector: `build_vision_projector` lifted from the patch (multimodal_projector/builder.py, a new file there),
    built for mm_projector_type = 'mlp2x_gelu' with the oracle's weights, against the oracle's projector (exact-erf GELU)."""
    if not os.path.exists(PATCH):
        pytest.skip("needs /root/reference")
    import re
    import torch.nn as nn
    with open(PATCH) as f:
        lines = f.read().split("\n")
    block = _added_block(lines, "def build_vision_projector(", "diff -ruN")
    ns = dict(nn=nn, re=re, IdentityMap=nn.Identity)
    exec(compile("\n".join(block) + "\n", "llava_modifications.patch[builder.py]", "exec"), ns)
    cfg = O.OracleConfig()
    w = O.init_weights(cfg)
    proj = ns["build_vision_projector"](SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=cfg.v_d, hidden_size=cfg.d_model))
    assert [type(m).__name__ for m in proj] == ["Linear", "GELU", "Linear"]
    with torch.no_grad():
        for i in (0, 2):
            proj[i].weight.copy_(w[f"model.mm_projector.{i}.weight"])
            proj[i].bias.copy_(w[f"model.mm_projector.{i}.bias"])
        feats = torch.randn(2, cfg.n_patches, cfg.v_d, generator=torch.Generator().manual_seed(1))
        want = proj(feats)
        got = O.projector(cfg, w, None, feats)
    assert torch.allclose(got, want, atol=1e-6, rtol=1e-5)
