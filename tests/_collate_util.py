"""Helpers shared by the CPU and GPU collator tests (input generators, golden accessors)."""
import json

import numpy as np

from oracle import collate_oracle as C


def unpack_rows(G, prefix):
    """flat segment arrays -> the oracle's list-of-lists form."""
    tok, so = G[prefix + "tokens"], G[prefix + "seg_offsets"]
    sc, rel, fl, ro = G[prefix + "seg_score"], G[prefix + "seg_relation"], G[prefix + "seg_flags"], G[prefix + "row_offsets"]
    rows = []
    for i in range(len(ro) - 1):
        rows.append([(tok[so[j]:so[j + 1]].tolist(), float(sc[j]), float(rel[j]), int(fl[j])) for j in range(ro[i], ro[i + 1])])
    return rows


def random_rows(rs, b, max_segs, max_len, R):
    """Ragged random reports incl. empty rows, BOS-only entries, leading-29871 entries, pad / EOS ids inside the text."""
    rows = []
    for i in range(b):
        segs = []
        for j in range(rs.randint(0, max_segs + 1)):
            n = rs.randint(0, max_len + 1)
            raw = [1] + rs.randint(3, 32000, n).tolist()
            if n and rs.rand() < 0.3:
                raw[1] = C.SPACE_TOKEN
            if n > 2 and rs.rand() < 0.15:
                raw[rs.randint(1, n + 1)] = rs.choice([0, 2])
            fl = (C.FLAG_KEEP_BOS if j == 0 else 0) | (C.FLAG_STRIP_SPACE if rs.rand() < 0.5 else 0)
            segs.append((raw, float(rs.choice([1.0, 1.5, 2.0, 2.5, 0.0])), float(rs.choice([1.0, 3.0])), fl))
        rows.append(segs)
    return rows


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
