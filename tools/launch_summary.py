"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` launch list to per-kernel totals of ONE step (the window between
the last two adamw_kernel launches, or --between A B for another delimiter / the whole file with --all).
  python tools/launch_summary.py gpurun_out/launches_step.csv [--all] [--delim adamw_kernel] > profiles/rNN_launch_summary_step.txt
The per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes."""
import argparse
import collections
import csv
import re


def short(name: str) -> str:
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    name = name.split("(")[0] if "gemm_kernel" not in name and "attn_" not in name else name
    return name.replace("opadpo::<unnamed>::", "opadpo::")[:110]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--delim", default="adamw_kernel")
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    for r in rd:
        if len(r) <= vi:
            continue
        try:
            rows.append((r[ki], float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3)))
        except ValueError:
            pass
    lo, hi = 0, len(rows)
    if not a.all:
        marks = [i for i, (n, _) in enumerate(rows) if a.delim in n]
        if len(marks) >= 2:
            lo, hi = marks[-2] + 1, marks[-1] + 1
    win = rows[lo:hi]
    tot = sum(t for _, t in win)
    agg = collections.defaultdict(lambda: [0.0, 0])
    for n, t in win:
        s = short(n)
        agg[s][0] += t
        agg[s][1] += 1
    print(f"launches {lo}..{hi} of {len(rows)}: {len(win)} launches, {tot / 1e3:.1f} ms summed kernel time (cold-cache, serialised: compare SHARES)")
    for n, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{t / 1e3:9.2f} ms {100 * t / tot:5.1f}%  n={c:5d}  avg={t / c:9.1f} us  {n}")


if __name__ == "__main__":
    main()
