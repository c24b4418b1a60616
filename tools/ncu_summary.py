"""Reduce an `ncu --set full` report to the handful of per-launch numbers DESIGN.md quotes.
Usage: python tools/ncu_summary.py report.ncu-rep [out.json]   (needs ncu on PATH; runs without a GPU)"""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    out = []
    for r in rows[2:]:
        d = {"kernel": r[col["Kernel Name"]].split("(")[0]}
        for m, k in KEEP.items():
            if m in col:
                d[k] = f"{r[col[m]]} {units[col[m]]}".strip()
        try:
            us = float(r[col["gpu__time_duration.sum"]]) * SCALE[units[col["gpu__time_duration.sum"]]]
            by = sum(float(r[col[m]]) * SCALE[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            d.update(dur_us=round(us, 1), dram_bytes=int(by), dram_GBps=round(by / us / 1e3, 1))
        except (KeyError, ValueError):
            pass
        out.append(d)
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    else:
        print(text)


if __name__ == "__main__":
    main()
