"""Reduce `ncu -i report.ncu-rep --page raw --csv` files (one row per profiled launch) to the per-kernel evidence rows
DESIGN.md quotes: duration, DRAM bytes and GB/s against the measured copy peak, tensor-pipe activity, SM throughput.
  python tools/ncu_csv_summary.py gpurun_out/r02_ncu_full_*.csv > profiles/r02_ncu_full_kernels.txt
Numbers come from a profiler run (serialised, cold cache, unlocked clocks): evidence of where a kernel stands against its
roofline, never a bench value."""
import csv
import json
import re
import sys

HBM_PEAK = 6551.0  # GB/s, MEASURED_PEAKS.json copy bandwidth
M = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2hit": "lts__t_sector_hit_rate.pct",
    "regs": "launch__registers_per_thread",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "clk": "sm__cycles_elapsed.max.per_second",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def short(name):
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((int|bool)\)", "", name)
    name = name.replace("void opadpo::<unnamed>::", "").replace("opadpo::<unnamed>::", "")
    return name.split("(")[0][:64] if "<" not in name.split("(")[0] else name[:64]


def main():
    rows_out = []
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path, newline="")))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(head)}

        def val(r, key, scale=True):
            m = M[key]
            if m not in col or r[col[m]] in ("", "n/a"):
                return None
            v = float(r[col[m]].replace(",", ""))
            return v * SCALE.get(units[col[m]], 1.0) if scale else v

        for r in rows[2:]:
            if len(r) <= col["Kernel Name"]:
                continue
            us = val(r, "dur")
            by = (val(r, "rd") or 0.0) + (val(r, "wr") or 0.0)
            rows_out.append({"kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                             "us": us, "dram_MB": by / 1e6, "GBps": by / us / 1e3 if us else None,
                             "hbm_frac": by / us / 1e3 / HBM_PEAK if us else None, "tensor_pct": val(r, "tensor", False),
                             "sm_pct": val(r, "sm", False), "l2_hit": val(r, "l2hit", False), "regs": val(r, "regs", False),
                             "occ_pct": val(r, "occ", False), "ghz": val(r, "clk", False), "src": path.split("/")[-1]})
    print(f"{'kernel':64s} {'grid':>14s} {'us':>9s} {'dramMB':>9s} {'GB/s':>7s} {'ofHBM':>6s} {'tensor%':>7s} {'SM%':>6s} {'L2hit':>6s} {'regs':>5s}")
    for d in rows_out:
        f = lambda x, w, p=1: (f"{x:{w}.{p}f}" if x is not None else " " * (w - 1) + "-")
        print(f"{d['kernel']:64s} {d['grid']:>14s} {f(d['us'], 9)} {f(d['dram_MB'], 9)} {f(d['GBps'], 7, 0)} {f(d['hbm_frac'], 6, 2)} "
              f"{f(d['tensor_pct'], 7)} {f(d['sm_pct'], 6)} {f(d['l2_hit'], 6)} {f(d['regs'], 5, 0)}")
    if "--json" in sys.argv:
        json.dump(rows_out, sys.stderr)


if __name__ == "__main__":
    main()
