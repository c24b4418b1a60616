"""Grep-able SASS evidence for the tcgen05 / TMA / TMEM kernels (the full `cuobjdump -sass` of gemm.o alone is 31 MB).
For every kernel of the given object files: instruction count, counts of the Blackwell-specific mnemonics
(UTCHMMA[.2CTA] = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
MUFU.EX2) and of the legacy tensor path (HMMA: must be 0), followed by every SASS line that carries one of the tensor / TMA /
TMEM mnemonics, verbatim.   python tools/sass_summary.py build/csrc/gemm.o > profiles/r02_sass_gemm.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP", "SYNCS", "MUFU.EX2",
        "RED", "HMMA", "HGMMA"]
VERBATIM = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR")


def main():
    for obj in sys.argv[1:]:
        out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True, check=True).stdout
        name, kernels = None, collections.OrderedDict()
        for ln in out.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
                name = re.sub(r"\(CUtensorMap_st.*", "", name)
                kernels[name] = []
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and name is not None:
                kernels[name].append((m.group(1), m.group(2).strip()))
        print(f"==== {obj}: {len(kernels)} kernels (cuobjdump -sass, sm_100a)")
        for name, ins in kernels.items():
            cnt = collections.Counter()
            for _, text in ins:
                op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
                for k in KEYS:
                    if op == k or op.startswith(k + ".") or (k == "UTCHMMA.2CTA" and op.startswith("UTCHMMA") and ".2CTA" in op):
                        cnt[k] += 1
            print(f"\n-- {name}\n   instructions {len(ins)}; " + ", ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k] or k in ("HMMA",)))
            shown = 0
            for addr, text in ins:
                op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
                if op.startswith(VERBATIM) and shown < 24:
                    print(f"     /*{addr}*/ {text}")
                    shown += 1
            if shown == 24:
                print("     ... (further tensor / TMA / TMEM lines of this kernel omitted)")


if __name__ == "__main__":
    main()
