"""Per-kernel counts of the Blackwell-native SASS mnemonics in the shipped library (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): UTC*MMA = tcgen05.mma, UTMALDG / UTMASTG = TMA loads / stores, LDTM / STTM = tcgen05.ld / st,
HMMA = legacy mma.sync (none expected), UBLKCP = cp.async.bulk (1-D bulk copies), UCGABAR = cluster barriers.  Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "depth_completion_b200", "lib", "libmdc_b200.so")
PAT = {"UTC*MMA": re.compile(r"\bUTC[A-Z]*MMA\b"), "UTMALDG": re.compile(r"\bUTMALDG\b"), "UTMASTG": re.compile(r"\bUTMASTG\b"),
       "LDTM": re.compile(r"\bLDTM\b"), "STTM": re.compile(r"\bSTTM\b"), "HMMA": re.compile(r"\bHMMA\b"),
       "UBLKCP": re.compile(r"\bUBLKCP\b"), "UCGABAR": re.compile(r"\bUCGABAR"), "MUFU": re.compile(r"\bMUFU\b"),
       "FFMA2": re.compile(r"\bFFMA2\b")}
NATIVE = ("UTC*MMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "UBLKCP", "UCGABAR")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    widths = collections.defaultdict(collections.Counter)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = cur.replace("mdc::", "").replace("void ", "")
            if cur not in order:
                order.append(cur)
            continue
        if cur:
            for k, p in PAT.items():
                if p.search(line):
                    counts[cur][k] += 1
            m2 = re.search(r"\b(LDG|STG)(\.[A-Z0-9]+)*", line)
            if m2:
                op = m2.group(0)
                w = "128" if ".128" in op else "64" if ".64" in op else None if (".U8" in op or ".U16" in op or ".S8" in op or ".S16" in op) else "32"
                if w:
                    widths[cur][m2.group(1) + w] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel")
    print("| kernel | " + " | ".join(PAT) + " |")
    print("|---|" + "---|" * len(PAT))
    for k in order:
        c = counts[k]
        if any(c[x] for x in NATIVE):
            print(f"| {k} | " + " | ".join(str(c[x]) for x in PAT) + " |")
    others = [k for k in order if not any(counts[k][x] for x in NATIVE)]
    print(f"\n{len(others)} further kernels are bandwidth-bound CUDA-core kernels without tensor / TMA / bulk-copy / cluster instructions "
          f"(GroupNorm, LayerNorm, GEGLU, the step tail, ...): " + ", ".join(sorted(set(o.split("<")[0].split("::")[-1] for o in others))))
    # global access widths of the bandwidth kernels (a BF8 channel vector must move as ONE 16-byte access; before the copy
    # operations of BF8 moved a uint4 these kernels issued four 32-bit accesses per vector)
    print("\n## global load / store widths of the non-GEMM kernels that move channel vectors (instruction counts)")
    print("| kernel | LDG.128 | LDG.64 | LDG.32 | STG.128 | STG.64 | STG.32 |")
    print("|---|---|---|---|---|---|---|")
    for k in order:
        if widths[k]["LDG128"] + widths[k]["STG128"] == 0 or any(counts[k][x] for x in ("UTC*MMA",)) or "cub::" in k:
            continue
        w = widths[k]
        print(f"| {k} | {w['LDG128']} | {w['LDG64']} | {w['LDG32']} | {w['STG128']} | {w['STG64']} | {w['STG32']} |")
    f2 = [k for k in order if counts[k]["FFMA2"]]
    print("\nkernels using packed fp32x2 FMAs (FFMA2): " + ", ".join(f"{k.split('<')[0]} ({counts[k]['FFMA2']})" for k in f2))


if __name__ == "__main__":
    main()
