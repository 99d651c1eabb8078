"""Per-class summary of an ops-profile CSV written by mdc_dbg_profile_ops (tools/gpu_profile_step.py, OPS_CSV=...)."""
import collections
import csv
import re
import sys


def cls(n):
    if "norm" in n and "transformer" not in n:
        return "groupnorm"
    if re.search(r"norm[123]$", n):
        return "layernorm"
    if "attn1" in n and "to_" not in n:
        return "attn1 (flash)"
    if "attn2" in n and "to_" not in n:
        return "attn2"
    if "sdpa" in n:
        return "sdpa d512"
    if "ff.net.0.proj" in n:
        return "ff_in"
    if "ff.net.0" in n:
        return "geglu"
    if "ff.net.2" in n:
        return "ff_out"
    if "to_qkv" in n:
        return "qkv"
    if "to_q" in n:
        return "attn2.to_q"
    if "to_out" in n:
        return "to_out"
    if "proj_in" in n or "proj_out" in n:
        return "proj_in/out"
    if "conv_shortcut" in n:
        return "shortcut"
    if "upsamplers" in n:
        return "upconv"
    if "downsamplers" in n:
        return "downsample"
    if "conv" in n:
        return "conv3x3"
    if n.startswith("cat"):
        return "cat"
    return "other"


def main(path):
    rows = list(csv.DictReader(open(path)))
    print(f"| tape | class | ops | fwd us | bwd us | GEMM GFLOP fwd | launches |")
    print("|---|---|---|---|---|---|---|")
    for tape in ("unet", "dec"):
        d = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
        for r in rows:
            if r["tape"] != tape:
                continue
            e = d[cls(r["name"])]
            e[0] += float(r["fwd_us"]); e[1] += float(r["bwd_us"]); e[2] += 1
            e[3] += float(r["gemm_gflop_fwd"]); e[4] += int(r["launches_fwd"]) + int(r["launches_bwd"])
        tot = sum(e[0] + e[1] for e in d.values())
        for k, e in sorted(d.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
            if e[2]:
                print(f"| {tape} | {k} | {e[2]} | {e[0]:.0f} | {e[1]:.0f} | {e[3]:.1f} | {e[4]} |")
        print(f"| {tape} | **total** | | **{tot:.0f}** | | | |")


if __name__ == "__main__":
    main(sys.argv[1])
