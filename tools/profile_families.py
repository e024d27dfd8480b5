"""Aggregate a HYVAE_PROFILE_DUMP csv (bench.py) by kernel family: ms per step and algorithmic rate."""
import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/profile_dump.csv"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
names = ["conv_tc", "conv_direct", "gn_stats", "gn_apply", "pad_upsample", "softmax", "layout", "blend", "temporal", "attn", "attn_proj"]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(open(path)):
    tag, c = r["tag"], int(r["class"])
    if c == 0:
        m = re.match(r"k\d (\d+)->(\d+)", tag)
        if tag.startswith("up"):
            key = "conv: upsample phases (pair kernel)"
        elif "wino" in tag:
            key = f"conv: wino {m.group(1)}->{m.group(2)}"
        elif "halo" in tag:
            key = f"conv: halo {m.group(1)}->{m.group(2)}"
        elif tag.startswith("k1"):
            key = "conv: k=1 (shortcuts, attention GEMMs)"
        elif "s111" not in tag:
            key = "conv: strided (downsample)"
        else:
            key = f"conv: pair {m.group(1)}->{m.group(2)}"
    else:
        key = names[c]
    a = agg[key]; a[0] += 1; a[1] += float(r["work"]); a[2] += float(r["ms"])
tot = sum(a[2] for a in agg.values())
print(f"total device ms per step in profiled kernels: {tot / steps:.1f}")
for k, (n, w, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    unit = "TFLOP/s (algorithmic)" if (k.startswith("conv") or k.startswith("attn")) else "TB/s (algorithmic)"
    print(f"{k:42s} launches/step={n // steps:6d}  ms/step={ms / steps:8.1f}  {100 * ms / tot:5.1f}%  {w / ms / 1e9:8.1f} {unit}")
