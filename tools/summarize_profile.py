"""Aggregate a HYVAE_PROFILE_DUMP csv by (class, tag): total ms, TFLOP/s or GB/s."""
import csv, sys, collections
names = ["conv_tc", "conv_direct", "gn_stats", "gn_apply", "pad_upsample", "softmax", "layout", "blend", "temporal", "attn", "attn_proj"]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(open(sys.argv[1])):
    a = agg[(int(r["class"]), r["tag"])]
    a[0] += 1; a[1] += float(r["work"]); a[2] += float(r["ms"])
tot = sum(a[2] for a in agg.values())
print(f"total device ms in profiled kernels: {tot:.1f}")
print(f"{'class':12s} {'tag':52s} {'n':>6s} {'ms':>9s} {'%':>6s} {'rate':>10s}")
for (c, tag), (n, w, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    rate = w / ms / 1e9 if ms > 0 else 0
    unit = "TFLOP/s" if c < 2 else "TB/s"
    print(f"{names[c]:12s} {tag:52s} {n:6d} {ms:9.2f} {100 * ms / tot:6.2f} {rate:8.1f} {unit}")
