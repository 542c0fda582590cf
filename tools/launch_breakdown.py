#!/usr/bin/env python3
"""Per-kernel shares of ONE bench step from an ncu launch list (gpu__time_duration.sum CSV)."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"], r.get("Grid Size", ""), r.get("Block Size", "")))
        except Exception:
            pass
    # a bench step = one commitment batch + one proof batch: from a k_build_commit_scalars launch to the next
    # k_prove_finish_c.  Every such segment of the modal length is the same work; report the second one (the first
    # belongs to the trapdoor verification, which also runs a single-proof call in front).
    segs, start = [], None
    for i, r in enumerate(rows):
        if "k_build_commit_scalars" in r[0] and start is None:
            start = i
        elif "k_build_commit_scalars" in r[0]:
            start = i if not any("k_prove" in x[0] for x in rows[start:i]) else start
        if "k_prove_finish_c" in r[0] and start is not None:
            segs.append((start, i + 1))
            start = None
    lens = collections.Counter(b - a for a, b in segs)
    modal = lens.most_common(1)[0][0]
    same = [(a, b) for a, b in segs if b - a == modal]
    a, b = same[1] if len(same) > 1 else same[0]
    seg = rows[a:b]
    tot = sum(r[1] for r in seg)
    agg = collections.OrderedDict()
    for name, v, u, g, b in seg:
        n = re.sub(r"<.*", "", name.split("(")[0]).replace("hk::", "").replace("void ", "")
        key = n + (" [G2]" if "Fp2T" in name else "")
        agg.setdefault(key, [0, 0.0])
        agg[key][0] += 1
        agg[key][1] += v
    print(f"one step (commitment batch + proof batch): {len(seg)} launches, {tot / 1e6:.2f} ms summed kernel time ({seg[0][2]})")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {v / 1e6:.3f} | {100 * v / tot:.1f} % |")
    if "-v" in sys.argv:
        for name, v, u, g, b in seg:
            n = re.sub(r"<.*", "", name.split("(")[0]).replace("hk::", "").replace("void ", "")
            print(f"   {n:28s} {'G2' if 'Fp2T' in name else '  '} {v/1e3:9.1f} us  grid {g} block {b}")


if __name__ == "__main__":
    main(sys.argv[1])
