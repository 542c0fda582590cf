#!/usr/bin/env python3
"""Steady-state statistics of a `bench.py --timeline FILE` capture: which kernels hold the GPU, how much of the
time a bucket accumulation is running, and an estimate of the register-file share each kernel family takes.
usage: timeline_stats.py FILE [lo_frac hi_frac]     (window of the capture to analyse, default 0.33 0.67)"""
import re
import sys
from collections import defaultdict

path = sys.argv[1]
lo_f, hi_f = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (0.33, 0.67)
rows = []
for ln in open(path):
    if ln.startswith("#"):
        continue
    p = ln.split(None, 6)
    if len(p) < 7:
        continue
    try:
        rows.append(dict(t=float(p[0]), d=float(p[1]), st=p[2], grid=p[3], block=p[4], regs=p[5], name=p[6].strip()))
    except ValueError:
        pass
end = max(r["t"] + r["d"] for r in rows)
lo, hi = lo_f * end, hi_f * end
span = hi - lo
SMS, RF = 148, 65536


def prod(s):
    v = 1
    for x in s.split("x"):
        v *= int(x) if x.isdigit() else 1
    return v


def fam(nm):
    nm = nm.replace("void ", "")
    m = re.match(r"hk::(\w+)", nm)
    base = m.group(1) if m else nm.split("(")[0][:24]
    if "Fp2T" in nm.split("(")[0]:
        base += "[G2]"
    return base


wall = defaultdict(float)
rf = defaultdict(float)
acc_iv = []
for r in rows:
    s, e = max(r["t"], lo), min(r["t"] + r["d"], hi)
    if e <= s:
        continue
    f = fam(r["name"])
    wall[f] += e - s
    if r["regs"].isdigit() and r["grid"] != "-":
        threads = prod(r["block"])
        regs_per_cta = max(int(r["regs"]), 16) * threads
        resident = min(prod(r["grid"]), SMS * max(1, min(RF // regs_per_cta, 2048 // threads, 32)))
        rf[f] += (e - s) * resident * regs_per_cta / (SMS * RF)
    if "k_msm_accum" in r["name"]:
        acc_iv.append((s, e))
acc_iv.sort()
cov, ce = 0.0, lo
for s, e in acc_iv:
    if e > ce:
        cov += e - max(s, ce)
        ce = e
print(f"window {lo / 1e3:.1f} .. {hi / 1e3:.1f} ms of {end / 1e3:.1f} ms; a bucket accumulation is running {100 * cov / span:.1f} % of it")
print(f"{'kernel':<28} {'wall ms':>9} {'x window':>9} {'RF-share ms (upper bound)':>26}")
for f, v in sorted(wall.items(), key=lambda x: -x[1])[:18]:
    print(f"{f:<28} {v / 1e3:9.2f} {v / span:9.2f} {rf.get(f, 0) / 1e3:26.2f}")
