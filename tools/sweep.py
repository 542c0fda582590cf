#!/usr/bin/env python3
"""BASELINE.json config 2: standalone G1/G2 MSM and Fr NTT sweep, 2^16 .. 2^24, every point CHECKED.

    python bench.py --sweep [--curve bls12_381|bn254] [--sweep-args "--lo 16 --hi 24 --cpu-max 20 --out FILE"]
(run through bench.py: the CPU column is bench's cpu_baseline leg, the one place the oracle may be executed)

MSM inputs: bases k_i * G generated on the device from a seeded stream, scalars
(a) uniform in [0, 2^253) and (b) witness-like (45 % zeros, 45 % ones, 10 % uniform).  Each result is
checked exactly: the expected point is (sum k_i s_i mod r) * G, one fixed-base multiplication, so
the check does not depend on any MSM algorithm.  Both modes are timed: `plain` (bases as uploaded,
one bucket set per window, Horner combine) and `table` (window-shifted table resident in HBM).
NTT: forward transform, checked by out[0] = sum x_j, the round trip, a known answer for a delta input and -- for the
sizes that have a CPU column -- element by element against the oracle's transform of the same input; reports
algorithmic GB/s (64 B per element) and butterflies/s.
The CPU column is the oracle's C restatement of the arkworks-0.4 algorithms on all host cores
(sizes <= 2^cpu-max so the sweep stays bounded).  Times are CUDA events on the context stream,
best of 3 after one warm-up, inputs resident in HBM."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from hekaton_system_b200 import codec
from hekaton_system_b200.backend import Context


def best_ms(stream, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def limb_dot_mod(ks, sc, r):
    """sum_i ks[i] * sc[i] mod r for (n, 4) uint64 limb arrays, in chunks of Python big integers."""
    tot = 0
    step = 1 << 16
    for o in range(0, len(ks), step):
        a = codec.limbs_to_ints(ks[o:o + step])
        b = codec.limbs_to_ints(sc[o:o + step])
        tot += sum(x * y for x, y in zip(a, b))
    return tot % r


def limb_sum_mod(a, r):
    """sum of (n, 4) uint64 limb rows as integers, mod r (numpy on 32-bit halves)."""
    lo = (a & np.uint64(0xffffffff)).sum(axis=0, dtype=np.uint64)
    hi = (a >> np.uint64(32)).sum(axis=0, dtype=np.uint64)
    tot = 0
    for j in range(4):
        tot += (int(lo[j]) + (int(hi[j]) << 32)) << (64 * j)
    return tot % r


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--curve", default="bls12_381")
    ap.add_argument("--lo", type=int, default=16)
    ap.add_argument("--hi", type=int, default=24)
    ap.add_argument("--g2-hi", type=int, default=24)
    ap.add_argument("--cpu-max", type=int, default=20, help="largest log size timed on the CPU (0 = none)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--ops", default="msm_g1,msm_g2,ntt", help="comma-separated subset of msm_g1,msm_g2,ntt")
    args = ap.parse_args(argv)
    ops = set(args.ops.split(","))

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = Context(args.curve, 0, stream=stream.cuda_stream)
    cv = ctx.curve
    L = cv.fq_limbs
    imad_peak = ctx.bench_imad_wide(8192)
    hbm = 6551.4
    pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        hbm = json.load(open(pk)).get("hbm_gbs", hbm)
    cref = None
    threads = os.cpu_count() or 1
    if args.cpu_max:
        from oracle import cref as _cref          # checker / CPU baseline only
        cref = _cref
    rows = []
    rng = np.random.default_rng(0x48454B41)

    def emit(row):
        rows.append(row)
        print(json.dumps(row), flush=True)

    # ---------------------------------------------------------------- MSM
    for group in (1, 2):
        if f"msm_g{group}" not in ops:
            continue
        enc = codec.g1_affine_to_limbs if group == 1 else codec.g2_affine_to_limbs
        gen_pt = cv.g1 if group == 1 else cv.g2
        gen = enc([gen_pt], cv)[0]
        n32 = 2 * L
        mul_macs = 2 * n32 * n32 + n32
        sqr_macs = n32 * (n32 + 1) // 2 + n32 * n32 + n32
        # mixed XYZZ addition: 8M + 2S in Fq (dedicated squaring); in Fq2 8 Karatsuba multiplies (3 Fq mul) + 2 squares (2 Fq mul)
        add_macs = (8 * mul_macs + 2 * sqr_macs) if group == 1 else 28 * mul_macs
        for lg in range(args.lo, (args.hi if group == 1 else min(args.hi, args.g2_hi)) + 1):
            n = 1 << lg
            ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
            ks[:, 3] >>= 2
            h = ctx.bases_from_scalars(group, gen, ks)
            uni = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
            uni[:, 3] >>= 2
            wit = uni.copy()
            u = rng.random(n)
            wit[u >= 0.1] = 0
            wit[(u >= 0.1) & (u < 0.55), 0] = 1
            exp = {}
            for kind, sc in (("uniform", uni), ("witness", wit)):
                e = limb_dot_mod(ks, sc, cv.r)
                he = ctx.bases_from_scalars(group, gen, codec.ints_to_limbs([e], 4))
                exp[kind] = ctx.bases_download(he, group)[0]
                ctx.bases_free(he)
            d_out = torch.zeros(128, dtype=torch.int64, device="cuda")
            cpu_ms = {}
            if cref is not None and lg <= args.cpu_max:
                pts = ctx.bases_download(h, group)
                for kind, sc in (("uniform", uni), ("witness", wit)):
                    t0 = time.time()
                    cref.msm(cv.cid, group, pts, None, sc, 0, threads=threads)
                    cpu_ms[kind] = (time.time() - t0) * 1e3
                del pts
            for mode in ("plain", "table"):
                for kind, sc, dens in (("uniform", uni, 1.0), ("witness", wit, 0.1)):
                    if mode == "table":
                        # the table geometry follows the expected scalar density, as hk_pk_create does
                        ctx.bases_free(h)
                        h = ctx.bases_from_scalars(group, gen, ks)
                        ctx.bases_precompute(h, dens)
                    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
                    ms = best_ms(stream, lambda: ctx.msm(group, h, d_sc.data_ptr(), n=n, out=d_out.data_ptr()))
                    jac = d_out.cpu().numpy().view(np.uint64)[: 3 * (L if group == 1 else 2 * L)].reshape(1, -1)
                    aff, inf = ctx.into_affine(group, jac)
                    ok = bool(np.array_equal(aff[0], exp[kind]) and not inf[0])
                    row = {"op": f"msm_g{group}", "curve": args.curve, "log_n": lg, "mode": mode, "scalars": kind,
                           "ms": round(ms, 4), "Mpts_per_s": round(n / ms / 1e3, 2), "exact": ok}
                    if kind == "uniform":
                        # canonical algorithmic work: c = 16 -> 16 windows, 10 Fq-mul per mixed addition
                        macs = n * 16 * add_macs
                        row["frac_of_imad_peak_c16"] = round(macs / (ms * 1e-3) / imad_peak, 4)
                    if kind in cpu_ms:
                        row["cpu_ms"] = round(cpu_ms[kind], 1)
                        row["cpu_threads"] = threads
                        row["speedup_vs_cpu"] = round(cpu_ms[kind] / ms, 1)
                    emit(row)
                    del d_sc
            ctx.bases_free(h)

    # ---------------------------------------------------------------- NTT
    one = codec.to_mont([1], cv.r, 4)[0]
    fr_macs = 2 * 8 * 8 + 8
    for lg in (range(args.lo, args.hi + 1) if "ntt" in ops else ()):
        N = 1 << lg
        a = rng.integers(0, 1 << 60, size=(N, 4), dtype=np.uint64)      # raw Montgomery residues < r
        d = torch.from_numpy(a.view(np.int64)).cuda().view(-1)
        res = {}
        for name, kw in (("forward", {}), ("inverse", {"inverse": True}), ("coset_forward", {"coset": True}),
                         ("coset_inverse", {"inverse": True, "coset": True})):
            res[name] = best_ms(stream, lambda: ctx.ntt(d.data_ptr(), lg, **kw))
        # checks on a fresh copy: out[0] = sum x_j (the transform is linear, so this holds on Montgomery
        # residues as well), round trip, delta known answer
        d.copy_(torch.from_numpy(a.view(np.int64)).view(-1))
        ctx.ntt(d.data_ptr(), lg)
        ctx.sync()
        out0 = codec.limbs_to_ints(d.view(-1)[:4].cpu().numpy().view(np.uint64).reshape(1, 4))[0]
        ok = out0 == limb_sum_mod(a, cv.r)
        ctx.ntt(d.data_ptr(), lg, inverse=True)
        ctx.sync()
        ok = ok and bool(np.array_equal(d.cpu().numpy().view(np.uint64).reshape(N, 4), a))
        dl = np.zeros((N, 4), dtype=np.uint64)
        dl[1] = one
        dd = torch.from_numpy(dl.view(np.int64)).cuda()
        ctx.ntt(dd.data_ptr(), lg)
        ctx.sync()
        w = pow(cv.fr_generator, (cv.r - 1) >> lg, cv.r)       # = TWO_ADIC_ROOT_OF_UNITY^(2^(adicity - lg))
        idx = [0, 1, 2, N // 2, N - 1]
        got = codec.from_mont(dd.cpu().numpy().view(np.uint64).reshape(N, 4)[idx], cv.r)
        ok = ok and got == [pow(w, k, cv.r) for k in idx]
        cpu_ms = None
        elementwise = None
        if cref is not None and lg <= args.cpu_max:
            b = a.copy()
            t0 = time.time()
            cref.ntt(cv.cid, b, lg, threads=threads)
            cpu_ms = (time.time() - t0) * 1e3
            # the CPU column's output is the oracle's forward transform of the same input: compare every element
            d.copy_(torch.from_numpy(a.view(np.int64)).view(-1))
            ctx.ntt(d.data_ptr(), lg)
            ctx.sync()
            elementwise = bool(np.array_equal(d.cpu().numpy().view(np.uint64).reshape(N, 4), b))
            ok = ok and elementwise
        ms = res["forward"]
        row = {"op": "ntt", "curve": args.curve, "log_n": lg, "ms": {k: round(v, 4) for k, v in res.items()},
               "GBps_algorithmic": round(64 * N / (ms * 1e-3) / 1e9, 1),
               "frac_of_hbm_peak": round(64 * N / (ms * 1e-3) / 1e9 / hbm, 4),
               "Gbutterflies_per_s": round(N / 2 * lg / ms / 1e6, 2),
               "frac_of_imad_peak": round(N / 2 * lg * fr_macs / (ms * 1e-3) / imad_peak, 4), "checked": bool(ok)}
        if cpu_ms is not None:
            row.update(cpu_ms=round(cpu_ms, 1), cpu_threads=threads, speedup_vs_cpu=round(cpu_ms / ms, 1),
                       elementwise_vs_oracle=elementwise)
        emit(row)
        del d, dd
    summary = {"imad_peak_T_mac_s": imad_peak / 1e12, "hbm_peak_GBps": hbm, "all_checked":
               all(r.get("exact", r.get("checked")) for r in rows), "rows": rows}
    if args.out:
        json.dump(summary, open(args.out, "w"), indent=1)
    ctx.close()
    return 0 if summary["all_checked"] else 1


if __name__ == "__main__":
    sys.exit(main())
