#!/usr/bin/env python3
"""One-shot hardware probe + timing sweep (run under gpurun; prints JSON lines)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from hekaton_system_b200 import backend as B, codec
from hekaton_system_b200.backend import Context


def ev_time(fn, stream, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    what = sys.argv[1:] or ["imad", "mul", "msm", "ntt"]
    torch.cuda.init()
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "sms": torch.cuda.get_device_properties(0).multi_processor_count}))
    for cname in ("bls12_381",):
        ctx = Context(cname, device=0, stream=stream.cuda_stream)
        cv = ctx.curve
        if "imad" in what:
            for it in (2048, 8192):
                print(json.dumps({"probe": "imad_wide_mac_per_s", "iters": it, "value": ctx.bench_imad_wide(it)}))
        if "mul" in what:
            for field, nm, wide in ((B.FIELD_FQ, "fq", 288 + 12), (B.FIELD_FR, "fr", 128 + 8)):
                for thr in (148 * 256, 148 * 1024, 148 * 2048):
                    iters = 4000
                    ms = ctx.bench_field_mul(field, thr, iters)
                    muls = thr * iters
                    print(json.dumps({"probe": f"{nm}_mul", "threads": thr, "ms": ms, "Gmul_per_s": muls / ms / 1e6,
                                      "wide_mac_per_s": muls * wide / (ms * 1e-3)}))
        if "msm" in what:
            gen1 = codec.g1_affine_to_limbs([cv.g1], cv)[0]
            gen2 = codec.g2_affine_to_limbs([cv.g2], cv)[0]
            rng = np.random.default_rng(1)
            logs1 = tuple(int(x) for x in os.environ.get("HK_PROBE_LOGS1", "12,16,18,20,22").split(",") if x)
            logs2 = tuple(int(x) for x in os.environ.get("HK_PROBE_LOGS2", "16,18,20").split(",") if x)
            for group, gen, logs in ((1, gen1, logs1), (2, gen2, logs2)):
                for lg in logs:
                    n = 1 << lg
                    ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
                    ks[:, 3] >>= 2          # < 2^253 < r
                    t0 = time.time()
                    h = ctx.bases_from_scalars(group, gen, ks)
                    t_gen = time.time() - t0
                    if os.environ.get("HK_PROBE_TABLE", "1") == "1":
                        ctx.bases_precompute(h)
                    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
                    sc[:, 3] >>= 2
                    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
                    d_out = torch.zeros(64, dtype=torch.int64, device="cuda")
                    wit = sc.copy()
                    mask = rng.random(n) < 0.9
                    wit[mask] = 0
                    wit[mask, 0] = rng.integers(0, 2, size=int(mask.sum()), dtype=np.uint64)
                    d_wit = torch.from_numpy(wit.view(np.int64)).cuda()
                    for kind, d in (("uniform", d_sc), ("witness90", d_wit)):
                        l0 = ctx.launches
                        best, med = ev_time(lambda: ctx.msm(group, h, d.data_ptr(), n=n, out=d_out.data_ptr()), stream)
                        print(json.dumps({"probe": "msm", "group": group, "log_n": lg, "scalars": kind, "ms_best": best,
                                          "ms_med": med, "Mpts_per_s": n / best / 1e3,
                                          "launches_per_msm": (ctx.launches - l0) // 4, "gen_s": round(t_gen, 3)}))
                    ctx.bases_free(h)
        if "ntt" in what:
            rng = np.random.default_rng(2)
            for lg in (16, 18, 20, 22, 24):
                n = 1 << lg
                a = rng.integers(0, 1 << 60, size=(n, 4), dtype=np.uint64)
                d = torch.from_numpy(a.view(np.int64)).cuda()
                ctx.ntt(d.data_ptr(), lg)     # builds the domain tables
                best, med = ev_time(lambda: ctx.ntt(d.data_ptr(), lg), stream)
                print(json.dumps({"probe": "ntt_fwd", "log_n": lg, "ms_best": best, "GBps_alg": 64 * n / best / 1e6,
                                  "Gbutterfly_per_s": n / 2 * lg / best / 1e6}))
                abc = torch.from_numpy(rng.integers(0, 1 << 60, size=(3 * n, 4), dtype=np.uint64).view(np.int64)).cuda()
                hh = torch.zeros(n * 4, dtype=torch.int64, device="cuda")
                fn = lambda: ctx.qap_h(abc.data_ptr(), abc.data_ptr() + 32 * n, abc.data_ptr() + 64 * n, lg, h=hh.data_ptr())
                best, med = ev_time(fn, stream)
                print(json.dumps({"probe": "qap_h_7ntt", "log_n": lg, "ms_best": best}))
        ctx.close()


if __name__ == "__main__":
    main()
