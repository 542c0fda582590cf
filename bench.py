#!/usr/bin/env python3
"""Benchmark of the Hekaton per-subcircuit CP-Groth16 hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

Metric (BASELINE.json): subcircuit proofs/s.  A "step" proves one batch of synthetic subcircuits
(stage-0 commitment + stage-1 proof each, as the reference's workers do: mpi-snark/src/worker.rs:48-85)
of the big-merkle shape (BASELINE config 3: 64 subcircuits x 4 SHA2 iters, 8 portals ->
n = 200 000 constraints, domain N = 2^18, BLS12-381), `--batch` subcircuits per GPU per step;
subcircuits are independent, so ranks shard them with no data-path collective and only gather
the finished proofs (weak scaling).  One JSON line is printed by rank 0:
  value  proofs/s with every input already resident in HBM (hk_prove, HK_DEVICE)
  e2e    proofs/s through the host-buffer C-ABI calls (hk_commit_batch / hk_prove_batch,
         HK_HOST_PINNED_ASYNC): H2D of the assignments from pinned host memory and D2H of the commitments
         and proofs inside the timed region
  roofline      dominant kernel (G1 bucket accumulation), integer-multiply bound
  roofline_ntt  NTT passes against measured HBM bandwidth (north_star's figure)
  cpu_baseline  the oracle's C restatement of the reference prover (arkworks-0.4 algorithms) on
                the host cores, on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="big-merkle")
    ap.add_argument("--batch", type=int, default=8, help="subcircuits per GPU per step")
    ap.add_argument("--curve", default="bls12_381")
    ap.add_argument("--streams", type=int, default=1,
                    help="concurrent sub-batches per GPU (forked contexts: shared key, own stream + scratch); "
                         "each sub-batch of batch/streams proofs is ONE hk_prove_batch pipeline")
    ap.add_argument("--inflight", type=int, default=2,
                    help="batches in flight per GPU: step i runs on pipeline i %% inflight (own contexts, streams and "
                         "scratch), so the latency-bound tail of one batch overlaps the head of the next")
    ap.add_argument("--opt", action="append", default=[],
                    help="MSM tuning knob for A/B runs, name=value with name in waves|smin|fold|chunk|gsort "
                         "(hk_ctx_set_option; results never depend on them)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the standalone MSM / NTT measurements")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE config 2 instead of the proof bench: checked G1/G2 MSM + Fr NTT sweep 2^16..2^24 "
                         "(tools/sweep.py; one JSON line per point)")
    ap.add_argument("--sweep-args", default="", help="extra arguments for tools/sweep.py")
    ap.add_argument("--timeline-host", action="store_true",
                    help="with --timeline: profile the end-to-end (host-buffer) path instead of the device-resident one")
    ap.add_argument("--timeline", default=None,
                    help="diagnostic: write the kernel timeline (name, stream, start, duration) of ONE warm step to "
                         "this file (torch.profiler / CUPTI) and exit; prints no bench line")
    return ap.parse_args()


# dram__bytes_read.sum + dram__bytes_write.sum of k_msm_accum<G1> per mixed addition, from the committed
# `ncu --set full` capture (profiles/r01f_accum_setfull_raw.csv: 3.379 GB read + 0.196 GB written by a launch of
# 2^20 x 16 = 16 777 216 additions)
NCU_ACCUM_DRAM_BYTES_PER_ADD = (3.379023e9 + 0.195592e9) / 16777216


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d
    return 6650.0, "fallback", {}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.split(",") for ln in open(self.f.name).read().strip().splitlines() if ln.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(wl, batch, rank):
    """Per-proof (z, r, s, kappa) for this rank's shard, Montgomery limbs."""
    import random
    from hekaton_system_b200 import codec
    cv = wl.curve
    zs = [codec.to_mont(z, cv.r, 4) for z in wl.assignments]
    rng = random.Random(1000 + rank)
    items = []
    for i in range(batch):
        k = (rank * batch + i) % len(zs)
        r_, s_, kap = rng.randrange(1, cv.r), rng.randrange(1, cv.r), rng.randrange(1, cv.r)
        items.append(dict(k=k, z=zs[k], r=r_, s=s_, kappa=kap,
                          rm=codec.to_mont([r_], cv.r, 4)[0], sm=codec.to_mont([s_], cv.r, 4)[0],
                          km=codec.to_mont([kap], cv.r, 4)))
    return items


def cpu_prover_for(ctx, wl, pk):
    """Host copy of the proving key for the CPU restatement (same bases as the device)."""
    from oracle import cref
    h = pk.handles
    dl = lambda hd, g: ctx.bases_download(hd, g)
    return cref.CpuProver(ctx.curve.cid, wl.n, wl.l, wl.nw, wl.csr, wl.stage_start, dl(h["a_g"], 1), dl(h["b_g"], 1),
                          dl(h["b_h"], 2), dl(h["h_g"], 1), [dl(c, 1) for c in h["ck"]], dl(h["deltas_g"], 1),
                          pk.points["alpha_g"], pk.points["beta_g"], pk.points["beta_h"], pk.points["last_delta_h"])


def measure_extras(ctx, stream, imad_peak, hbm_gbs):
    """Standalone G1 MSM at 2^20 (uniform 255-bit scalars, resident bases with window table) and Fr NTT
    at 2^20 / 2^24, device-resident, CUDA events on the context stream, best of 5 after warm-up."""
    import torch
    from hekaton_system_b200 import codec
    cv = ctx.curve
    L = cv.fq_limbs
    rng = np.random.default_rng(7)
    out = {}

    def best_ms(fn, reps=5):
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

    n = 1 << 20
    ks = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    ks[:, 3] >>= 2
    h = ctx.bases_from_scalars(1, codec.g1_affine_to_limbs([cv.g1], cv)[0], ks)
    ctx.bases_precompute(h, 1.0)
    sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    sc[:, 3] >>= 2
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(64, dtype=torch.int64, device="cuda")
    ms = best_ms(lambda: ctx.msm(1, h, d_sc.data_ptr(), n=n, out=d_out.data_ptr()))
    n32 = 2 * L
    W, c = 16, 16                           # planner's choice at this size (msm_plan, table mode)
    macs = n * W * (8 * (2 * n32 * n32 + n32) + 2 * (n32 * (n32 + 1) // 2 + n32 * n32 + n32))
    out["msm_g1_2p20"] = {"ms": ms, "Mpts_per_s": n / ms / 1e3, "window_bits": c, "windows": W,
                          "algorithmic_wide_macs": macs, "frac_of_imad_peak": macs / (ms * 1e-3) / imad_peak,
                          "note": "whole MSM (sort + accumulate + reduce), scalars and bases resident in HBM"}
    ctx.bases_free(h)
    del d_sc
    for lg in (20, 24):
        N = 1 << lg
        a = rng.integers(0, 1 << 60, size=(N, 4), dtype=np.uint64)
        d = torch.from_numpy(a.view(np.int64)).cuda()
        ms = best_ms(lambda: ctx.ntt(d.data_ptr(), lg))
        gbs = 64 * N / (ms * 1e-3) / 1e9
        out[f"ntt_fwd_2p{lg}"] = {"ms": ms, "GBps_algorithmic": gbs, "frac_of_hbm_peak": gbs / hbm_gbs,
                                  "Gbutterflies_per_s": N / 2 * lg / ms / 1e6}
        del d
    return out


def emit(line):
    """The one JSON line goes to the REAL stdout; everything else (library banners such as NCCL's version
    line, build output) was redirected to stderr at start-up."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    if not args.sweep:
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)              # fd 1 -> stderr for the rest of the process (C libraries included)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return 0          # the CPU arm runs on rank 0 only; other ranks exit without work

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    if args.sweep:
        if rank != 0:
            return 0
        import __graft_entry__ as entry
        entry.build()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sweep
        return sweep.main(["--curve", args.curve] + args.sweep_args.split())
    torch.cuda.set_device(local)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1 and args.impl == "ours":
        dist.barrier()

    from hekaton_system_b200 import codec, workload
    from hekaton_system_b200.backend import Context
    from hekaton_system_b200.shard import gather_records, shard_range

    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx = Context(args.curve, device=local, stream=stream.cuda_stream)
    if args.opt:
        from hekaton_system_b200 import backend as _B
        names = {"waves": _B.OPT_ACCUM_WAVES, "smin": _B.OPT_ACCUM_SEG_MIN, "fold": _B.OPT_FOLD_SEG,
                 "chunk": _B.OPT_REDUCE_CHUNK, "gsort": _B.OPT_SORT_GLOBAL_ATOMICS}
        for kv in args.opt:
            k_, v_ = kv.split("=")
            ctx.set_option(names[k_], int(v_))         # forks inherit the settings
    nstreams = max(1, min(args.streams, args.batch))
    inflight = max(1, args.inflight)
    # pipelines: each has, per sub-batch, a proof context and a commitment context (forks of `ctx`: shared
    # keys, own stream + scratch).  Stage-0 commitments run on their own context so the small latency-bound
    # commitment MSM overlaps the head of the stage-1 proof pipeline.
    pipes = []
    for _ in range(inflight):
        ps = [torch.cuda.Stream(device=local) for _ in range(nstreams)]
        cs = [torch.cuda.Stream(device=local) for _ in range(nstreams)]
        pipes.append(dict(pstreams=ps, cstreams=cs, ctxs=[ctx.fork(stream=s_.cuda_stream) for s_ in ps],
                          cctxs=[ctx.fork(stream=s_.cuda_stream) for s_ in cs]))
    all_ctxs = [c_ for pp in pipes for c_ in pp["ctxs"] + pp["cctxs"]]
    all_streams = [s_ for pp in pipes for s_ in pp["pstreams"] + pp["cstreams"]]
    cv = ctx.curve
    t0 = time.time()
    wl = workload.make_workload(cv, args.shape, seed=1, num_assignments=2)
    workload.key_scalars(wl, seed=7)
    pk = workload.build_device_key(ctx, wl)
    setup_s = time.time() - t0
    items = build_inputs(wl, args.batch, rank)
    L = cv.fq_limbs
    P1, P2 = 2 * L * 8, 4 * L * 8
    config = {"workload": f"{args.shape}: synthetic 2-stage R1CS, n={wl.n} constraints, m={wl.m} variables, "
                          f"N=2^{wl.log_n}, {args.curve}, bit-heavy witnesses",
              "subcircuits_per_gpu_per_step": args.batch, "global_batch": args.batch * world,
              "parallelism": f"subcircuits sharded over {world} GPU(s), no data-path collective; "
                             f"{inflight} batch(es) in flight per GPU, each {nstreams} sub-batch(es) of "
                             f"{args.batch // nstreams} proofs (hk_commit_batch + hk_prove_batch on forked contexts)",
              "l2": "working set (resident bases ~150 MB + MSM workspaces) exceeds the 126 MB L2",
              "setup_s": round(setup_s, 1)}

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        from oracle import cref
        cp = cpu_prover_for(ctx, wl, pk)
        threads = os.cpu_count() or 1
        it = items[0]
        for _ in range(min(args.warmup, 1)):
            cp.commit(0, it["z"], it["km"], threads=threads)
            cp.prove(it["z"], it["rm"], it["sm"], it["km"], threads=threads)
        t0 = time.time()
        for i in range(args.steps):
            it = items[i % len(items)]
            cp.commit(0, it["z"], it["km"], threads=threads)                    # stage 0 (committer.rs:55-98)
            cp.prove(it["z"], it["rm"], it["sm"], it["km"], threads=threads)    # stage 1 (committer.rs:100-122)
        dt = time.time() - t0
        v = args.steps / dt
        sample = (f"{args.steps} subcircuit(s) (commit + prove) of the same workload, one per step "
                  f"(bounded sample of the {args.batch}-subcircuit step)")
        emit(({
            "impl": "reference", "metric": "subcircuit proofs/s", "value": v, "unit": "proofs/s", "n_gpus": 0,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Montgomery, 64-bit limbs)",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": threads, "kind": "port", "sample": sample,
                             "note": "C restatement of the arkworks-0.4 algorithms the reference runs (window-parallel "
                                     "Pippenger, radix-2 FFT); not the arkworks binary (no Rust toolchain)"},
            "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        ctx.close()
        return 0

    # ------------------------------------------------------------------ device-resident inputs
    dev = torch.device("cuda", local)
    assert args.batch % nstreams == 0, "--batch must be a multiple of --streams"
    per = args.batch // nstreams
    OUTB = 2 * P1 + P2 + 3
    COMB = P1 + 1
    chunks = []          # per forked context: contiguous device + host inputs for its sub-batch
    for t in range(nstreams):
        its = items[t * per:(t + 1) * per]
        # e2e inputs live in pinned host memory (numpy views of torch pinned tensors)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).pin_memory()
        pins = [pin(np.stack([it[k] for it in its])) for k in ("z", "rm", "sm", "km")]
        hz, hr, hs, hk_ = [t_.numpy().view(np.uint64) for t_ in pins]
        hz = hz.reshape(len(its), -1, 4)
        dz = torch.from_numpy(hz.view(np.int64)).to(dev)
        drsk = torch.from_numpy(np.concatenate([hr.ravel(), hs.ravel(), hk_.ravel()]).view(np.int64)).to(dev)
        # one output record block per step of a run: the proofs of the whole job are gathered ONCE at its end
        nout = max(args.steps, args.warmup, inflight, 4)
        outs = [torch.zeros((per * (OUTB + COMB) + 63) // 8 + 4, dtype=torch.int64, device=dev) for _ in range(nout)]
        chunks.append(dict(hz=hz, hr=hr, hs=hs, hk=hk_, dz=dz, drsk=drsk, outs=outs, pins=pins))
    my_subcircuits = shard_range(args.batch * world, world, rank)      # contiguous chunk per rank (node.rs:471-493)
    assert len(my_subcircuits) == args.batch

    def step_device(i):
        # asynchronous: everything is enqueued on the pipeline's own streams (in order per context)
        pp = pipes[i % inflight]
        for t, ch in enumerate(chunks):
            q = ch["drsk"].data_ptr()
            p = ch["outs"][i % len(ch["outs"])].data_ptr()
            pc = p + ((per * OUTB + 15) // 16) * 16
            # stage 0: commitments (committer.rs:55-98); stage 1: proofs (committer.rs:100-122)
            pp["cctxs"][t].commit_batch_device(pk.handle, 0, per, ch["dz"].data_ptr(), wl.m, q + 64 * per, pc, pc + per * P1)
            pp["ctxs"][t].prove_batch_device(pk.handle, per, ch["dz"].data_ptr(), q, q + 32 * per, q + 64 * per,
                                             p, p + per * P1, p + per * (P1 + P2), p + per * (2 * P1 + P2))

    import threading

    def run_device(steps):
        for s_ in all_streams:
            s_.wait_stream(stream)
        for i in range(steps):
            step_device(i)
        for s_ in all_streams:
            stream.wait_stream(s_)
        if world > 1:
            # the path's one exchange step (node.rs:687-715 gather_responses: once per stage, after every worker
            # has finished its chunk): the proof records of the whole run go to every rank
            with torch.cuda.stream(stream):
                recs = [ch["outs"][i % len(ch["outs"])] for i in range(min(steps, len(chunks[0]["outs"]))) for ch in chunks]
                gather_records(torch.cat(recs).view(1, -1), world)

    # pinned output blocks for the end-to-end path, one per step of a run
    def pinned_u64(*shape):
        return torch.zeros(shape, dtype=torch.int64).pin_memory().numpy().view(np.uint64)

    nhost = max(args.steps, 2 * inflight, 4)
    for ch in chunks:
        ch["hout"] = [dict(a=pinned_u64(per, 2 * L), b=pinned_u64(per, 4 * L), c=pinned_u64(per, 2 * L),
                           inf=torch.zeros((per, 3), dtype=torch.uint8).pin_memory().numpy(),
                           com=pinned_u64(per, 2 * L), cinf=torch.zeros(per, dtype=torch.uint8).pin_memory().numpy())
                      for _ in range(nhost)]

    def run_host(steps):
        # End to end through the host-buffer C ABI (HK_HOST_PINNED_ASYNC): every step's assignments are DMA'd from
        # pinned host memory and its commitments / proofs are DMA'd back into pinned host memory, all inside the timed
        # region.  The calls only enqueue, so the upload of a batch overlaps the kernels of the previous one and
        # `inflight` batches overlap like in the device-resident run; one sync per context at the end.
        for s_ in all_streams:
            s_.wait_stream(stream)
        for i in range(steps):
            pp = pipes[i % inflight]
            for t, ch in enumerate(chunks):
                o = ch["hout"][i % nhost]
                pp["cctxs"][t].commit_batch_async(pk.handle, 0, ch["hz"], ch["hk"], o["com"], o["cinf"])
                pp["ctxs"][t].prove_batch_async(pk.handle, ch["hz"], ch["hr"].reshape(per, 4), ch["hs"].reshape(per, 4),
                                                ch["hk"].reshape(per, -1, 4), o["a"], o["b"], o["c"], o["inf"])
        for c_ in all_ctxs:
            c_.sync()
        for s_ in all_streams:
            stream.wait_stream(s_)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn(steps)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # untimed parity spot check of the full-size workload: device proof == trapdoor known answer
    it0 = items[0]
    a, b, c, inf = ctx.prove(pk.handle, it0["z"], it0["rm"], it0["sm"], it0["km"])
    exp = workload.expected_exponents(wl, wl.assignments[it0["k"]], it0["r"], it0["s"], [it0["kappa"]])
    td = wl.trapdoor
    gen1 = codec.g1_affine_to_limbs([cv.g1], cv)[0]
    gen2 = codec.g2_affine_to_limbs([cv.g2], cv)[0]
    h1 = ctx.bases_from_scalars(1, gen1, codec.ints_to_limbs([exp["A"] * td["g1"] % cv.r, exp["C"] * td["g1"] % cv.r], 4))
    h2 = ctx.bases_from_scalars(2, gen2, codec.ints_to_limbs([exp["B"] * td["g2"] % cv.r], 4))
    e1p, e2p = ctx.bases_download(h1, 1), ctx.bases_download(h2, 2)
    verified = bool(np.array_equal(e1p[0], a) and np.array_equal(e1p[1], c) and np.array_equal(e2p[0], b))
    ctx.bases_free(h1)
    ctx.bases_free(h2)

    imad_peak = ctx.bench_imad_wide(8192)          # wide multiply-accumulates / s (roofline denominator)

    run_device(args.warmup)
    torch.cuda.synchronize()
    if args.timeline:
        from torch.profiler import ProfilerActivity, profile
        nrun = 3 * inflight                       # steady state: the middle third is what tools/timeline_stats.py reads
        if args.timeline_host:
            run_host(2 * inflight)
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            if args.timeline_host:
                run_host(nrun)
            else:
                run_device(nrun)
            torch.cuda.synchronize()
        tmp = args.timeline + ".chrome.json"
        tp.export_chrome_trace(tmp)
        tr = json.load(open(tmp))
        os.unlink(tmp)
        evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        t0_ = min(e["ts"] for e in evs)
        sids = {s_: i for i, s_ in enumerate(sorted({e["args"].get("stream") for e in evs}))}
        with open(args.timeline, "w") as f:
            f.write(f"# {nrun} steps of {args.batch} proofs, {inflight} in flight\n")
            f.write("# start_us dur_us stream grid block regs kernel\n")
            for e in sorted(evs, key=lambda e: e["ts"]):
                a_ = e["args"]
                g_ = "x".join(str(v) for v in a_.get("grid", [])) or "-"
                b_ = "x".join(str(v) for v in a_.get("block", [])) or "-"
                f.write(f"{e['ts'] - t0_:10.1f} {e['dur']:9.1f} s{sids[a_.get('stream')]} {g_} {b_} "
                        f"{a_.get('registers per thread', '-')} {e['name'][:100]}\n")
        ctx.close()
        return 0
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = sum(c_.launches for c_ in all_ctxs)
    ms_dev = timed(run_device, args.steps)
    launches = sum(c_.launches for c_ in all_ctxs) - l0
    # Roofline pass: the same step once more with the profiling events on, (a) as it runs in the pipeline
    # (launches of different lanes / batches overlap and share the GPU) and (b) with the lanes of ONE batch
    # serialised, so that every launch of the dominant kernels is timed alone.  Outside the timed region:
    # the event records and the device-counter read-backs perturb the pipeline.
    def profiled(serial, steps):
        cs = [pipes[0]["ctxs"][0], pipes[0]["cctxs"][0]] if serial else all_ctxs
        for c_ in cs:
            c_.profile_reset()
            c_.profile(True, serial=serial)
        if serial:
            for s_ in all_streams:
                s_.wait_stream(stream)
            q, p_ = chunks[0]["drsk"].data_ptr(), chunks[0]["outs"][0].data_ptr()
            pc_ = p_ + ((per * OUTB + 15) // 16) * 16
            for _ in range(steps):
                cs[1].commit_batch_device(pk.handle, 0, per, chunks[0]["dz"].data_ptr(), wl.m, q + 64 * per, pc_, pc_ + per * P1)
                torch.cuda.synchronize()
                cs[0].prove_batch_device(pk.handle, per, chunks[0]["dz"].data_ptr(), q, q + 32 * per, q + 64 * per,
                                         p_, p_ + per * P1, p_ + per * (P1 + P2), p_ + per * (2 * P1 + P2))
                torch.cuda.synchronize()
            ms = None
        else:
            ms = timed(run_device, steps)
        out = {}
        for k in (0, 1, 2):
            rs = [c_.profile_read(k) for c_ in cs]
            out[k] = tuple(sum(x[j] for x in rs) for j in range(3))
        for c_ in cs:
            c_.profile(False)
            c_.profile_reset()
        return out, ms

    prof, ms_prof = profiled(False, max(2, min(args.steps, 4)))
    prof_alone, _ = profiled(True, 2)
    # end-to-end through the host-buffer C-ABI call
    run_host(min(args.warmup, 2) * inflight)
    ms_e2e = timed(run_host, args.steps)
    clocks = sampler.stop() if sampler else None

    # secondary headline figures (BASELINE.json metric: "G1 MSM Mpts/s at 2^20"; north_star: MSM / NTT
    # throughput with roofline fractions), measured per GPU on rank 0 after the timed regions
    extras = None
    if rank == 0 and not args.no_extras:
        extras = measure_extras(ctx, stream, imad_peak, peaks()[0])

    proofs = args.batch * world * args.steps
    value = proofs / (ms_dev * 1e-3)
    e2e_v = proofs / (ms_e2e * 1e-3)

    if rank == 0:
        hbm_gbs, hbm_src, _ = peaks()
        # roofline of the dominant kernel: G1 bucket accumulation.
        # algorithmic work per unit (one mixed addition XYZZ += affine): 8 Fq multiplications of 2 N^2 + N wide
        # MACs and 2 squarings of N(N+1)/2 + N^2 + N, N = 2 * fq_limbs 32-bit limbs (BLS12-381: 8*300 + 2*234 = 2868;
        # SURVEY 8d counts 10 * 300 = 3000 for the same addition -- the kernel executes the smaller number)
        n32 = 2 * L
        mul_macs = 2 * n32 * n32 + n32                          # Montgomery multiply
        sqr_macs = n32 * (n32 + 1) // 2 + n32 * n32 + n32       # dedicated squaring (field_gen: *_sqr)
        macs_per_add = 8 * mul_macs + 2 * sqr_macs              # XYZZ mixed addition: 8M + 2S
        # what this build executes: with the fused dual product (hk_build_flags & HK_BUILD_MUL2, BLS12-381 Fq only)
        # Y3 = R (Q - X3) - Y1 PPP is ONE 3 N^2 + N pass instead of two multiplies: 6M + 2S + 1 dual = 2712
        from hekaton_system_b200 import backend as hk_backend
        flags = hk_backend.load_library().hk_build_flags()
        dual = bool(flags & 1) and args.curve == "bls12_381"
        exec_macs = (6 * mul_macs + 2 * sqr_macs + 3 * n32 * n32 + n32) if dual else macs_per_add
        ms_acc, n_acc, adds = prof_alone[0]
        ms_acc_p, n_acc_p, adds_p = prof[0]
        roof = None
        if n_acc:
            ach = adds * macs_per_add / (ms_acc * 1e-3)
            roof = {"kernel": "k_msm_accum<G1>", "bound": "imad", "achieved": ach / 1e12, "peak": imad_peak / 1e12,
                    "unit": "T wide-MAC/s (32x32+64 IMAD.WIDE)", "frac": ach / imad_peak,
                    "traffic": NCU_ACCUM_DRAM_BYTES_PER_ADD * adds / n_acc,
                    "traffic_note": f"dram read+write bytes per launch = {NCU_ACCUM_DRAM_BYTES_PER_ADD:.0f} B per addition "
                                    "(ncu --set full capture of this kernel, profiles/r01f_accum_setfull_raw.csv: 3.57 GB for a launch of "
                                    "16.8 M additions) x this run's additions per launch; algorithmic = 104 B per addition "
                                    "(one 96-B affine base + one 8-B sorted pair)",
                    "launches": n_acc, "avg_launch_ms": ms_acc / n_acc, "units_per_launch": adds / n_acc,
                    "algorithmic_macs_per_unit": macs_per_add,
                    "executed_macs_per_unit": exec_macs,
                    "frac_executed": ach / imad_peak * exec_macs / macs_per_add,
                    "share_of_step": (ms_acc_p / ms_prof) if ms_prof else None,
                    "in_pipeline": {"frac": (adds_p * macs_per_add / (ms_acc_p * 1e-3) / imad_peak) if n_acc_p else None,
                                    "launches": n_acc_p, "avg_launch_ms": (ms_acc_p / n_acc_p) if n_acc_p else None},
                    "note": "frac: algorithmic MACs (8M + 2S per mixed addition, SURVEY 8d) / time / peak; frac_executed: "
                            "the multiplies this build issues (fused dual product: one reduction fewer per addition) / "
                            "time / peak = multiplier-pipe utilisation.  CUDA events around every launch with the lanes of a batch serialised (each launch "
                            "alone on the GPU, same kernels and inputs, right after the timed region); in_pipeline: the "
                            "same events while lanes and batches overlap, so a launch's duration includes the time it "
                            "shares the SMs with other kernels; share_of_step = summed in-pipeline launch time / wall time",
                    "peak_source": "measured in this run by hk_bench_imad_wide (MEASURED_PEAKS.json has no integer peak)"}
        ms_ntt, n_ntt, elems = prof_alone[2]
        ms_ntt_p, n_ntt_p, elems_p = prof[2]
        roof_ntt = None
        if n_ntt:
            gbs = elems * 64 / (ms_ntt * 1e-3) / 1e9
            roof_ntt = {"kernel": "k_ntt_pass", "bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                        "frac": gbs / hbm_gbs, "traffic": None, "launches": n_ntt, "avg_launch_ms": ms_ntt / n_ntt,
                        "share_of_step": (ms_ntt_p / ms_prof) if ms_prof else None,
                        "peak_source": f"{hbm_src} (MEASURED_PEAKS.json hbm_gbs)",
                        "note": "64 B per element per pass (read + write once); launches timed alone (lanes serialised). "
                                "A pass is multiplier-bound, not HBM-bound (136 wide MACs per butterfly): see extras for "
                                "the IMAD fraction"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cp = cpu_prover_for(ctx, wl, pk)
                threads = os.cpu_count() or 1
                t0 = time.time()
                cp.commit(0, it0["z"], it0["km"], threads=threads)
                ca, cb, cc = cp.prove(it0["z"], it0["rm"], it0["sm"], it0["km"], threads=threads)
                dt = time.time() - t0
                same = bool(np.array_equal(ca, a) and np.array_equal(cb, b) and np.array_equal(cc, c))
                cpu = {"value": 1.0 / dt, "unit": "proofs/s", "cores": threads, "kind": "port",
                       "sample": "1 subcircuit (commit + prove) of the same workload (same key, same assignment as step 0)",
                       "seconds": dt, "bit_exact_with_gpu": same}
            except Exception as e:       # the baseline is a report, never a dependency of the product path
                cpu = {"value": None, "unit": "proofs/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": "subcircuit proofs/s", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32x2 (Montgomery Fq/Fr on 32x32+64 IMAD.WIDE lanes)", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": "proofs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": args.batch * (wl.m * 32 + 96), "d2h_bytes_per_step": args.batch * (OUTB + COMB)},
            "gpu_launches": launches, "roofline": roof, "roofline_ntt": roof_ntt, "cpu_baseline": cpu, "extras": extras,
            "verified_vs_trapdoor": verified,
        }
        emit(line)
    for c_ in all_ctxs:
        c_.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
