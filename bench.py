#!/usr/bin/env python3
"""Benchmark of the Hekaton per-subcircuit CP-Groth16 hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

Metric (BASELINE.json): subcircuit proofs/s.  A "step" is one batch of `--batch` synthetic subcircuits per GPU of the
big-merkle shape (BASELINE config 3: 64 subcircuits x 4 SHA2 iters, 8 portals -> n = 200 000 constraints, domain
N = 2^18, BLS12-381); each subcircuit is a stage-0 commitment plus a stage-1 proof, as the reference's workers produce
them (mpi-snark/src/worker.rs:48-85).  A timed run of K steps follows the reference's protocol order
(mpi-snark/src/bin/node.rs:488-539):

    phase A   every stage-0 commitment of the run (hk_commit_batch, K batches)
    barrier   the commitments of ALL ranks are gathered (NCCL all_gather) and read back by the host, which derives
              the stage-1 requests from them (stub of coordinator.rs:315-351; what matters is the dependency:
              no proof starts before every commitment is known)
    phase B   every stage-1 proof of the run (hk_prove_batch, K batches, `--inflight` batches overlapping)
    gather    the proofs of all ranks (NCCL all_gather), for the aggregation step

Subcircuits are independent within a phase, so ranks shard them with no data-path collective (weak scaling).
One JSON line is printed by rank 0:
  value         proofs/s with every input already resident in HBM (HK_DEVICE)
  e2e           proofs/s through the host-buffer C-ABI calls (HK_HOST_PINNED_ASYNC): H2D of the assignments from pinned
                host memory and D2H of the commitments / proofs inside the timed region; e2e.pageable: the same from
                ordinary (pageable) buffers through blocking HK_HOST calls driven by one host thread per pipeline
  roofline      dominant kernel (G1 bucket accumulation), integer-multiply bound
  roofline_ntt  NTT passes against measured HBM bandwidth (north_star's figure)
  cpu_baseline  the oracle's C restatement of the reference prover (arkworks-0.4 algorithms) on the host cores, on a
                bounded sample of the same workload
  extras        standalone G1 MSM 2^20 / NTT figures and `shapes`: the other BASELINE configs (VKD 2^14 x 64,
                big-merkle-small 2^16 x 16, VM 2^21 x 4 per step), a few steps each, trapdoor-verified
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# subcircuits per GPU per step for the shapes of extras.shapes
SHAPE_BATCH = {"big-merkle": 8, "big-merkle-small": 16, "vkd": 64, "vm": 4, "tiny": 8}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="big-merkle")
    ap.add_argument("--batch", type=int, default=0, help="subcircuits per GPU per step (default: 8 for big-merkle)")
    ap.add_argument("--curve", default="bls12_381")
    ap.add_argument("--inflight", type=int, default=2,
                    help="proof batches in flight per GPU: step i runs on pipeline i %% inflight (own contexts, streams and "
                         "scratch), so the latency-bound tail of one batch overlaps the head of the next")
    ap.add_argument("--opt", action="append", default=[],
                    help="MSM tuning knob for A/B runs, name=value with name in waves|smin|fold|chunk|gsort|g2pair|quad|noresolve "
                         "(hk_ctx_set_option; results never depend on them)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the standalone MSM / NTT measurements and the other shapes")
    ap.add_argument("--shapes", default="vkd,big-merkle-small,vm", help="extras.shapes: comma-separated shapes ('' = none)")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE config 2 instead of the proof bench: checked G1/G2 MSM + Fr NTT sweep 2^16..2^24 "
                         "(tools/sweep.py; one JSON line per point)")
    ap.add_argument("--sweep-args", default="", help="extra arguments for tools/sweep.py")
    ap.add_argument("--timeline-host", action="store_true",
                    help="with --timeline: profile the end-to-end (host-buffer) path instead of the device-resident one")
    ap.add_argument("--timeline", default=None,
                    help="diagnostic: write the kernel timeline (name, stream, start, duration) of a few warm steps to "
                         "this file (torch.profiler / CUPTI) and exit; prints no bench line")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured", d
    return 6650.0, "fallback", {}


def ncu_traffic():
    """Per-unit DRAM traffic of the dominant kernels from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the raw CSV pages)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


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

    def wait_first(self, timeout=5.0):
        """Block until nvidia-smi has printed its first sample: its start-up (NVML initialisation on the GPU under test)
        stalls kernel launches for ~100 ms -- seen once as a 43 ms/step outlier when it overlapped the timed region."""
        t0 = time.time()
        while self.p is not None and self.p.poll() is None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.02)

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


def workload_string(shape, wl, curve):
    return (f"{shape}: synthetic 2-stage R1CS, n={wl.n} constraints, m={wl.m} variables, N=2^{wl.log_n}, {curve}, "
            f"bit-heavy witnesses")


def make_config(args, wl, world, batch):
    """The `config` object of BOTH arms (the reference arm is timed "on your arm's config", so it prints the same
    dict): the workload, how the job is laid out over the GPUs, and why no L2 flush is needed."""
    inflight = max(1, args.inflight)
    return {"workload": workload_string(args.shape, wl, args.curve),
            "subcircuits_per_gpu_per_step": batch, "global_batch": batch * world,
            "parallelism": f"subcircuits sharded over {world} GPU(s), no data-path collective; protocol order per run: "
                           f"all stage-0 commitments -> all_gather + host read-back (stage barrier, node.rs:488-539) -> "
                           f"all stage-1 proofs -> all_gather; {inflight} batch(es) in flight per GPU within a phase "
                           f"(hk_commit_batch / hk_prove_batch on forked contexts)",
            "l2": "working set (resident window tables ~3 GB + MSM workspaces) exceeds the 126 MB L2: no flush needed"}


# ------------------------------------------------------------------------------------------------------------
class ShapeBench:
    """One subcircuit shape on one GPU: workload, resident proving key, per-pipeline contexts and buffers, and the
    protocol-ordered runs (device-resident / pinned-async host buffers / pageable host buffers)."""

    def __init__(self, ctx, stream, shape, batch, inflight, max_steps, rank, world, local, num_assignments=2):
        import torch
        from hekaton_system_b200 import workload
        self.torch = torch
        self.ctx, self.stream, self.shape, self.batch = ctx, stream, shape, batch
        self.inflight, self.rank, self.world, self.local = inflight, rank, world, local
        self.dev = torch.device("cuda", local)
        cv = self.cv = ctx.curve
        t0 = time.time()
        self.wl = wl = workload.make_workload(cv, shape, seed=1, num_assignments=num_assignments)
        workload.key_scalars(wl, seed=7)
        self.pk = workload.build_device_key(ctx, wl)
        self.setup_s = time.time() - t0
        self.items = build_inputs(wl, batch, rank)
        L = self.L = cv.fq_limbs
        self.P1, self.P2 = 2 * L * 8, 4 * L * 8
        self.OUTB = 2 * self.P1 + self.P2 + 3          # A, B, C + 3 infinity flags per proof
        self.COMB = self.P1 + 1                        # commitment + flag
        # pipelines: a proof context and a commitment context each (forks of `ctx`: shared key, own streams + scratch)
        self.pipes = []
        for _ in range(inflight):
            ps, cs = torch.cuda.Stream(device=local), torch.cuda.Stream(device=local)
            self.pipes.append(dict(pstream=ps, cstream=cs, ctx=ctx.fork(stream=ps.cuda_stream),
                                   cctx=ctx.fork(stream=cs.cuda_stream)))
        # Phase A is latency-bound (a commitment MSM has a few dozen points): the K commitment batches of a run are
        # independent, so they go to up to 8 commitment contexts and run side by side instead of queueing behind each
        # other on two (3.1 -> ~1 ms for the phase at K = 6..8)
        for _ in range(min(max(max_steps, inflight), 8) - inflight):
            cs = torch.cuda.Stream(device=local)
            self.pipes.append(dict(pstream=None, cstream=cs, ctx=None, cctx=ctx.fork(stream=cs.cuda_stream)))
        self.cpipes = self.pipes                      # every entry has a commitment context
        self.pipes = self.pipes[:inflight]            # the proof pipelines
        self.all_ctxs = [pp["ctx"] for pp in self.pipes] + [pp["cctx"] for pp in self.cpipes]
        self.pstreams = [pp["pstream"] for pp in self.pipes]
        self.cstreams = [pp["cstream"] for pp in self.cpipes]
        # inputs: pinned host copies (e2e) and device copies (device-resident run)
        its = self.items
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).pin_memory()
        self.pins = [pin(np.stack([it[k] for it in its])) for k in ("z", "rm", "sm", "km")]
        hz, hr, hs, hk_ = [t_.numpy().view(np.uint64) for t_ in self.pins]
        self.hz = hz.reshape(batch, -1, 4)
        self.hr, self.hs, self.hk = hr.reshape(batch, 4), hs.reshape(batch, 4), hk_.reshape(batch, 4)
        self.dz = torch.from_numpy(self.hz.view(np.int64)).to(self.dev)
        self.drsk = torch.from_numpy(np.concatenate([hr.ravel(), hs.ravel(), hk_.ravel()]).view(np.int64)).to(self.dev)
        # pageable copies for the blocking HK_HOST path
        self.pz, self.pr, self.ps_, self.pkap = self.hz.copy(), self.hr.copy(), self.hs.copy(), self.hk.copy()
        # one output block per step of a run: commitments and proofs of the whole job are gathered once per phase
        self.nout = max(max_steps, inflight, 4)
        w_p = ((batch * self.OUTB + 15) // 16 + 1) * 2       # int64 words per block; rows stay 16-byte aligned
        w_c = ((batch * self.COMB + 15) // 16 + 1) * 2
        self.d_proofs = torch.zeros((self.nout, w_p), dtype=torch.int64, device=self.dev)
        self.d_coms = torch.zeros((self.nout, w_c), dtype=torch.int64, device=self.dev)
        pinned_u64 = lambda *sh: torch.zeros(sh, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
        pinned_u8 = lambda *sh: torch.zeros(sh, dtype=torch.uint8).pin_memory().numpy()
        self.hout = [dict(a=pinned_u64(batch, 2 * L), b=pinned_u64(batch, 4 * L), c=pinned_u64(batch, 2 * L),
                          inf=pinned_u8(batch, 3), com=pinned_u64(batch, 2 * L), cinf=pinned_u8(batch))
                     for _ in range(self.nout)]
        # bytes that cross PCIe per step: the full assignment + (r, s, kappa) for the proofs, the committed prefix +
        # kappa for the commitments; commitments + proofs + flags back
        self.h2d_per_step = batch * (wl.m * 32 + 96) + batch * ((wl.l + wl.stage_start[1]) * 32 + 32)
        self.d2h_per_step = batch * (self.OUTB + self.COMB)
        self.challenge = 0
        self.prime()

    def prime(self):
        """One-off set-up of EVERY context (not a warm-up step): a forked context allocates its MSM / NTT scratch on
        first use, and cudaMalloc synchronises the device.  A run of K steps spreads its commitment batches over up to 8
        commitment contexts while W = 3 warm-up steps touch only three of them -- the other five used to take their
        first-use allocations inside the timed region (device-resident `value` 281-293 proofs/s run to run against a
        steady 296-300 end to end, which ran later)."""
        b, P1, P2, m = self.batch, self.P1, self.P2, self.wl.m
        q, pc, p = self.drsk.data_ptr(), self.d_coms[0].data_ptr(), self.d_proofs[0].data_ptr()
        for s_ in self.cstreams + self.pstreams:
            s_.wait_stream(self.stream)
        for pp in self.cpipes:
            pp["cctx"].commit_batch_device(self.pk.handle, 0, b, self.dz.data_ptr(), m, q + 64 * b, pc, pc + b * P1)
            pp["cctx"].sync()
        for pp in self.pipes:
            pp["ctx"].prove_batch_device(self.pk.handle, b, self.dz.data_ptr(), q, q + 32 * b, q + 64 * b, p, p + b * P1,
                                         p + b * (P1 + P2), p + b * (2 * P1 + P2))
            pp["ctx"].sync()
        o = self.hout[0]                                 # ... and their host-buffer staging (end-to-end path)
        for pp in self.cpipes:
            pp["cctx"].commit_batch_async(self.pk.handle, 0, self.hz, self.hk, o["com"], o["cinf"])
            pp["cctx"].sync()
        for pp in self.pipes:
            pp["ctx"].prove_batch_async(self.pk.handle, self.hz, self.hr, self.hs, self.hk.reshape(b, -1, 4), o["a"], o["b"],
                                        o["c"], o["inf"])
            pp["ctx"].sync()
        self.torch.cuda.synchronize()

    # ---- the stage barrier ----------------------------------------------------------------------------------
    def _barrier_device(self, steps):
        """After phase A: gather the commitments of all ranks, read them back, derive the stage-1 requests."""
        torch = self.torch
        for s_ in self.cstreams:
            self.stream.wait_stream(s_)
        from hekaton_system_b200.shard import stage_barrier
        with torch.cuda.stream(self.stream):             # the host holds every rank's commitments when this returns
            self.challenge, _ = stage_barrier(self.d_coms[:min(steps, self.nout)], self.world)
        for s_ in self.pstreams:
            s_.wait_stream(self.stream)

    def _gather_proofs_device(self, steps):
        torch = self.torch
        for s_ in self.pstreams:
            self.stream.wait_stream(s_)
        if self.world > 1:
            from hekaton_system_b200.shard import gather_records
            with torch.cuda.stream(self.stream):
                gather_records(self.d_proofs[:min(steps, self.nout)].reshape(1, -1), self.world)

    # ---- device-resident run -------------------------------------------------------------------------------
    def run_device(self, steps):
        b, P1, P2, m = self.batch, self.P1, self.P2, self.wl.m
        q = self.drsk.data_ptr()
        for s_ in self.cstreams + self.pstreams:
            s_.wait_stream(self.stream)
        for i in range(steps):                               # phase A: stage 0 (committer.rs:55-98)
            pc = self.d_coms[i % self.nout].data_ptr()
            self.cpipes[i % len(self.cpipes)]["cctx"].commit_batch_device(self.pk.handle, 0, b, self.dz.data_ptr(), m,
                                                                          q + 64 * b, pc, pc + b * P1)
        self._barrier_device(steps)
        for i in range(steps):                               # phase B: stage 1 (committer.rs:100-122)
            p = self.d_proofs[i % self.nout].data_ptr()
            self.pipes[i % self.inflight]["ctx"].prove_batch_device(self.pk.handle, b, self.dz.data_ptr(), q, q + 32 * b,
                                                                    q + 64 * b, p, p + b * P1, p + b * (P1 + P2),
                                                                    p + b * (2 * P1 + P2))
        self._gather_proofs_device(steps)

    # ---- end to end, page-locked caller buffers, enqueue-only calls -------------------------------------------
    def _exchange_host(self, arrays):
        """all_gather of host-resident records at N > 1 (through a device staging copy over NCCL)."""
        flat = np.concatenate([np.ascontiguousarray(a).view(np.uint8).ravel() for a in arrays])
        if self.world == 1:
            return flat
        torch = self.torch
        from hekaton_system_b200.shard import gather_records
        pad = (-flat.size) % 8
        t = torch.from_numpy(np.concatenate([flat, np.zeros(pad, np.uint8)]).view(np.int64)).to(self.dev)
        with torch.cuda.stream(self.stream):
            return gather_records(t.reshape(1, -1), self.world).cpu().numpy().view(np.uint8).ravel()

    def run_host(self, steps):
        from hekaton_system_b200.shard import derive_stage1_requests
        b = self.batch
        for s_ in self.cstreams + self.pstreams:
            s_.wait_stream(self.stream)
        for i in range(steps):
            o = self.hout[i % self.nout]
            self.cpipes[i % len(self.cpipes)]["cctx"].commit_batch_async(self.pk.handle, 0, self.hz, self.hk, o["com"], o["cinf"])
        for pp in self.cpipes:
            pp["cctx"].sync()                                # commitments are in host memory
        n = min(steps, self.nout)
        self.challenge = derive_stage1_requests(self._exchange_host([self.hout[i]["com"] for i in range(n)]))
        for i in range(steps):
            o = self.hout[i % self.nout]
            self.pipes[i % self.inflight]["ctx"].prove_batch_async(self.pk.handle, self.hz, self.hr, self.hs,
                                                                   self.hk.reshape(b, -1, 4), o["a"], o["b"], o["c"], o["inf"])
        for pp in self.pipes:
            pp["ctx"].sync()
        if self.world > 1:
            self._exchange_host([self.hout[i][k] for i in range(n) for k in ("a", "b", "c")])
        for s_ in self.cstreams + self.pstreams:
            self.stream.wait_stream(s_)

    # ---- end to end, pageable caller buffers, blocking HK_HOST calls, one host thread per pipeline ------------
    def run_host_pageable(self, steps):
        from hekaton_system_b200.shard import derive_stage1_requests
        b = self.batch
        res, coms = [None] * steps, [None] * steps

        def phase(fn):
            def work(p):
                for i in range(p, steps, self.inflight):
                    fn(i, self.pipes[p])
            ts = [threading.Thread(target=work, args=(p,)) for p in range(self.inflight)]
            [t.start() for t in ts]
            [t.join() for t in ts]

        def commit(i, pp):
            coms[i] = pp["cctx"].commit_batch(self.pk.handle, 0, self.pz, self.pkap)[0]

        def prove(i, pp):
            res[i] = pp["ctx"].prove_batch(self.pk.handle, self.pz, self.pr, self.ps_, self.pkap.reshape(b, -1, 4))

        phase(commit)
        self.challenge = derive_stage1_requests(self._exchange_host(coms[:min(steps, self.nout)]))
        phase(prove)
        if self.world > 1:
            self._exchange_host([x for r_ in res[:min(steps, self.nout)] for x in r_[:3]])
        return res

    # ---- checks ----------------------------------------------------------------------------------------------
    def verify_against_trapdoor(self):
        """Proof and commitment of item 0 (fresh blocking calls) == the trapdoor known answer."""
        from hekaton_system_b200 import codec, workload
        ctx, cv, wl, it0 = self.ctx, self.cv, self.wl, self.items[0]
        a, b, c, inf = ctx.prove(self.pk.handle, it0["z"], it0["rm"], it0["sm"], it0["km"])
        com, _ = ctx.commit(self.pk.handle, 0, it0["z"], it0["km"])
        exp = workload.expected_exponents(wl, wl.assignments[it0["k"]], it0["r"], it0["s"], [it0["kappa"]])
        td = wl.trapdoor
        gen1 = codec.g1_affine_to_limbs([cv.g1], cv)[0]
        gen2 = codec.g2_affine_to_limbs([cv.g2], cv)[0]
        g1 = lambda e: e * td["g1"] % cv.r
        h1 = ctx.bases_from_scalars(1, gen1, codec.ints_to_limbs([g1(exp["A"]), g1(exp["C"]), g1(exp["D"][0])], 4))
        h2 = ctx.bases_from_scalars(2, gen2, codec.ints_to_limbs([exp["B"] * td["g2"] % cv.r], 4))
        e1p, e2p = ctx.bases_download(h1, 1), ctx.bases_download(h2, 2)
        ok = bool(np.array_equal(e1p[0], a) and np.array_equal(e1p[1], c) and np.array_equal(e2p[0], b) and
                  np.array_equal(e1p[2], com))
        ctx.bases_free(h1)
        ctx.bases_free(h2)
        self.proof0 = (a, b, c)
        return ok

    def batch_outputs_match_single(self, step=0):
        """The device-resident run's output block of `step` holds, for proof 0, exactly the blocking single call's A,
        B and C (same inputs): the timed path produces the verified results."""
        b, P1, P2 = self.batch, self.P1, self.P2
        raw = self.d_proofs[step % self.nout].cpu().numpy().view(np.uint8)
        A = raw[0:P1].view(np.uint64)
        Bp = raw[b * P1:b * P1 + P2].view(np.uint64)
        Cp = raw[b * (P1 + P2):b * (P1 + P2) + P1].view(np.uint64)
        a0, b0, c0 = self.proof0
        return bool(np.array_equal(A, a0) and np.array_equal(Bp, b0) and np.array_equal(Cp, c0))

    def host_outputs_match_single(self, step=0):
        """After run_host: the page-locked output block of `step` holds, for proof 0, the blocking single call's
        A, B, C (the end-to-end path produces the verified results too)."""
        o = self.hout[step % self.nout]
        a0, b0, c0 = self.proof0
        return bool(np.array_equal(o["a"][0], a0) and np.array_equal(o["b"][0], b0) and np.array_equal(o["c"][0], c0)
                    and not o["inf"].any() and not o["cinf"].any())

    def close(self):
        for c_ in self.all_ctxs:
            c_.close()
        try:
            self.ctx.pk_free(self.pk.handle)
            for h in list(self.pk.handles.values()):
                for hh in (h if isinstance(h, list) else [h]):
                    self.ctx.bases_free(hh)
        except Exception:
            pass


def timed(torch, dist, stream, dev, world, fn, steps):
    """W/K contract: barrier + synchronize on both sides, CUDA events on `stream`, max over ranks."""
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


def timed_wall(torch, dist, dev, world, fn, steps):
    """Host-blocking paths (pageable e2e): wall clock between two full synchronisations, max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(steps)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def profiled(sb, torch, run_timed, serial, steps):
    """Roofline pass with the per-launch events on: (a) as the launches run in the pipeline, (b) with the lanes of
    ONE batch serialised so that every launch of the dominant kernels is timed alone."""
    cs = [sb.pipes[0]["ctx"], sb.pipes[0]["cctx"]] if serial else sb.all_ctxs
    for c_ in cs:
        c_.profile_reset()
        c_.profile(True, serial=serial)
    ms = None
    if serial:
        b, P1, P2 = sb.batch, sb.P1, sb.P2
        q, p_, pc_ = sb.drsk.data_ptr(), sb.d_proofs[0].data_ptr(), sb.d_coms[0].data_ptr()
        for _ in range(steps):
            cs[1].commit_batch_device(sb.pk.handle, 0, b, sb.dz.data_ptr(), sb.wl.m, q + 64 * b, pc_, pc_ + b * P1)
            torch.cuda.synchronize()
            cs[0].prove_batch_device(sb.pk.handle, b, sb.dz.data_ptr(), q, q + 32 * b, q + 64 * b,
                                     p_, p_ + b * P1, p_ + b * (P1 + P2), p_ + b * (2 * P1 + P2))
            torch.cuda.synchronize()
    else:
        ms = run_timed(sb.run_device, steps)
    out = {}
    for k in (0, 1, 2):
        rs = [c_.profile_read(k) for c_ in cs]
        out[k] = tuple(sum(x[j] for x in rs) for j in range(3))
    for c_ in cs:
        c_.profile(False)
        c_.profile_reset()
    return out, ms


def accum_macs(L, dual):
    """Wide multiply-accumulates of one mixed XYZZ addition over a field of L u64 limbs: (algorithmic 8M + 2S,
    executed by this build)."""
    n32 = 2 * L
    mul = 2 * n32 * n32 + n32                          # Montgomery multiply
    sqr = n32 * (n32 + 1) // 2 + n32 * n32 + n32       # dedicated squaring (field_gen: *_sqr)
    alg = 8 * mul + 2 * sqr
    return alg, (6 * mul + 2 * sqr + 3 * n32 * n32 + n32) if dual else alg


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
    W, c = 16, 16                           # planner's choice at this size (msm_plan, table mode)
    macs = n * W * accum_macs(L, False)[0]
    out["msm_g1_2p20"] = {"ms": ms, "Mpts_per_s": n / ms / 1e3, "window_bits": c, "windows": W,
                          "algorithmic_wide_macs": macs, "frac_of_imad_peak": macs / (ms * 1e-3) / imad_peak,
                          "note": "whole MSM (sort + accumulate + reduce), scalars and bases resident in HBM"}
    ctx.bases_free(h)
    del d_sc
    fr_macs = 2 * 8 * 8 + 8
    for lg in (20, 24):
        N = 1 << lg
        a = rng.integers(0, 1 << 60, size=(N, 4), dtype=np.uint64)
        d = torch.from_numpy(a.view(np.int64)).cuda()
        ms = best_ms(lambda: ctx.ntt(d.data_ptr(), lg))
        gbs = 64 * N / (ms * 1e-3) / 1e9
        out[f"ntt_fwd_2p{lg}"] = {"ms": ms, "GBps_algorithmic": gbs, "frac_of_hbm_peak": gbs / hbm_gbs,
                                  "Gbutterflies_per_s": N / 2 * lg / ms / 1e6,
                                  "frac_of_imad_peak": N / 2 * lg * fr_macs / (ms * 1e-3) / imad_peak}
        del d
    return out


def emit(line):
    """The one JSON line goes to the REAL stdout; everything else (library banners such as NCCL's version
    line, build output) was redirected to stderr at start-up."""
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


# ------------------------------------------------------------------------------------------------------------
def field_mul_ns(cref, cid):
    return {"fq": round(cref.bench_field_mul(cid, 1), 1), "fr": round(cref.bench_field_mul(cid, 0), 1)}


def reference_arm(args):
    """`--impl reference`: the reference's CPU path for the same workload, timed on the host cores.  Uses ONLY the
    oracle (C restatement; the arkworks binary cannot be built here): the synthetic proving key is generated on the
    CPU as well, so this process never loads libhekaton_b200.so and never touches the GPU.
    Concurrency follows the reference's worker (mpi-snark/src/bin/node.rs:745-795 pool_and_chunk_size): with T
    threads and R requests it runs min(R, T) proofs concurrently, each on a pool of max(1, T / R) threads."""
    from hekaton_system_b200 import codec, workload          # numpy-only modules (no CUDA library is loaded)
    from hekaton_system_b200.curves import get_curve
    from oracle import cref
    cv = get_curve(args.curve)
    batch = args.batch or SHAPE_BATCH.get(args.shape, 8)
    t0 = time.time()
    wl = workload.make_workload(cv, args.shape, seed=1, num_assignments=2)
    td, ks = workload.key_scalars(wl, seed=7)
    threads = os.cpu_count() or 1
    gen1 = codec.g1_affine_to_limbs([cv.g1], cv)[0][0]
    gen2 = codec.g2_affine_to_limbs([cv.g2], cv)[0][0]
    enc = lambda vals, g: codec.ints_to_limbs([v * g % cv.r for v in vals], 4)
    fb = lambda group, vals: cref.fixed_base_mul(cv.cid, group, gen1 if group == 1 else gen2,
                                                 enc(vals, td["g1"] if group == 1 else td["g2"]), threads)
    L = cv.fq_limbs
    ck = [fb(1, v) if v else np.zeros((0, 2 * L), dtype=np.uint64) for v in ks["ck"]]
    p1 = fb(1, [ks["alpha"], ks["beta"]])
    p2 = fb(2, [ks["beta"], ks["deltas"][-1]])
    cp = cref.CpuProver(cv.cid, wl.n, wl.l, wl.nw, wl.csr, wl.stage_start, fb(1, ks["a_g"]), fb(1, ks["b_g"]),
                        fb(2, ks["b_g"]), fb(1, ks["h_g"]), ck, fb(1, ks["deltas"]), p1[0], p1[1], p2[0], p2[1])
    setup_s = time.time() - t0
    items = build_inputs(wl, batch, 0)
    conc = min(batch, threads)
    pool = max(1, threads // batch)

    def one(it):
        cp.commit(0, it["z"], it["km"], threads=pool)                    # stage 0 (committer.rs:55-98)
        return cp.prove(it["z"], it["rm"], it["sm"], it["km"], threads=pool)    # stage 1 (committer.rs:100-122)

    def step():
        """One bounded sample of the step: `conc` subcircuits of the batch, concurrently, as the reference's worker."""
        out = [None] * conc
        ts = [threading.Thread(target=lambda j=j: out.__setitem__(j, one(items[j % len(items)]))) for j in range(conc)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        return out

    last = None
    for _ in range(min(args.warmup, 1)):
        last = step()
    t0 = time.time()
    for _ in range(args.steps):
        last = step()
    dt = time.time() - t0
    v = args.steps * conc / dt
    # the CPU proof is itself checked against the trapdoor known answer (Python oracle group arithmetic)
    from oracle.ec import groups
    it0 = items[0]
    exp = workload.expected_exponents(wl, wl.assignments[it0["k"]], it0["r"], it0["s"], [it0["kappa"]])
    G1 = groups(cv.name)[0]
    a_pt = codec.g1_affine_from_limbs(last[0][0].reshape(1, -1), cv)[0]
    ok = a_pt == G1.mul(G1.mul(G1.gen, td["g1"]), exp["A"])
    sample = (f"{conc} subcircuit(s) (commit + prove) of the {batch}-subcircuit step per timed step, run concurrently "
              f"on pools of {pool} thread(s) each (node.rs:745-795)")
    config = make_config(args, wl, max(1, args.gpus), batch)       # our arm's config, verbatim
    emit({
        "impl": "reference", "metric": "subcircuit proofs/s", "value": v, "unit": "proofs/s", "n_gpus": 0,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Montgomery, 64-bit limbs)",
        "data": "synthetic", "config": config, "setup_s": round(setup_s, 1), "proof_matches_trapdoor": bool(ok),
        "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": threads, "kind": "port", "sample": sample,
                         "concurrent_proofs": conc, "threads_per_proof": pool,
                         "field_mul_ns_single_thread": field_mul_ns(cref, cv.cid), "mulx_adx_asm": cref.uses_asm(),
                         "note": "C restatement of the arkworks-0.4 algorithms the reference runs (window-parallel "
                                 "Pippenger, radix-2 FFT), -O3, MULX/ADCX/ADOX Montgomery multiplication where the CPU "
                                 "has it (= ark-ff's `asm` feature); not the arkworks binary (no Rust toolchain)"},
        "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return 0


def cpu_baseline_sample(sb):
    """Bounded CPU sample inside the GPU arm (N = 1 only): the oracle's prover on the same key (downloaded from the
    device -- this leg runs after every timed region) and the same assignments; `conc` proofs concurrently as the
    reference's worker would (node.rs:745-795).  Checks the first proof bit for bit against the GPU's."""
    from oracle import cref
    ctx, wl, pk, cv = sb.ctx, sb.wl, sb.pk, sb.cv
    h = pk.handles
    dl = lambda hd, g: ctx.bases_download(hd, g)
    cp = cref.CpuProver(cv.cid, wl.n, wl.l, wl.nw, wl.csr, wl.stage_start, dl(h["a_g"], 1), dl(h["b_g"], 1),
                        dl(h["b_h"], 2), dl(h["h_g"], 1), [dl(c, 1) for c in h["ck"]], dl(h["deltas_g"], 1),
                        pk.points["alpha_g"], pk.points["beta_g"], pk.points["beta_h"], pk.points["last_delta_h"])
    threads = os.cpu_count() or 1
    conc = min(sb.batch, threads)
    pool = max(1, threads // sb.batch)
    out = [None] * conc

    def one(j):
        it = sb.items[j % len(sb.items)]
        cp.commit(0, it["z"], it["km"], threads=pool)
        out[j] = cp.prove(it["z"], it["rm"], it["sm"], it["km"], threads=pool)

    t0 = time.time()
    ts = [threading.Thread(target=one, args=(j,)) for j in range(conc)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.time() - t0
    a, b, c = sb.proof0
    same = bool(np.array_equal(out[0][0], a) and np.array_equal(out[0][1], b) and np.array_equal(out[0][2], c))
    return {"value": conc / dt, "unit": "proofs/s", "cores": threads, "kind": "port",
            "sample": f"{conc} subcircuit(s) (commit + prove) of the same workload, concurrently on pools of {pool} "
                      f"thread(s) each (same key, same assignments as the timed step)",
            "seconds": dt, "concurrent_proofs": conc, "threads_per_proof": pool, "bit_exact_with_gpu": same,
            "field_mul_ns_single_thread": field_mul_ns(cref, cv.cid), "mulx_adx_asm": cref.uses_asm()}


# ------------------------------------------------------------------------------------------------------------
def main():
    global _REAL_STDOUT
    args = parse()
    if not args.sweep:
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)              # fd 1 -> stderr for the rest of the process (C libraries included)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0          # the CPU arm runs on rank 0 only; other ranks exit without work
        import __graft_entry__ as entry
        entry.build_oracle()
        return reference_arm(args)

    import torch
    import torch.distributed as dist
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()

    from hekaton_system_b200 import backend as hk_backend
    from hekaton_system_b200.backend import Context
    from hekaton_system_b200.shard import shard_range

    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ctx = Context(args.curve, device=local, stream=stream.cuda_stream)
    if args.opt:
        names = {"waves": hk_backend.OPT_ACCUM_WAVES, "smin": hk_backend.OPT_ACCUM_SEG_MIN, "fold": hk_backend.OPT_FOLD_SEG,
                 "chunk": hk_backend.OPT_REDUCE_CHUNK, "gsort": hk_backend.OPT_SORT_GLOBAL_ATOMICS, "g2pair": hk_backend.OPT_G2_LANE_PAIR, "quad": hk_backend.OPT_QUAD_MAX_THREADS, "noresolve": hk_backend.OPT_NO_FOLD_RESOLVE}
        for kv in args.opt:
            k_, v_ = kv.split("=")
            ctx.set_option(names[k_], int(v_))         # forks inherit the settings
    inflight = max(1, args.inflight)
    batch = args.batch or SHAPE_BATCH.get(args.shape, 8)
    cv = ctx.curve
    L = cv.fq_limbs
    run_timed = lambda fn, steps: timed(torch, dist, stream, dev, world, fn, steps)

    sb = ShapeBench(ctx, stream, args.shape, batch, inflight, max(args.steps, args.warmup), rank, world, local)
    wl = sb.wl
    assert len(shard_range(batch * world, world, rank)) == batch      # contiguous chunk per rank (node.rs:471-493)
    config = make_config(args, wl, world, batch)

    verified = sb.verify_against_trapdoor()
    imad_peak = ctx.bench_imad_wide(8192)          # wide multiply-accumulates / s (roofline denominator)

    # started before the warm-up so that its start-up cost never lands in the timed region; samples every 100 ms
    sampler = ClockSampler(local) if rank == 0 and not args.timeline else None
    sb.run_device(args.warmup)
    torch.cuda.synchronize()
    if sampler:
        sampler.wait_first()
    if args.timeline:
        from torch.profiler import ProfilerActivity, profile
        nrun = 3 * inflight                       # steady state: the middle third is what tools/timeline_stats.py reads
        if args.timeline_host:
            sb.run_host(2 * inflight)
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            (sb.run_host if args.timeline_host else sb.run_device)(nrun)
            torch.cuda.synchronize()
        tmp = args.timeline + ".chrome.json"
        tp.export_chrome_trace(tmp)
        tr = json.load(open(tmp))
        os.unlink(tmp)
        evs = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        t0_ = min(e["ts"] for e in evs)
        sids = {s_: i for i, s_ in enumerate(sorted({e["args"].get("stream") for e in evs}))}
        with open(args.timeline, "w") as f:
            f.write(f"# {nrun} steps of {batch} proofs, {inflight} in flight\n")
            f.write("# start_us dur_us stream grid block regs kernel\n")
            for e in sorted(evs, key=lambda e: e["ts"]):
                a_ = e["args"]
                g_ = "x".join(str(v) for v in a_.get("grid", [])) or "-"
                b_ = "x".join(str(v) for v in a_.get("block", [])) or "-"
                f.write(f"{e['ts'] - t0_:10.1f} {e['dur']:9.1f} s{sids[a_.get('stream')]} {g_} {b_} "
                        f"{a_.get('registers per thread', '-')} {e['name'][:100]}\n")
        sb.close()
        ctx.close()
        return 0

    l0 = sum(c_.launches for c_ in sb.all_ctxs)
    ms_dev = run_timed(sb.run_device, args.steps)
    launches = sum(c_.launches for c_ in sb.all_ctxs) - l0
    same_as_single = sb.batch_outputs_match_single(0)
    # Roofline pass, outside the timed region (the event records and counter read-backs perturb the pipeline)
    prof, ms_prof = profiled(sb, torch, run_timed, False, max(2, min(args.steps, 4)))
    prof_alone, _ = profiled(sb, torch, run_timed, True, 2)
    # end to end through the host-buffer C-ABI calls
    sb.run_host(min(args.warmup, 2) * inflight)
    ms_e2e = run_timed(sb.run_host, args.steps)
    e2e_same = sb.host_outputs_match_single(0)
    clocks = sampler.stop() if sampler else None
    sb.run_host_pageable(inflight)
    ms_pageable = timed_wall(torch, dist, dev, world, sb.run_host_pageable, args.steps)

    proofs = batch * world * args.steps
    value = proofs / (ms_dev * 1e-3)
    e2e_v = proofs / (ms_e2e * 1e-3)
    hbm_gbs, hbm_src, _ = peaks()
    flags = hk_backend.load_library().hk_build_flags()
    dual = bool(flags & 1) and args.curve == "bls12_381"
    macs_per_add, exec_macs = accum_macs(L, dual)

    # the other shapes (BASELINE configs 1, 4, 5): every rank runs them (weak scaling), rank 0 reports
    shapes = {}
    if not args.no_extras and args.shapes:
        for shp in [s_ for s_ in args.shapes.split(",") if s_ and s_ != args.shape]:
            try:
                b2 = SHAPE_BATCH.get(shp, 8)
                st2 = 3 if shp == "vm" else 6
                s2 = ShapeBench(ctx, stream, shp, b2, inflight, st2, rank, world, local)
                ok2 = s2.verify_against_trapdoor()
                s2.run_device(2)
                ms2 = run_timed(s2.run_device, st2)
                pa, _ = profiled(s2, torch, run_timed, True, 1)
                ms_a, n_a, adds_a = pa[0]
                shapes[shp] = {"workload": workload_string(shp, s2.wl, args.curve), "subcircuits_per_gpu_per_step": b2,
                               "steps": st2, "proofs_per_s": b2 * world * st2 / (ms2 * 1e-3), "ms_per_step": ms2 / st2,
                               "verified_vs_trapdoor": ok2, "setup_s": round(s2.setup_s, 1),
                               "roofline_frac_accum_g1": (adds_a * macs_per_add / (ms_a * 1e-3) / imad_peak) if n_a else None,
                               "accum_g1_avg_launch_ms": (ms_a / n_a) if n_a else None}
                s2.close()
                del s2
                torch.cuda.empty_cache()
            except Exception as e:       # an extra shape must never take the headline line down
                shapes[shp] = {"error": f"{type(e).__name__}: {e}"}

    extras = None
    if rank == 0 and not args.no_extras:
        extras = measure_extras(ctx, stream, imad_peak, hbm_gbs)
        extras["shapes"] = shapes

    if rank == 0:
        traffic = ncu_traffic()
        ms_acc, n_acc, adds = prof_alone[0]
        ms_acc_p, n_acc_p, adds_p = prof[0]
        roof = None
        if n_acc:
            ach = adds * macs_per_add / (ms_acc * 1e-3)
            t_acc = traffic.get("k_msm_accum_g1", {})
            roof = {"kernel": "k_msm_accum<G1>", "bound": "imad", "achieved": ach / 1e12, "peak": imad_peak / 1e12,
                    "unit": "T wide-MAC/s (32x32+64 IMAD.WIDE)", "frac": ach / imad_peak,
                    "traffic": (t_acc["dram_bytes_per_unit"] * adds / n_acc) if t_acc else None,
                    "traffic_note": (f"dram read+write bytes per launch = {t_acc['dram_bytes_per_unit']:.0f} B per addition "
                                     f"(ncu --set full capture {t_acc['source']}) x this run's additions per launch; "
                                     "algorithmic = 104 B per addition (one 96-B affine base + one 8-B sorted pair)")
                    if t_acc else None,
                    "launches": n_acc, "avg_launch_ms": ms_acc / n_acc, "units_per_launch": adds / n_acc,
                    "algorithmic_macs_per_unit": macs_per_add, "executed_macs_per_unit": exec_macs,
                    "frac_executed": ach / imad_peak * exec_macs / macs_per_add,
                    "share_of_step": (ms_acc_p / ms_prof) if ms_prof else None,
                    "in_pipeline": {"frac": (adds_p * macs_per_add / (ms_acc_p * 1e-3) / imad_peak) if n_acc_p else None,
                                    "launches": n_acc_p, "avg_launch_ms": (ms_acc_p / n_acc_p) if n_acc_p else None},
                    "note": "frac: algorithmic MACs (8M + 2S per mixed addition, SURVEY 8d) / time / peak; frac_executed: "
                            "the multiplies this build issues (fused dual product: one reduction fewer per addition) / "
                            "time / peak = multiplier-pipe utilisation.  CUDA events around every launch with the lanes of "
                            "a batch serialised (each launch alone on the GPU, same kernels and inputs, right after the "
                            "timed region); in_pipeline: the same events while lanes and batches overlap, so a launch's "
                            "duration includes the time it shares the SMs with other kernels; share_of_step = summed "
                            "in-pipeline launch time / wall time",
                    "peak_source": "measured in this run by hk_bench_imad_wide (MEASURED_PEAKS.json has no integer peak)"}
        ms_ntt, n_ntt, elems = prof_alone[2]
        ms_ntt_p, n_ntt_p, elems_p = prof[2]
        roof_ntt = None
        if n_ntt:
            gbs = elems * 64 / (ms_ntt * 1e-3) / 1e9
            t_ntt = traffic.get("k_ntt_pass", {})
            roof_ntt = {"kernel": "k_ntt_pass", "bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s",
                        "frac": gbs / hbm_gbs,
                        "traffic": (t_ntt["dram_bytes_per_unit"] * elems / n_ntt) if t_ntt else None,
                        "traffic_note": (f"dram read+write = {t_ntt['dram_bytes_per_unit']:.1f} B per element per pass "
                                         f"({t_ntt['source']}); algorithmic 64 B") if t_ntt else None,
                        "launches": n_ntt, "avg_launch_ms": ms_ntt / n_ntt,
                        "share_of_step": (ms_ntt_p / ms_prof) if ms_prof else None,
                        "peak_source": f"{hbm_src} (MEASURED_PEAKS.json hbm_gbs)",
                        "note": "64 B per element per pass (read + write once); launches timed alone (lanes serialised). "
                                "A pass is multiplier-bound, not HBM-bound (136 wide MACs per butterfly): see extras for "
                                "the IMAD fraction"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline_sample(sb)
            except Exception as e:       # the baseline is a report, never a dependency of the product path
                cpu = {"value": None, "unit": "proofs/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": "subcircuit proofs/s", "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32x2 (Montgomery Fq/Fr on 32x32+64 IMAD.WIDE lanes)", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": "proofs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": sb.h2d_per_step, "d2h_bytes_per_step": sb.d2h_per_step,
                    "buffers": "page-locked caller buffers, enqueue-only calls (HK_HOST_PINNED_ASYNC)",
                    "pageable": {"value": proofs / (ms_pageable * 1e-3), "unit": "proofs/s",
                                 "ms_per_step": ms_pageable / args.steps,
                                 "buffers": f"pageable caller buffers, blocking HK_HOST calls from {inflight} host "
                                            "thread(s), wall clock"}},
            "gpu_launches": launches, "roofline": roof, "roofline_ntt": roof_ntt, "cpu_baseline": cpu, "extras": extras,
            "setup_s": round(sb.setup_s, 1), "verified_vs_trapdoor": verified, "batch_equals_single_call": same_as_single,
            "e2e_equals_single_call": e2e_same,
        }
        emit(line)
    sb.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
