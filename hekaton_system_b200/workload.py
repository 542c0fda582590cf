"""Synthetic subcircuit workloads with the statistics of the reference's experiments.

The reference's real circuits (SHA-256 Merkle tree `tree_hash_circuit.rs`, VKD, VM) are
synthesised by arkworks gadgets that cannot run here, so the benchmark workload is a synthetic
2-stage R1CS of matched SHAPE: l = 4 instance variables (ROM circuits:
one, two challenges, root -- subcircuit_circuit.rs:99-102), 4 stage-0 witnesses per portal op
(rom_transcript.rs:298-303), roughly one new variable per constraint, and bit-heavy witnesses
(booleanity gates, as in SHA-256 bit decompositions).  The proving key is a VALID CP-Groth16
key for this R1CS, built as generator.rs:18-238 does but from explicit toxic waste, with the
group elements produced on the device (hk_bases_from_scalars) -- so every proof of the
benchmark is checkable against the trapdoor known answer.

This is data generation (host, Python integers), not the hot path.
"""
import random
from dataclasses import dataclass, field

import numpy as np

from . import codec
from .curves import get_curve

SHAPES = {
    # name: (constraints, portal ops in stage 0)        domain size
    "big-merkle-small": (60_000, 1),    # config 1: 1 SHA2 iter, 1 portal  -> N = 2^16
    "big-merkle": (200_000, 8),         # config 3: 4 SHA2 iters, 8 portals -> N = 2^18
    "vm": (1_140_000, 54_320),          # config 5: 1024 cycles x 1114 -> N = 2^21; stage 0 = 70 x 3104 = 217 280
                                        # committed trace values (a 2^17.7-point commitment MSM)
    "vkd": (12_000, 4),                 # config 4: 32 Poseidon two-to-one hashes + exec-tree path -> N = 2^14
    "tiny": (500, 2),                   # tests / smoke
}


@dataclass
class Workload:
    curve: object
    name: str
    n: int                      # constraints
    l: int                      # instance variables (incl. one)
    nw: int                     # witness variables
    stage_start: list           # [0, |stage0|, nw]
    csr: list                   # [(ptr, col, val_mont)] * 3
    rows: tuple                 # (a_rows, b_rows, c_rows) python lists (coeff, col)
    assignments: list = field(default_factory=list)   # list of full assignments (python ints)
    trapdoor: dict = None
    key_scalars: dict = None

    @property
    def m(self):
        return self.l + self.nw

    @property
    def log_n(self):
        return max(1, (self.n + self.l - 1).bit_length())


def _lc(rng, nvars, r):
    terms = []
    for _ in range(rng.randrange(1, 4)):
        cf = rng.choice((1, 1, 1, 2, r - 1))
        terms.append((cf, rng.randrange(nvars)))
    return terms


def make_workload(curve, shape="big-merkle", seed=1, num_assignments=1, bool_frac=0.5, n=None, portals=None):
    """Build the R1CS (shared by every subcircuit of this shape) and `num_assignments`
    satisfying assignments (different witnesses, same matrices)."""
    cv = get_curve(curve)
    r = cv.r
    if n is None:
        n, portals = SHAPES[shape]
    rng = random.Random(seed)
    l = 4
    s0 = 4 * portals
    a_rows, b_rows, c_rows = [], [], []
    kinds = []                      # per constraint: None for a boolean gate
    nvars = l + s0
    for _ in range(n):
        if rng.random() < bool_frac:
            j = nvars
            a_rows.append([(1, j)])
            b_rows.append([(1, 0), (r - 1, j)])
            c_rows.append([])
            kinds.append(None)
        else:
            ra, rb = _lc(rng, nvars, r), _lc(rng, nvars, r)
            a_rows.append(ra)
            b_rows.append(rb)
            c_rows.append([(1, nvars)])
            kinds.append((ra, rb))
        nvars += 1
    nw = nvars - l
    csr = []
    for rows in (a_rows, b_rows, c_rows):
        ptr = np.zeros(n + 1, dtype=np.uint32)
        cols, vals = [], []
        for i, row in enumerate(rows):
            for cf, j in row:
                cols.append(j)
                vals.append(cf)
            ptr[i + 1] = len(cols)
        # few distinct coefficients: encode each once
        uniq = {v: codec.to_mont([v], r, 4)[0] for v in set(vals)}
        val = np.stack([uniq[v] for v in vals]) if vals else np.zeros((0, 4), dtype=np.uint64)
        csr.append((ptr, np.array(cols, dtype=np.uint32), val))
    wl = Workload(cv, shape, n, l, nw, [0, s0, nw], csr, (a_rows, b_rows, c_rows))
    for k in range(num_assignments):
        wl.assignments.append(_assign(wl, kinds, random.Random(seed * 1000 + k)))
    return wl


def _assign(wl, kinds, rng):
    r = wl.curve.r
    z = [1] + [rng.randrange(r) for _ in range(wl.l - 1)]
    s0 = wl.stage_start[1]
    z += [rng.randrange(2) if rng.random() < 0.5 else rng.randrange(r) for _ in range(s0)]
    for kd in kinds:
        if kd is None:
            z.append(rng.randrange(2))
        else:
            ra, rb = kd
            va = sum(cf * z[j] for cf, j in ra) % r
            vb = sum(cf * z[j] for cf, j in rb) % r
            z.append(va * vb % r)
    return z


def _batch_inverse(vals, r):
    pref = [1] * (len(vals) + 1)
    for i, v in enumerate(vals):
        pref[i + 1] = pref[i] * v % r
    inv = pow(pref[-1], -1, r)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % r
        inv = inv * vals[i] % r
    return out


def key_scalars(wl, seed=7):
    """Exponents of every proving-key element w.r.t. the generators (generator.rs:60-196 with
    explicit toxic waste).  Returns (trapdoor dict, scalars dict)."""
    cv = wl.curve
    r = cv.r
    rng = random.Random(seed)
    nz = lambda: rng.randrange(1, r)
    td = dict(alpha=nz(), beta=nz(), gamma=nz(), deltas=[nz(), nz()], tau=nz(), g1=nz(), g2=nz())
    n, l = wl.n, wl.l
    N = 1 << wl.log_n
    tau = td["tau"]
    root = pow(cv.fr_generator, (r - 1) >> cv.two_adicity, r)
    w = pow(root, 1 << (cv.two_adicity - wl.log_n), r)
    zt = (pow(tau, N, r) - 1) % r
    # Lagrange coefficients L_i(tau) = zt * w^i / (N (tau - w^i)), only rows < n + l are used
    ws, x = [], 1
    for _ in range(n + l):
        ws.append(x)
        x = x * w % r
    den = _batch_inverse([N * (tau - wi) % r for wi in ws], r)
    u = [zt * wi % r * d % r for wi, d in zip(ws, den)]
    a = [0] * wl.m
    b = [0] * wl.m
    c = [0] * wl.m
    for i in range(l):
        a[i] = u[n + i]
    ar, br, cr = wl.rows
    for i in range(n):
        ui = u[i]
        for cf, j in ar[i]:
            a[j] = (a[j] + ui * cf) % r
        for cf, j in br[i]:
            b[j] = (b[j] + ui * cf) % r
        for cf, j in cr[i]:
            c[j] = (c[j] + ui * cf) % r
    inv = lambda v: pow(v, -1, r)
    al, be = td["alpha"], td["beta"]
    abc = [(be * a[i] + al * b[i] + c[i]) % r for i in range(wl.m)]
    ck = []
    for d, (s, e) in zip(td["deltas"], zip(wl.stage_start[:-1], wl.stage_start[1:])):
        di = inv(d)
        ck.append([abc[i] * di % r for i in range(s + l, e + l)])
    hq, sc, x = [], zt * inv(td["deltas"][-1]) % r, 1
    for _ in range(N - 1):
        hq.append(sc * x % r)
        x = x * tau % r
    ks = dict(a_g=a, b_g=b, h_g=hq, ck=ck, deltas=list(td["deltas"]), alpha=al, beta=be,
              gamma_abc=[abc[i] * inv(td["gamma"]) % r for i in range(l)], c_col=c, zt=zt)
    wl.trapdoor, wl.key_scalars = td, ks
    return td, ks


def expected_exponents(wl, z, r_, s_, kappas, cache=None):
    """Discrete logs (w.r.t. G resp. H) of D_i, A, B, C from O(m) scalar arithmetic.
    `cache` (a dict owned by the caller, one per assignment) keeps the O(m) inner products, which do not
    depend on (r, s, kappa)."""
    r = wl.curve.r
    td, ks = wl.trapdoor, wl.key_scalars
    inv = lambda v: pow(v, -1, r)
    l = wl.l
    if cache is None:
        cache = {}
    if "a_tau" not in cache:
        cache["a_tau"] = sum(zi * ai for zi, ai in zip(z, ks["a_g"])) % r
        cache["b_tau"] = sum(zi * bi for zi, bi in zip(z, ks["b_g"])) % r
        cache["c_tau"] = sum(zi * ci for zi, ci in zip(z, ks["c_col"])) % r
        cache["stage"] = [sum(z[l + s0 + j] * x for j, x in enumerate(ck)) % r
                          for ck, s0 in zip(ks["ck"], wl.stage_start[:-1])]
    a_tau, b_tau, c_tau = cache["a_tau"], cache["b_tau"], cache["c_tau"]
    d_last = td["deltas"][-1]
    a_exp = (td["alpha"] + a_tau + r_ * d_last) % r
    b_exp = (td["beta"] + b_tau + s_ * d_last) % r
    h_tau = (a_tau * b_tau - c_tau) % r * inv(ks["zt"]) % r
    ds = [(cache["stage"][i] + kap * d_last) % r for i, kap in enumerate(kappas)]
    l_exp = cache["stage"][-1]
    h_exp = h_tau * ks["zt"] % r * inv(d_last) % r
    c_exp = (s_ * a_exp + r_ * b_exp - r_ * s_ * d_last + l_exp + h_exp
             - sum(k * d for k, d in zip(kappas, td["deltas"]))) % r
    return dict(A=a_exp, B=b_exp, C=c_exp, D=ds)


def build_device_key(ctx, wl):
    """Materialise the proving key on the device from its exponents and register it with the
    fused path.  Returns a hekaton_system_b200.cp_groth16.ProvingKey plus host copies of the
    few points bench/tests need."""
    from .cp_groth16 import ProvingKey, StagedAssignment
    cv = ctx.curve
    r = cv.r
    td, ks = wl.trapdoor, wl.key_scalars
    gen1 = codec.g1_affine_to_limbs([cv.g1], cv)[0]
    gen2 = codec.g2_affine_to_limbs([cv.g2], cv)[0]
    g1s, g2s = td["g1"], td["g2"]
    enc = lambda vals, g: codec.ints_to_limbs([v * g % r for v in vals], 4)
    h_a = ctx.bases_from_scalars(1, gen1, enc(ks["a_g"], g1s))
    h_b1 = ctx.bases_from_scalars(1, gen1, enc(ks["b_g"], g1s))
    h_b2 = ctx.bases_from_scalars(2, gen2, enc(ks["b_g"], g2s))
    h_h = ctx.bases_from_scalars(1, gen1, enc(ks["h_g"], g1s))
    h_ck = [ctx.bases_from_scalars(1, gen1, enc(v, g1s)) if v else ctx.upload_g1(np.zeros((0, 2 * cv.fq_limbs), dtype=np.uint64))
            for v in ks["ck"]]
    h_d = ctx.bases_from_scalars(1, gen1, enc(ks["deltas"], g1s))
    # the four loose points (alpha_g, beta_g in G1; beta_h, last_delta_h in G2)
    h_p1 = ctx.bases_from_scalars(1, gen1, enc([ks["alpha"], ks["beta"]], g1s))
    h_p2 = ctx.bases_from_scalars(2, gen2, enc([ks["beta"], ks["deltas"][-1]], g2s))
    p1 = ctx.bases_download(h_p1, 1)
    p2 = ctx.bases_download(h_p2, 2)
    staged = _StagedView(wl)
    pk = ProvingKey.from_handles_raw(ctx, staged, h_a, h_b1, h_h, h_b2, h_ck, h_d, p1[0], p1[1], p2[0], p2[1])
    pk.handles = dict(a_g=h_a, b_g=h_b1, b_h=h_b2, h_g=h_h, ck=h_ck, deltas_g=h_d)
    pk.points = dict(alpha_g=p1[0], beta_g=p1[1], beta_h=p2[0], last_delta_h=p2[1])
    ctx.bases_free(h_p1)
    ctx.bases_free(h_p2)
    return pk


class _StagedView:
    """The parts of StagedAssignment that ProvingKey needs, backed by a Workload."""

    def __init__(self, wl):
        self.wl = wl
        self.l = wl.l
        self.witness = range(wl.nw)
        self.stage_ranges = list(zip(wl.stage_start[:-1], wl.stage_start[1:]))

    @property
    def num_constraints(self):
        return self.wl.n

    def csr(self):
        return self.wl.csr


def dump_workload(wl, path, subs):
    """Write the binary workload file read by host/hk_node.cpp (the C++ `node work`-style driver).
    subs: list of dicts with Montgomery limb arrays z (m,4), kappa, r, s (4,) and the expected
    outputs D, A (2L), B (4L), C (2L)."""
    cv = wl.curve
    r = cv.r
    td, ks = wl.trapdoor, wl.key_scalars
    u64 = lambda *v: np.array(v, dtype="<u8").tobytes()
    enc = lambda vals, g: codec.ints_to_limbs([v * g % r for v in vals], 4).tobytes()
    with open(path, "wb") as f:
        f.write(u64(0x314c574b48, cv.cid, wl.n, wl.l, wl.nw, len(wl.stage_start) - 1))
        f.write(np.array(wl.stage_start, dtype="<u8").tobytes())
        for ptr, col, val in wl.csr:
            f.write(u64(len(col)))
            f.write(np.ascontiguousarray(ptr, dtype="<u4").tobytes())
            f.write(np.ascontiguousarray(col, dtype="<u4").tobytes())
            f.write(np.ascontiguousarray(val, dtype="<u8").tobytes())
        f.write(codec.g1_affine_to_limbs([cv.g1], cv)[0].tobytes())
        f.write(codec.g2_affine_to_limbs([cv.g2], cv)[0].tobytes())
        g1s, g2s = td["g1"], td["g2"]
        f.write(enc(ks["a_g"], g1s))
        f.write(enc(ks["b_g"], g1s))
        f.write(enc(ks["b_g"], g2s))
        f.write(enc(ks["h_g"], g1s))
        for v in ks["ck"]:
            f.write(enc(v, g1s))
        f.write(enc(ks["deltas"], g1s))
        f.write(enc([ks["alpha"], ks["beta"]], g1s))
        f.write(enc([ks["beta"], ks["deltas"][-1]], g2s))
        f.write(u64(len(subs)))
        for s in subs:
            for k in ("z", "kappa", "r", "s", "D", "A", "B", "C"):
                f.write(np.ascontiguousarray(s[k], dtype="<u8").tobytes())
