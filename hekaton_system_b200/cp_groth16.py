"""Host-side mirror of the reference's cp-groth16 operator interface for the hot path, over the
C ABI.  Same names, argument meaning and error behaviour as the Rust (the tests read like
cp-groth16/src/lib.rs:139-180):

    reference (Rust)                                   here
    -------------------------------------------------  -----------------------------------------
    ProvingKey<E>            data_structures.rs:66-83  ProvingKey (device-resident base sets)
    MultiStageConstraintSystem                          StagedAssignment (matrices + assignment +
        constraint_synthesizer.rs:14-117                   per-stage witness ranges)
    CommitmentBuilder::new   committer.rs:38-52        CommitmentBuilder(assignment, pk)
    CommitmentBuilder::commit  committer.rs:55-98      .commit(rng | kappa=)
    CommitmentBuilder::prove   committer.rs:100-122    .prove(comms, comm_rands, rng | r=, s=)
    Proof{a,b,c,ds}          data_structures.rs:7-16   Proof

Constraint synthesis (gadgets) is outside the hot path and stays on the host: a
StagedAssignment is what `generate_constraints` + `cs.finalize()` leave behind.  Randomness is
drawn on the host in the reference's order (kappa in commit: committer.rs:85; r then s in
prove: prover.rs:28-29); the device never sees an RNG.
"""
from dataclasses import dataclass, field

import numpy as np

from . import codec
from .backend import SCALAR_MONT


# ---- ark-serialize 0.4 encodings (CanonicalSerialize) -----------------------------------------------------
# Restated from the published crates (third-party, un-vendored).
# * BN254 (ark-bn254 0.4 uses ark-ec's generic short-Weierstrass impl): field elements are canonical
#   little-endian bytes (8 per limb); a point is x (|| y when uncompressed) with two flag bits OR-ed into the
#   most significant byte written last: 0x80 = "y is the larger of {y, -y}", 0x40 = point at infinity (then
#   x = y = 0); Fq2 = c0 || c1.
# * BLS12-381 (ark-bls12-381 0.4 overrides SWCurveConfig::{serialize,deserialize}_with_mode for G1 and G2 with
#   the zcash / IETF pairing-friendly-curves encoding, curves/util.rs): field elements are BIG-endian, Fq2 is
#   written c1 || c0, and three flag bits sit in the top of the FIRST byte: 0x80 = compressed, 0x40 = infinity,
#   0x20 = y is the lexicographically larger root (set only for compressed, finite points).
# Both orders compare Fq2 elements by c1 first, then c0.  Vec<T> = u64 LE length + items.
# PINNING: the BLS12-381 encoders reproduce the published zcash serialisations of the G1 / G2 generators
# (tests/test_serialize.py); the BN254 flag layout is restated from ark-serialize and has no external vector.
def _y_is_negative_fq(y, p):
    return y > (p - y) % p


def _y_is_negative_fq2(y, p):
    ny = ((p - y[0]) % p, (p - y[1]) % p)
    return (y[1], y[0]) > (ny[1], ny[0])


def _is_zcash(curve):
    return curve.name == "bls12_381"


def _zcash_point(coords, negative, nb, compress):
    """coords: field elements in file order (x [, y]; Fq2 already as c1, c0) or None for infinity."""
    b = bytearray(b"".join(c.to_bytes(nb, "big") for c in coords))
    if compress:
        b[0] |= 0x80
        if negative:
            b[0] |= 0x20
    return bytes(b)


def serialize_g1(P, curve, compress=False):
    nb = 8 * curve.fq_limbs
    if _is_zcash(curve):
        if P is None:
            b = bytearray(nb if compress else 2 * nb)
            b[0] |= 0x40 | (0x80 if compress else 0)
            return bytes(b)
        x, y = P
        return _zcash_point([x] if compress else [x, y], _y_is_negative_fq(y, curve.p), nb, compress)
    if P is None:
        b = bytearray(nb if compress else 2 * nb)
        b[-1] |= 0x40
        return bytes(b)
    x, y = P
    b = bytearray(x.to_bytes(nb, "little") + (b"" if compress else y.to_bytes(nb, "little")))
    if _y_is_negative_fq(y, curve.p):
        b[-1] |= 0x80
    return bytes(b)


def serialize_g2(P, curve, compress=False):
    nb = 8 * curve.fq_limbs
    if _is_zcash(curve):
        if P is None:
            b = bytearray(2 * nb if compress else 4 * nb)
            b[0] |= 0x40 | (0x80 if compress else 0)
            return bytes(b)
        (x0, x1), (y0, y1) = P
        return _zcash_point([x1, x0] if compress else [x1, x0, y1, y0], _y_is_negative_fq2((y0, y1), curve.p), nb,
                            compress)
    if P is None:
        b = bytearray(2 * nb if compress else 4 * nb)
        b[-1] |= 0x40
        return bytes(b)
    (x0, x1), (y0, y1) = P
    b = bytearray(x0.to_bytes(nb, "little") + x1.to_bytes(nb, "little") +
                  (b"" if compress else y0.to_bytes(nb, "little") + y1.to_bytes(nb, "little")))
    if _y_is_negative_fq2((y0, y1), curve.p):
        b[-1] |= 0x80
    return bytes(b)


def _deser_fq(buf, nb, last=False):
    b = bytearray(buf[:nb])
    flags = 0
    if last:
        flags = b[-1] & 0xC0
        b[-1] &= 0x3F
    return int.from_bytes(b, "little"), flags


def _zcash_fields(buf, nb, count):
    """-> (flags byte, [count big-endian field elements]) of an UNCOMPRESSED zcash-format point."""
    if len(buf) < nb * count:
        raise ValueError("point truncated")
    flags = buf[0] & 0xE0
    if flags & 0x80:
        raise ValueError("compressed flag set on an uncompressed point")
    first = bytes([buf[0] & 0x1F]) + bytes(buf[1:nb])
    vals = [int.from_bytes(first, "big")] + [int.from_bytes(buf[i * nb:(i + 1) * nb], "big") for i in range(1, count)]
    return flags, vals


def deserialize_g1_uncompressed(buf, curve):
    nb = 8 * curve.fq_limbs
    if _is_zcash(curve):
        flags, (x, y) = _zcash_fields(buf, nb, 2)
        return None if flags & 0x40 else (x, y)
    x, _ = _deser_fq(buf, nb)
    y, flags = _deser_fq(buf[nb:], nb, last=True)
    return None if flags & 0x40 else (x, y)


def deserialize_g2_uncompressed(buf, curve):
    nb = 8 * curve.fq_limbs
    if _is_zcash(curve):
        flags, (x1, x0, y1, y0) = _zcash_fields(buf, nb, 4)
        return None if flags & 0x40 else ((x0, x1), (y0, y1))
    x0, _ = _deser_fq(buf, nb)
    x1, _ = _deser_fq(buf[nb:], nb)
    y0, _ = _deser_fq(buf[2 * nb:], nb)
    y1, flags = _deser_fq(buf[3 * nb:], nb, last=True)
    return None if flags & 0x40 else ((x0, x1), (y0, y1))


@dataclass
class Proof:
    a: tuple
    b: tuple
    c: tuple
    ds: list = field(default_factory=list)

    def serialize(self, curve, compress=False):
        """`#[derive(CanonicalSerialize)] struct Proof { a, b, c, ds }` (data_structures.rs:7-16)."""
        out = serialize_g1(self.a, curve, compress) + serialize_g2(self.b, curve, compress)
        out += serialize_g1(self.c, curve, compress)
        out += len(self.ds).to_bytes(8, "little")
        for d in self.ds:
            out += serialize_g1(d, curve, compress)
        return out

    @classmethod
    def deserialize_uncompressed(cls, buf, curve):
        nb = 8 * curve.fq_limbs
        a = deserialize_g1_uncompressed(buf, curve)
        b = deserialize_g2_uncompressed(buf[2 * nb:], curve)
        c = deserialize_g1_uncompressed(buf[6 * nb:], curve)
        n = int.from_bytes(buf[8 * nb:8 * nb + 8], "little")
        off = 8 * nb + 8
        ds = [deserialize_g1_uncompressed(buf[off + 2 * nb * i:], curve) for i in range(n)]
        return cls(a, b, c, ds)


class StagedAssignment:
    """Matrices in CSR (columns over [instance || witness]), the assignment, and the witness
    range of every stage (constraint_synthesizer.rs:21,53-64,96-106)."""

    def __init__(self, curve, num_instance, a_rows, b_rows, c_rows, instance, witness, stage_ranges):
        self.curve = curve
        self.l = num_instance
        self.rows = (a_rows, b_rows, c_rows)
        self.instance = list(instance)
        self.witness = list(witness)
        self.stage_ranges = list(stage_ranges)
        assert len(self.instance) == self.l and self.instance[0] == 1
        for (s0, e0), (s1, _) in zip(self.stage_ranges, self.stage_ranges[1:]):
            assert e0 == s1, "stage witness ranges must be contiguous"

    @classmethod
    def from_circuit(cls, circ, curve):
        """Adopt a circuit object exposing .r1cs{num_instance,a,b,c}, .instance, .witness, .stage_ranges."""
        from .curves import get_curve
        r1 = circ.r1cs
        return cls(get_curve(curve), r1.num_instance, r1.a, r1.b, r1.c, circ.instance, circ.witness,
                   circ.stage_ranges)

    @property
    def num_constraints(self):
        return len(self.rows[0])

    def full_assignment(self):
        return self.instance + self.witness

    def csr(self):
        r = self.curve.r
        out = []
        for rows in self.rows:
            ptr = np.zeros(len(rows) + 1, dtype=np.uint32)
            cols, vals = [], []
            for i, row in enumerate(rows):
                for cf, j in row:
                    cols.append(j)
                    vals.append(cf % r)
                ptr[i + 1] = len(cols)
            out.append((ptr, np.array(cols, dtype=np.uint32), codec.to_mont(vals, r, 4) if vals
                        else np.zeros((0, 4), dtype=np.uint64)))
        return out


class ProvingKey:
    """Device-resident proving key: every vector of data_structures.rs:66-114 is uploaded once."""

    def __init__(self, ctx, handle, num_stages, l, nw, log_n, vk=None):
        self.ctx, self.handle, self.num_stages, self.l, self.nw, self.log_n = ctx, handle, num_stages, l, nw, log_n
        self.vk = vk

    @classmethod
    def from_points(cls, ctx, staged, a_g, b_g, b_h, h_g, deltas_abc_g, deltas_g, alpha_g, beta_g, beta_h,
                    last_delta_h):
        cv = ctx.curve
        up1 = lambda pts: ctx.upload_g1(*codec.g1_affine_to_limbs(pts, cv))
        up2 = lambda pts: ctx.upload_g2(*codec.g2_affine_to_limbs(pts, cv))
        h_a, h_b1, h_b2, h_h = up1(a_g), up1(b_g), up2(b_h), up1(h_g)
        h_ck = [up1(v) for v in deltas_abc_g]
        h_d = up1(deltas_g)
        return cls.from_handles(ctx, staged, h_a, h_b1, h_h, h_b2, h_ck, h_d, alpha_g, beta_g, beta_h, last_delta_h)

    @classmethod
    def from_handles(cls, ctx, staged, h_a, h_b1, h_h, h_b2, h_ck, h_d, alpha_g, beta_g, beta_h, last_delta_h):
        cv = ctx.curve
        return cls.from_handles_raw(ctx, staged, h_a, h_b1, h_h, h_b2, h_ck, h_d,
                                    codec.g1_affine_to_limbs([alpha_g], cv)[0][0], codec.g1_affine_to_limbs([beta_g], cv)[0][0],
                                    codec.g2_affine_to_limbs([beta_h], cv)[0][0],
                                    codec.g2_affine_to_limbs([last_delta_h], cv)[0][0])

    @classmethod
    def from_handles_raw(cls, ctx, staged, h_a, h_b1, h_h, h_b2, h_ck, h_d, alpha_g, beta_g, beta_h, last_delta_h):
        """Same, with the four loose vk points already as Montgomery limb arrays."""
        n = staged.num_constraints
        nw = len(staged.witness)
        r1 = ctx.r1cs_upload(n, staged.l, nw, staged.csr())
        starts = [s for s, _ in staged.stage_ranges] + [staged.stage_ranges[-1][1]]
        pk = ctx.pk_create(h_a, h_b1, h_h, h_b2, h_ck, h_d, alpha_g, beta_g, beta_h, last_delta_h, r1, starts)
        log_n = max(1, (n + staged.l - 1).bit_length())
        return cls(ctx, pk, len(h_ck), staged.l, nw, log_n)

    @classmethod
    def from_key(cls, ctx, opk, circ):
        staged = StagedAssignment.from_circuit(circ, ctx.curve)
        return cls.from_points(ctx, staged, opk.a_g, opk.b_g, opk.b_h, opk.h_g, opk.deltas_abc_g, opk.deltas_g,
                               opk.alpha_g, opk.beta_g, opk.beta_h, opk.last_delta_h)


class CommitmentBuilder:
    """committer.rs:15-122.  `assignment` plays the role of (circuit, cs) after synthesis."""

    def __init__(self, assignment, pk):
        self.cs = assignment
        self.pk = pk
        self.cur_stage = 0
        cv = pk.ctx.curve
        self._z = codec.to_mont(assignment.full_assignment(), cv.r, 4)

    def _fr(self, v):
        return codec.to_mont([v], self.pk.ctx.curve.r, 4)[0]

    def commit(self, rng=None, kappa=None):
        """-> (Comm, CommRandomness).  Raises like the reference's `expect` when no stage is left."""
        if self.cur_stage >= self.pk.num_stages:
            raise IndexError("no more values left in committing key")
        if kappa is None:
            kappa = rng.fr_rand()
        ctx = self.pk.ctx
        out, inf = ctx.commit(self.pk.handle, self.cur_stage, self._z, self._fr(kappa))
        self.cur_stage += 1
        com = None if inf else codec.g1_affine_from_limbs(out, ctx.curve)[0]
        return com, kappa

    def prove(self, comms, comm_rands, rng=None, r=None, s=None):
        """committer.rs:100-122: requires deltas_g.len() == comm_rands.len() + 1."""
        assert self.pk.num_stages == len(comm_rands) + 1, "deltas_g.len() must equal comm_rands.len() + 1"
        if r is None:
            r = rng.fr_rand()
        if s is None:
            s = rng.fr_rand()
        ctx = self.pk.ctx
        cv = ctx.curve
        kap = codec.to_mont(list(comm_rands), cv.r, 4) if comm_rands else np.zeros((0, 4), dtype=np.uint64)
        a, b, c, inf = ctx.prove(self.pk.handle, self._z, self._fr(r), self._fr(s), kap)
        A = None if inf[0] else codec.g1_affine_from_limbs(a, cv)[0]
        B = None if inf[1] else codec.g2_affine_from_limbs(b, cv)[0]
        Cc = None if inf[2] else codec.g1_affine_from_limbs(c, cv)[0]
        return Proof(A, B, Cc, list(comms))
