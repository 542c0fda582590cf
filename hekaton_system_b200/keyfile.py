"""Reader / writer of the reference's proving-key files.

`setup-big-merkle|vkd|vm|r1cs --key-out FILE` writes `ProvingKeys::serialize_uncompressed`
(/root/reference/mpi-snark/src/bin/node.rs:314-319,384-389,425-430,453-458) and `work` reads it back with
`deserialize_uncompressed_unchecked` (node.rs:230-237).  Layout, from the `#[derive(CanonicalSerialize)]`
field order (ark-serialize 0.4: fields in declaration order, `usize` as u64 LE, `Vec<T>` / `String` /
`BTreeMap<K, V>` = u64 LE length then the items, map entries as key then value in key order):

    ProvingKeys  (mpi-snark/src/data_structures.rs:41-50)
        circuit_id: String, serialized_circ_params: Vec<u8>,
        minimal_proving_keys: BTreeMap<usize, ProvingKey>, subcircuit_representative_map: BTreeMap<usize, usize>
    ProvingKey   (cp-groth16/src/data_structures.rs:66-83)
        vk: VerifyingKey, beta_g: G1, a_g: Vec<G1>, b_g: Vec<G1>, b_h: Vec<G2>, h_g: Vec<G1>,
        ck: CommitterKey, deltas_g: Vec<G1>
    VerifyingKey (data_structures.rs:32-46)
        alpha_g: G1, beta_h: G2, gamma_h: G2, last_delta_h: G2, gamma_abc_g: Vec<G1>, deltas_h: Vec<G2>
    CommitterKey (data_structures.rs:107-114)
        last_delta_g: G1, deltas_abc_g: Vec<Vec<G1>>

Uncompressed affine point, BN254 (ark-ec generic impl): x || y, canonical little-endian field elements
(Fq2 = c0 || c1), flags in the two top bits of the LAST byte (0x80: y is the larger of {y, -y}; 0x40: infinity,
x = y = 0).  BLS12-381 (ark-bls12-381 0.4 overrides the point encoding with the zcash format): x || y, BIG-endian
field elements, Fq2 = c1 || c0, flags in the top three bits of the FIRST byte (0x80 compressed -- never set in
these files --, 0x40 infinity, 0x20 unused when uncompressed).  See cp_groth16.py for the encoders.

Point vectors are decoded in bulk with numpy into canonical 64-bit limb arrays plus an infinity mask (a
2^18-constraint key holds ~10^6 points); `to_device` converts them to Montgomery form on the GPU
(hk_field_op TO_MONT) and hands them to `hk_pk_create`.  PINNING: unpinned -- the reference ships no key
file and cannot be built here; the tests check write -> read round trips, the flag rules above against
hand-built points, and a proof from a key that went through a file against the trapdoor known answer.
"""
import io
import struct
from dataclasses import dataclass, field

import numpy as np

from . import cp_groth16 as host
from .curves import get_curve


@dataclass
class PointVec:
    """n affine points: canonical little-endian u64 limbs (n, width) and an infinity mask (n,)."""
    limbs: np.ndarray
    inf: np.ndarray

    def __len__(self):
        return self.limbs.shape[0]

    def to_points(self, group):
        """-> list of (x, y) | None (G1) or ((x0, x1), (y0, y1)) | None (G2) as Python integers."""
        from . import codec
        L = self.limbs.shape[1] // (2 if group == 1 else 4)
        flat = codec.limbs_to_ints(self.limbs.reshape(-1, L))
        out = []
        step = 2 if group == 1 else 4
        for i in range(len(self)):
            if self.inf[i]:
                out.append(None)
            elif group == 1:
                out.append((flat[2 * i], flat[2 * i + 1]))
            else:
                c = flat[step * i:step * i + 4]
                out.append(((c[0], c[1]), (c[2], c[3])))
        return out


@dataclass
class VerifyingKeyData:
    alpha_g: PointVec
    beta_h: PointVec
    gamma_h: PointVec
    last_delta_h: PointVec
    gamma_abc_g: PointVec
    deltas_h: PointVec


@dataclass
class ProvingKeyData:
    vk: VerifyingKeyData
    beta_g: PointVec
    a_g: PointVec
    b_g: PointVec
    b_h: PointVec
    h_g: PointVec
    last_delta_g: PointVec
    deltas_abc_g: list          # [PointVec] per stage
    deltas_g: PointVec


@dataclass
class ProvingKeysData:
    circuit_id: str
    serialized_circ_params: bytes
    minimal_proving_keys: dict = field(default_factory=dict)          # representative idx -> ProvingKeyData
    subcircuit_representative_map: dict = field(default_factory=dict)

    def get_pk(self, subcircuit_idx):
        """ProvingKeys::get_pk (mpi-snark/src/data_structures.rs:93-101)."""
        if subcircuit_idx not in self.subcircuit_representative_map:
            raise KeyError("subcircuit index out of range")
        rep = self.subcircuit_representative_map[subcircuit_idx]
        if rep not in self.minimal_proving_keys:
            raise KeyError("missing proving key")
        return self.minimal_proving_keys[rep]

    def num_subcircuits(self):
        return len(self.subcircuit_representative_map)


# ---- decoding ------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf, curve):
        self.b = memoryview(buf)
        self.o = 0
        self.cv = curve
        self.nb = 8 * curve.fq_limbs

    def u64(self):
        if self.o + 8 > len(self.b):
            raise ValueError("key file truncated")
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def raw(self, n):
        if self.o + n > len(self.b):
            raise ValueError("key file truncated")
        v = bytes(self.b[self.o:self.o + n])
        self.o += n
        return v

    def points(self, group, n):
        """n consecutive uncompressed points of G1 / G2."""
        width = (2 if group == 1 else 4) * self.nb
        if self.o + n * width > len(self.b):
            raise ValueError("key file truncated")
        a = np.frombuffer(self.b, dtype=np.uint8, count=n * width, offset=self.o).reshape(n, width).copy()
        self.o += n * width
        if host._is_zcash(self.cv):
            flags = a[:, 0] & 0xE0
            if (flags & 0x80).any():
                raise ValueError("compressed flag set in an uncompressed key file")
            a[:, 0] &= 0x1F
            inf = ((flags & 0x40) != 0).astype(np.uint8)
            # big-endian field elements -> little-endian; G2 stores (c1, c0) pairs -> (c0, c1)
            el = a.reshape(n, width // self.nb, self.nb)[:, :, ::-1]
            if group == 2:
                el = el.reshape(n, 2, 2, self.nb)[:, :, ::-1, :].reshape(n, 4, self.nb)
            a = np.ascontiguousarray(el).reshape(n, width)
        else:
            flags = a[:, -1] & 0xC0
            a[:, -1] &= 0x3F
            inf = ((flags & 0x40) != 0).astype(np.uint8)
        limbs = a.view("<u8").reshape(n, width // 8)
        if inf.any():
            if limbs[inf != 0].any():
                raise ValueError("point at infinity with non-zero coordinates")
        return PointVec(np.ascontiguousarray(limbs), inf)

    def point(self, group):
        return self.points(group, 1)

    def vec(self, group):
        return self.points(group, self.u64())


def _read_vk(r):
    return VerifyingKeyData(r.point(1), r.point(2), r.point(2), r.point(2), r.vec(1), r.vec(2))


def _read_pk(r):
    vk = _read_vk(r)
    beta_g = r.point(1)
    a_g, b_g, b_h, h_g = r.vec(1), r.vec(1), r.vec(2), r.vec(1)
    last_delta_g = r.point(1)
    deltas_abc_g = [r.vec(1) for _ in range(r.u64())]
    deltas_g = r.vec(1)
    return ProvingKeyData(vk, beta_g, a_g, b_g, b_h, h_g, last_delta_g, deltas_abc_g, deltas_g)


def read_proving_key(buf, curve):
    """One `ProvingKey::serialize_uncompressed` blob -> ProvingKeyData."""
    r = _Reader(buf, get_curve(curve))
    pk = _read_pk(r)
    if r.o != len(r.b):
        raise ValueError("trailing bytes after the proving key")
    return pk


def read_proving_keys(buf, curve):
    """A `--key-out` file (ProvingKeys::serialize_uncompressed) -> ProvingKeysData."""
    r = _Reader(buf, get_curve(curve))
    cid = r.raw(r.u64()).decode("utf-8")
    params = r.raw(r.u64())
    out = ProvingKeysData(cid, params)
    for _ in range(r.u64()):
        k = r.u64()
        out.minimal_proving_keys[k] = _read_pk(r)
    for _ in range(r.u64()):
        k = r.u64()
        out.subcircuit_representative_map[k] = r.u64()
    if r.o != len(r.b):
        raise ValueError("trailing bytes after the proving keys")
    return out


def load(path, curve):
    with open(path, "rb") as f:
        return read_proving_keys(f.read(), curve)


# ---- encoding (from Python-integer points, e.g. the oracle's generate_parameters output) -----------------
def _w_vec(out, pts, ser, curve):
    out.write(struct.pack("<Q", len(pts)))
    for p in pts:
        out.write(ser(p, curve))


def write_proving_key(out, key, curve):
    """key: an object with the reference's field names holding integer points (oracle.cpgroth16 key)."""
    cv = get_curve(curve)
    g1, g2 = host.serialize_g1, host.serialize_g2
    out.write(g1(key.alpha_g, cv) + g2(key.beta_h, cv) + g2(key.gamma_h, cv) + g2(key.last_delta_h, cv))
    _w_vec(out, key.gamma_abc_g, g1, cv)
    _w_vec(out, key.deltas_h, g2, cv)
    out.write(g1(key.beta_g, cv))
    _w_vec(out, key.a_g, g1, cv)
    _w_vec(out, key.b_g, g1, cv)
    _w_vec(out, key.b_h, g2, cv)
    _w_vec(out, key.h_g, g1, cv)
    out.write(g1(key.deltas_g[-1], cv))                     # ck.last_delta_g (generator.rs: last delta)
    out.write(struct.pack("<Q", len(key.deltas_abc_g)))
    for v in key.deltas_abc_g:
        _w_vec(out, v, g1, cv)
    _w_vec(out, key.deltas_g, g1, cv)


def write_proving_keys(out, circuit_id, circ_params, minimal_keys, representative_map, curve):
    cid = circuit_id.encode("utf-8")
    out.write(struct.pack("<Q", len(cid)) + cid)
    out.write(struct.pack("<Q", len(circ_params)) + bytes(circ_params))
    out.write(struct.pack("<Q", len(minimal_keys)))
    for k in sorted(minimal_keys):
        out.write(struct.pack("<Q", k))
        write_proving_key(out, minimal_keys[k], curve)
    out.write(struct.pack("<Q", len(representative_map)))
    for k in sorted(representative_map):
        out.write(struct.pack("<QQ", k, representative_map[k]))


def dumps_proving_keys(*a, **kw):
    b = io.BytesIO()
    write_proving_keys(b, *a, **kw)
    return b.getvalue()


# ---- device upload -----------------------------------------------------------------------------------
def _mont(ctx, pv, group):
    """canonical limbs -> Montgomery limbs on the device (hk_field_op TO_MONT over the flat Fq array)."""
    from . import backend as B
    L = ctx.curve.fq_limbs
    if len(pv) == 0:
        return pv.limbs.copy()
    flat = np.ascontiguousarray(pv.limbs.reshape(-1, L))
    return ctx.field_op(B.FIELD_FQ, B.FOP_TO_MONT, flat).reshape(pv.limbs.shape)


def to_device(ctx, pk, staged):
    """ProvingKeyData + the subcircuit's StagedAssignment (matrices, stage ranges) -> host.ProvingKey."""
    up1 = lambda pv: ctx.upload_g1(_mont(ctx, pv, 1), pv.inf)
    up2 = lambda pv: ctx.upload_g2(_mont(ctx, pv, 2), pv.inf)
    h_a, h_b1, h_b2, h_h = up1(pk.a_g), up1(pk.b_g), up2(pk.b_h), up1(pk.h_g)
    h_ck = [up1(v) for v in pk.deltas_abc_g]
    h_d = up1(pk.deltas_g)
    one = lambda pv, g: _mont(ctx, pv, g)[0]
    return host.ProvingKey.from_handles_raw(ctx, staged, h_a, h_b1, h_h, h_b2, h_ck, h_d, one(pk.vk.alpha_g, 1),
                                            one(pk.beta_g, 1), one(pk.vk.beta_h, 2), one(pk.vk.last_delta_h, 2))
