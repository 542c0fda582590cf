"""circom `.r1cs` / witness readers and writers.

Mirrors /root/reference/circom-compat/src/lib.rs -- `R1CSFile::{new, write, generate_constraints}`
(:32-125, :144-157, :380-421), `Header` (:185-255), `read_witness` / `write_witness` (:336-377) and
`read_binary_wtns` (:467-545) -- which is how the reference ingests pre-partitioned circom
subcircuits (`distributed-prover/src/partitioned_r1cs_circuit.rs:65-99`, `node setup-r1cs`).
File format: https://github.com/iden3/r1csfile/blob/master/doc/r1cs_bin_format.md .
Like the reference, only 32-byte fields over the BN254 scalar field are accepted.

Differences from the reference, both deliberate: the wire-to-label map IS read (the reference has the call
commented out at :111 although its own test asserts on it), so `write` round-trips byte-exactly.
"""
import io
import struct

from .curves import BN254

_PRIME_BYTES = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")
assert int.from_bytes(_PRIME_BYTES, "little") == BN254.r


class FormatError(ValueError):
    pass


class Header:
    def __init__(self, field_size, prime_size, n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints):
        self.field_size, self.prime_size = field_size, prime_size
        self.n_wires, self.n_pub_out, self.n_pub_in, self.n_prv_in = n_wires, n_pub_out, n_pub_in, n_prv_in
        self.n_labels, self.n_constraints = n_labels, n_constraints

    @classmethod
    def read(cls, buf, size):
        (field_size,) = struct.unpack_from("<I", buf, 0)
        if field_size != 32:
            raise FormatError("This parser only supports 32-byte fields")
        if size != 32 + field_size:
            raise FormatError("Invalid header section size")
        prime = bytes(buf[4:4 + field_size])
        if prime != _PRIME_BYTES:
            raise FormatError("This parser only supports bn256")
        n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", buf, 4 + field_size)
        return cls(field_size, prime, n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints)

    def write(self):
        return (struct.pack("<I", self.field_size) + self.prime_size +
                struct.pack("<IIIIQI", self.n_wires, self.n_pub_out, self.n_pub_in, self.n_prv_in, self.n_labels,
                            self.n_constraints))


def _sections(data, magic, version):
    if data[:4] != magic:
        raise FormatError("Invalid magic number")
    ver, nsec = struct.unpack_from("<II", data, 4)
    if ver != version:
        raise FormatError("Unsupported version")
    off = 12
    secs = {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", data, off)
        off += 12
        secs[typ] = (off, size)
        off += size
    return ver, secs


class R1CSFile:
    """version, header, constraints [(A, B, C)] with each LC a list of (wire, coeff int), wire_mapping."""

    def __init__(self, version, header, constraints, wire_mapping, witness=None):
        self.version, self.header, self.constraints = version, header, constraints
        self.wire_mapping = wire_mapping
        self.witness = witness or []

    @classmethod
    def read(cls, data):
        data = bytes(data)
        ver, secs = _sections(data, b"r1cs", 1)
        for typ, what in ((1, "header"), (2, "constraint"), (3, "wire2label")):
            if typ not in secs:
                raise FormatError(f"No section offset for {what} type found")
        off, size = secs[1]
        header = Header.read(memoryview(data)[off:off + size], size)
        off, _ = secs[2]
        cons = []
        for _ in range(header.n_constraints):
            lcs = []
            for _ in range(3):
                (n,) = struct.unpack_from("<I", data, off)
                off += 4
                lc = []
                for _ in range(n):
                    (wire,) = struct.unpack_from("<I", data, off)
                    coeff = int.from_bytes(data[off + 4:off + 36], "little")
                    if coeff >= BN254.r:
                        raise FormatError("coefficient is not a canonical field element")
                    lc.append((wire, coeff))
                    off += 36
                lcs.append(lc)
            cons.append(tuple(lcs))
        off, size = secs[3]
        if size != header.n_wires * 8:
            raise FormatError("Invalid map section size")
        wmap = list(struct.unpack_from(f"<{header.n_wires}Q", data, off))
        if wmap and wmap[0] != 0:
            raise FormatError("Wire 0 should always be mapped to 0")
        return cls(ver, header, cons, wmap)

    def write(self):
        def section(typ, body):
            return struct.pack("<IQ", typ, len(body)) + body

        body2 = io.BytesIO()
        for lcs in self.constraints:
            for lc in lcs:
                body2.write(struct.pack("<I", len(lc)))
                for wire, coeff in lc:
                    body2.write(struct.pack("<I", wire) + int(coeff).to_bytes(32, "little"))
        body3 = b"".join(struct.pack("<Q", v) for v in self.wire_mapping)
        return (b"r1cs" + struct.pack("<II", 1, 3) + section(1, self.header.write()) + section(2, body2.getvalue()) +
                section(3, body3))

    # ---- generate_constraints (lib.rs:380-421): wires -> arkworks variables -----------------------------
    def to_staged(self, witness=None, num_stage0=0):
        """Matrices and assignment exactly as the reference's `generate_constraints` leaves them in a fresh
        constraint system: instance = [One, wire_0 .. wire_{num_inputs-1}] with num_inputs = n_pub_in +
        n_pub_out (the circom constant-one wire becomes an ordinary input), witness = the remaining wires.
        Returns a hekaton_system_b200.cp_groth16.StagedAssignment (single stage unless num_stage0 > 0)."""
        from .cp_groth16 import StagedAssignment
        h = self.header
        w = list(witness if witness is not None else self.witness)
        if len(w) != h.n_wires:
            raise FormatError(f"witness has {len(w)} entries, the circuit has {h.n_wires} wires")
        num_inputs = h.n_pub_in + h.n_pub_out
        num_aux = h.n_wires - num_inputs
        col = lambda wire: 1 + wire                      # [One | inputs | aux] is contiguous in wire order
        rows = ([], [], [])
        for lcs in self.constraints:
            for k in range(3):
                rows[k].append([(coeff, col(wire)) for wire, coeff in lcs[k]])
        instance = [1] + [x % BN254.r for x in w[:num_inputs]]
        wit = [x % BN254.r for x in w[num_inputs:]]
        ranges = [(0, num_aux)] if not num_stage0 else [(0, num_stage0), (num_stage0, num_aux)]
        return StagedAssignment(BN254, 1 + num_inputs, rows[0], rows[1], rows[2], instance, wit, ranges)


def read_witness(text):
    """`read_witness` (lib.rs:336-351): snarkjs JSON array of decimal strings, one per line."""
    out = []
    for line in text.splitlines()[1:]:
        if len(line) <= 1:
            continue
        out.append(int(line[2:len(line) - 1]) % BN254.r)
    return out


def write_witness(witness):
    """`write_witness` (lib.rs:353-377)."""
    lines = ["[", f' "{witness[0]}"'] + [f',"{v}"' for v in witness[1:]] + ["]"]
    return "\n".join(lines) + "\n"


def read_binary_wtns(data):
    """`read_binary_wtns` (lib.rs:467-545): snarkjs `.wtns` version 2."""
    data = bytes(data)
    _, secs = _sections(data, b"wtns", 2)
    if 1 not in secs:
        raise FormatError("No section offset for header type found")
    off, size = secs[1]
    (field_size,) = struct.unpack_from("<I", data, off)
    if field_size != 32:
        raise FormatError("This parser only supports 32-byte fields")
    if size != 8 + field_size:
        raise FormatError("Invalid header section size")
    if data[off + 4:off + 36] != _PRIME_BYTES:
        raise FormatError("This parser only supports bn256")
    (n_witness,) = struct.unpack_from("<I", data, off + 36)
    if 2 not in secs:
        raise FormatError("No section offset for constraint type found")
    off, _ = secs[2]
    return [int.from_bytes(data[off + 32 * i:off + 32 * (i + 1)], "little") for i in range(n_witness)]
