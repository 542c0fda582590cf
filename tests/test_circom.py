"""circom file formats, pinned against the reference's own fixtures (tests/golden/, extracted verbatim from
/root/reference/circom-compat/src/lib.rs:548-737).  These restate the reference tests `sample`, `test_write`
and `wtns_bin_file` assertion by assertion."""
import os

import pytest

from hekaton_system_b200 import circom
from hekaton_system_b200.curves import BN254

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DATA = bytes.fromhex(open(os.path.join(G, "circom_sample.r1cs.hex")).read().strip())
WTNS = bytes.fromhex(open(os.path.join(G, "circom_sample.wtns.hex")).read().strip())
WITNESS_FILE = open(os.path.join(G, "circom_sample_witness.json")).read()


def test_sample():                      # lib.rs:548-644
    f = circom.R1CSFile.read(DATA)
    assert f.version == 1
    assert f.header.field_size == 32
    assert f.header.prime_size == bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")
    assert (f.header.n_wires, f.header.n_pub_out, f.header.n_pub_in, f.header.n_prv_in) == (7, 1, 2, 3)
    assert f.header.n_labels == 0x03E8 and f.header.n_constraints == 3
    assert len(f.constraints) == 3
    assert len(f.constraints[0][0]) == 2
    assert f.constraints[0][0][0] == (5, 3)
    assert f.constraints[2][1][0] == (0, 6)
    assert len(f.constraints[1][2]) == 0
    assert len(f.wire_mapping) == 7 and f.wire_mapping[1] == 3
    w = circom.read_witness(WITNESS_FILE)
    assert len(w) == 5 and w[0] == 1 and w[4] == 0


def test_write():                       # lib.rs:646-720: byte-exact round trips
    f = circom.R1CSFile.read(DATA)
    assert f.write() == DATA
    w = circom.read_witness(WITNESS_FILE)
    assert circom.write_witness(w) == WITNESS_FILE


def test_wtns_bin_file():               # lib.rs:722-737
    assert circom.read_binary_wtns(WTNS) == [1]


def test_rejects_what_the_reference_rejects():
    with pytest.raises(circom.FormatError, match="magic"):
        circom.R1CSFile.read(b"xxxx" + DATA[4:])
    with pytest.raises(circom.FormatError, match="version"):
        circom.R1CSFile.read(DATA[:4] + (2).to_bytes(4, "little") + DATA[8:])
    bad_prime = bytearray(DATA)
    bad_prime[28] ^= 1
    with pytest.raises(circom.FormatError, match="bn256"):
        circom.R1CSFile.read(bytes(bad_prime))


def test_generate_constraints_mapping():
    """lib.rs:380-421: inputs = first n_pub_in + n_pub_out wires (the constant wire included), the rest witness."""
    f = circom.R1CSFile.read(DATA)
    wires = [1, 11, 22, 33, 44, 55, 66]
    st = f.to_staged(wires)
    assert st.l == 4 and st.instance == [1, 1, 11, 22] and st.witness == [33, 44, 55, 66]
    a0 = st.rows[0][0]                     # 3*w5 + 8*w6  -> columns 1 + wire
    assert a0 == [(3, 6), (8, 7)]
    assert st.stage_ranges == [(0, 4)]
    with pytest.raises(circom.FormatError):
        f.to_staged(wires[:5])
    assert all(cf < BN254.r for row in st.rows[1] for cf, _ in row)
