"""The C++ host mirror (host/cp_groth16.hpp: MultiStageConstraintSystem, CommitmentBuilder::{commit,
prove}, Proof) driven like `node work` (host/hk_node.cpp: worker threads on forked contexts):
every commitment and proof must equal the oracle's trapdoor known answer byte for byte."""
import json
import os
import subprocess

import numpy as np
import pytest

from hekaton_system_b200 import codec, workload
from oracle.ec import groups

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cname", ["bls12_381", "bn254"])
def test_cpp_driver_commit_prove(tmp_path, cname):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    wl = workload.make_workload(cname, "tiny", seed=6, num_assignments=5)
    td, _ = workload.key_scalars(wl, seed=8)
    cv = wl.curve
    G1, G2 = groups(cname)
    G = G1.mul(G1.gen, td["g1"])
    H = G2.mul(G2.gen, td["g2"])
    subs = []
    for i, z in enumerate(wl.assignments):
        kap, r_, s_ = 1000 + i, cv.r - 77 - i, 0xabcdef0123 + i
        e = workload.expected_exponents(wl, z, r_, s_, [kap])
        fr = lambda v: codec.to_mont([v], cv.r, 4)[0]
        subs.append(dict(z=codec.to_mont(z, cv.r, 4), kappa=fr(kap), r=fr(r_), s=fr(s_),
                         D=codec.g1_affine_to_limbs([G1.mul(G, e["D"][0])], cv)[0],
                         A=codec.g1_affine_to_limbs([G1.mul(G, e["A"])], cv)[0],
                         B=codec.g2_affine_to_limbs([G2.mul(H, e["B"])], cv)[0],
                         C=codec.g1_affine_to_limbs([G1.mul(G, e["C"])], cv)[0]))
    path = str(tmp_path / "wl.bin")
    workload.dump_workload(wl, path, subs)
    out = subprocess.run([os.path.join(ROOT, "host", "hk_node"), path, "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["subcircuits"] == 5 and res["mismatches"] == 0 and res["gpu_launches"] > 0
