"""CPU oracle for the Hekaton CP-Groth16 hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, with Python big integers, the arithmetic the reference
executes on the per-subcircuit commit + prove path
(/root/reference/cp-groth16/src/{committer.rs:55-122, prover.rs:53-171,
generator.rs:18-238, verifier.rs:23-71}) and the arkworks-0.4 semantics those
files call into (ark-ec VariableBaseMSM, ark-poly Radix2EvaluationDomain,
ark-groth16 LibsnarkReduction).  The arkworks crates are third-party and are
NOT vendored in /root/reference (versions "0.4", no Cargo.lock), so their
published algorithms are restated here and in oracle/c/.

PARITY PINNING STATUS: "parity unpinned" against the reference's own tests --
the reference holds no golden vector, known-answer test or numeric fixture for
MSM, NTT or proofs (all 33 #[test]s are boolean property tests; SURVEY.md
section 4 / 8c), and the reference cannot be built here (no Rust toolchain).
What the oracle IS pinned against (tests/test_oracle_*.py):
  * public curve constants: the standard BLS12-381 / BN254 G1 and G2
    generators lie on their curves and have order r;
  * the algebraic properties the reference's tests check: setup -> commit ->
    prove -> verify_proof accepts, tampered proofs are rejected (reduced ate
    pairing + verifier.rs:23-71 restated in oracle/pairing.py, itself pinned
    by bilinearity / non-degeneracy / order-r on the standard generators:
    tests/test_oracle_pairing.py);
  * RFC 7539 / ECRYPT ChaCha block vectors and the zcash-format BLS12-381
    generator encodings for the host-side RNG and serialisers
    (tests/test_rng.py, tests/test_serialize.py);
  * an MSM/NTT-free trapdoor known-answer for whole proofs (cpgroth16.py);
  * the reference's own two acceptance tests (cp-groth16/src/lib.rs:139-180,
    273-313) restated on its PolyEvalCircuit (reference_circuits.py,
    tests/test_reference_poly_commit.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (hekaton_system_b200)
never does.
"""
