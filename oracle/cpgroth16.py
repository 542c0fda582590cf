"""CP-Groth16 ("Mirage") setup / commit / prove / trapdoor known-answer
(oracle; test infrastructure only).

Follows, line by line where it matters for VALUES:
  generate_parameters   /root/reference/cp-groth16/src/generator.rs:18-238
  CommitmentBuilder     /root/reference/cp-groth16/src/committer.rs:55-122
  prove_last_stage      /root/reference/cp-groth16/src/prover.rs:53-171
  MultiStageConstraintSystem (assignment order, stage ranges)
                        /root/reference/cp-groth16/src/constraint_synthesizer.rs:53-106
Randomness (alpha, beta, gamma, deltas, tau, generators, kappa, r, s) is passed
in explicitly: the GPU boundary never touches the RNG.
"""
import random
from dataclasses import dataclass, field

from .params import get_curve
from .ec import groups
from . import poly


@dataclass
class Trapdoor:
    alpha: int
    beta: int
    gamma: int
    deltas: list
    tau: int
    g1_scalar: int   # generator G = g1_scalar * (standard G1 generator)   (generator.rs:35)
    g2_scalar: int


@dataclass
class ProvingKey:
    """Mirror of cp-groth16 data_structures.rs:66-114 (ProvingKey + CommitterKey + VerifyingKey)."""
    curve: object
    # vk
    alpha_g: tuple = None
    beta_h: tuple = None
    gamma_h: tuple = None
    last_delta_h: tuple = None
    gamma_abc_g: list = None
    deltas_h: list = None
    # pk
    beta_g: tuple = None
    a_g: list = None
    b_g: list = None
    b_h: list = None
    h_g: list = None
    deltas_g: list = None
    # ck
    last_delta_g: tuple = None
    deltas_abc_g: list = None
    # exponents of every base w.r.t. G / H (trapdoor view; for the KAT only)
    exps: dict = field(default_factory=dict)


@dataclass
class StagedCircuit:
    """What `generate_constraints` leaves in a MultiStageConstraintSystem."""
    r1cs: poly.R1CS
    stage_ranges: list     # witness index ranges [(start, end), ...] contiguous (constraint_synthesizer.rs:21,53-64)
    instance: list         # instance assignment incl. leading 1
    witness: list

    def full_assignment(self):          # constraint_synthesizer.rs:102-106
        return list(self.instance) + list(self.witness)

    def stage_witness(self, i):         # constraint_synthesizer.rs:96-99
        s, e = self.stage_ranges[i]
        return self.witness[s:e]


def synthetic_circuit(curve, num_constraints, num_instance=4, stage0_witnesses=8,
                      seed=1, bool_frac=0.5):
    """A satisfiable 2-stage R1CS with big-merkle-like statistics: `bool_frac` of the
    gates are booleanity checks (0/1 witnesses, like SHA-256 bit decompositions), the
    rest are random fan-in-2..3 multiplication gates.  One new witness per constraint."""
    c = get_curve(curve)
    r = c.r
    rng = random.Random(seed)
    l = num_instance
    instance = [1] + [rng.randrange(r) for _ in range(l - 1)]
    witness = []
    # stage 0: free witnesses (committed values), half of them bits
    for i in range(stage0_witnesses):
        witness.append(rng.randrange(2) if rng.random() < bool_frac else rng.randrange(r))
    a_rows, b_rows, c_rows = [], [], []
    z = instance + witness

    def lc(nterms):
        terms = []
        for _ in range(nterms):
            j = rng.randrange(len(z))
            cf = rng.choice([1, 1, 1, 2, r - 1, rng.randrange(r)])
            terms.append((cf, j))
        return terms

    for k in range(num_constraints):
        if rng.random() < bool_frac:
            # new bit b: b * (1 - b) = 0  ->  A = b, B = 1 - b, C = 0
            bit = rng.randrange(2)
            z.append(bit)
            j = len(z) - 1
            a_rows.append([(1, j)])
            b_rows.append([(1, 0), (r - 1, j)])
            c_rows.append([])
        else:
            ra = lc(rng.randrange(1, 4))
            rb = lc(rng.randrange(1, 4))
            va = sum(cf * z[j] for cf, j in ra) % r
            vb = sum(cf * z[j] for cf, j in rb) % r
            z.append(va * vb % r)
            j = len(z) - 1
            a_rows.append(ra)
            b_rows.append(rb)
            c_rows.append([(1, j)])
    witness = z[l:]
    r1cs = poly.R1CS(l, len(witness), a_rows, b_rows, c_rows)
    ranges = [(0, stage0_witnesses), (stage0_witnesses, len(witness))]
    return StagedCircuit(r1cs, ranges, instance, witness)


def sample_trapdoor(curve, num_stages, seed):
    c = get_curve(curve)
    rng = random.Random(seed)
    nz = lambda: rng.randrange(1, c.r)
    return Trapdoor(nz(), nz(), nz(), [nz() for _ in range(num_stages)], nz(), nz(), nz())


def generate_parameters(curve, circ, td, with_points=True):
    """generator.rs:18-238 with explicit toxic waste `td`.  Returns a ProvingKey whose
    `.exps` holds the discrete logs of every base (w.r.t. G resp. H)."""
    c = get_curve(curve)
    r = c.r
    G1, G2 = groups(c)
    r1cs = circ.r1cs
    l = r1cs.num_instance
    a, b, cc, zt, qap_num_variables, m_raw = poly.instance_map_with_evaluation(c, r1cs, td.tau)
    inv = lambda x: pow(x, -1, r)
    alpha, beta, gamma = td.alpha, td.beta, td.gamma
    abc = [(beta * a[i] + alpha * b[i] + cc[i]) % r for i in range(r1cs.num_variables)]
    deltas_abc = []
    for delta, (s, e) in zip(td.deltas, circ.stage_ranges):          # generator.rs:93-106
        dinv = inv(delta)
        deltas_abc.append([abc[i] * dinv % r for i in range(s + l, e + l)])
    gamma_abc = [abc[i] * inv(gamma) % r for i in range(l)]         # generator.rs:112-117
    last_delta_inv = inv(td.deltas[-1])
    h = poly.h_query_scalars(c, m_raw - 1, td.tau, zt, last_delta_inv)  # generator.rs:182

    exps = dict(alpha_g=alpha, beta_g=beta, beta_h=beta, gamma_h=gamma,
                deltas_g=list(td.deltas), deltas_h=list(td.deltas),
                gamma_abc_g=gamma_abc, a_g=a, b_g=b, b_h=b, h_g=h,
                deltas_abc_g=deltas_abc, g1=td.g1_scalar, g2=td.g2_scalar)
    pk = ProvingKey(curve=c, exps=exps)
    if not with_points:
        return pk
    G = G1.mul(G1.gen, td.g1_scalar)
    H = G2.mul(G2.gen, td.g2_scalar)
    m1 = lambda k: G1.mul(G, k % r)
    m2 = lambda k: G2.mul(H, k % r)
    pk.alpha_g = m1(alpha)
    pk.beta_g = m1(beta)
    pk.beta_h = m2(beta)
    pk.gamma_h = m2(gamma)
    pk.deltas_g = [m1(d) for d in td.deltas]
    pk.deltas_h = [m2(d) for d in td.deltas]
    pk.last_delta_g = pk.deltas_g[-1]
    pk.last_delta_h = pk.deltas_h[-1]
    pk.gamma_abc_g = [m1(x) for x in gamma_abc]
    pk.a_g = [m1(x) for x in a]
    pk.b_g = [m1(x) for x in b]
    pk.b_h = [m2(x) for x in b]
    pk.h_g = [m1(x) for x in h]
    pk.deltas_abc_g = [[m1(x) for x in v] for v in deltas_abc]
    return pk


def commit(pk, circ, stage, kappa):
    """committer.rs:85-91: MSM(ck_stage, w_stage) + kappa * last_delta_g  -> affine."""
    G1, _ = groups(pk.curve)
    w = circ.stage_witness(stage)
    ck = pk.deltas_abc_g[stage]
    assert len(w) == len(ck)
    return G1.add(G1.msm_naive(ck, w), G1.mul(pk.last_delta_g, kappa))


def _calculate_coeff(G, initial, query, vk_param, assignment):
    """prover.rs:158-171."""
    acc = G.msm_naive(query[1:], assignment)
    return G.add(G.add(G.add(initial, query[0]), acc), vk_param)


def prove(pk, circ, r_, s_, kappas):
    """prover.rs:53-156 followed by committer.rs:106-114.  Returns (A, B, C) affine."""
    c = pk.curve
    G1, G2 = groups(c)
    rr = c.r
    z = circ.full_assignment()
    assignment = z[1:]                                               # prover.rs:78-82
    a_g = _calculate_coeff(G1, G1.mul(pk.last_delta_g, r_), pk.a_g, pk.alpha_g, assignment)
    if r_ % rr == 0:                                                 # prover.rs:92
        b_g = None
    else:
        b_g = _calculate_coeff(G1, G1.mul(pk.last_delta_g, s_), pk.b_g, pk.beta_g, assignment)
    b_h = _calculate_coeff(G2, G2.mul(pk.last_delta_h, s_), pk.b_h, pk.beta_h, assignment)
    w_last = circ.stage_witness(len(circ.stage_ranges) - 1)
    l_aux = G1.msm_naive(pk.deltas_abc_g[-1], w_last)                # prover.rs:111-117
    h = poly.witness_map(c, circ.r1cs, z)                            # prover.rs:123
    assert len(h) == len(pk.h_g) + 1                                 # prover.rs:128
    h_acc = G1.msm_naive(pk.h_g, h[:len(pk.h_g)])
    r_s_delta = G1.mul(pk.last_delta_g, r_ * s_ % rr)
    c_g = G1.mul(a_g, s_)
    c_g = G1.add(c_g, G1.mul(b_g, r_))
    c_g = G1.sub(c_g, r_s_delta)
    c_g = G1.add(c_g, l_aux)
    c_g = G1.add(c_g, h_acc)
    # committer.rs:112-114: msm_unchecked(deltas_g, comm_rands) zips to the shorter
    c_g = G1.sub(c_g, G1.msm_naive(pk.deltas_g, kappas))
    return a_g, b_h, c_g


def trapdoor_expected(pk, circ, td, r_, s_, kappas):
    """MSM/NTT-free known answer: exponents of D_i, A, B, C from O(m+N)
    scalar arithmetic, then ONE scalar-mul each.  h(tau) is evaluated from the
    defining identity a(x)b(x) - c(x) = h(x) Z(x) at x = tau, never via an NTT."""
    c = pk.curve
    rr = c.r
    G1, G2 = groups(c)
    e = pk.exps
    z = circ.full_assignment()
    l = circ.r1cs.num_instance
    inv = lambda x: pow(x, -1, rr)
    a_tau = sum(zi * ai for zi, ai in zip(z, e["a_g"])) % rr
    b_tau = sum(zi * bi for zi, bi in zip(z, e["b_g"])) % rr
    a_exp = (td.alpha + a_tau + r_ * td.deltas[-1]) % rr
    b_exp = (td.beta + b_tau + s_ * td.deltas[-1]) % rr
    # c(tau) from the instance map's c column; h(tau) = (a(tau) b(tau) - c(tau)) / Z(tau)
    _, _, ccol, zt, _, _ = poly.instance_map_with_evaluation(c, circ.r1cs, td.tau)
    c_tau = sum(zi * ci for zi, ci in zip(z, ccol)) % rr
    h_tau = (a_tau * b_tau - c_tau) % rr * inv(zt) % rr
    ds = []
    for i, kappa in enumerate(kappas):
        s0, e0 = circ.stage_ranges[i]
        w = circ.witness[s0:e0]
        d = (sum(wi * x for wi, x in zip(w, e["deltas_abc_g"][i])) + kappa * td.deltas[-1]) % rr
        ds.append(d)
    sl, el = circ.stage_ranges[-1]
    l_exp = sum(wi * x for wi, x in zip(circ.witness[sl:el], e["deltas_abc_g"][-1])) % rr
    h_exp = h_tau * zt % rr * inv(td.deltas[-1]) % rr
    c_exp = (s_ * a_exp + r_ * b_exp - r_ * s_ * td.deltas[-1] + l_exp + h_exp
             - sum(k * d for k, d in zip(kappas, td.deltas))) % rr
    G = G1.mul(G1.gen, td.g1_scalar)
    H = G2.mul(G2.gen, td.g2_scalar)
    return dict(A=G1.mul(G, a_exp), B=G2.mul(H, b_exp), C=G1.mul(G, c_exp),
                D=[G1.mul(G, d) for d in ds],
                exps=dict(A=a_exp, B=b_exp, C=c_exp, D=ds))
