"""Curve / field parameters (oracle; test infrastructure only).

Values follow ark-bls12-381 0.4 and ark-bn254 0.4 (third-party crates named in
/root/reference/cp-groth16/Cargo.toml:9-23 and distributed-prover/Cargo.toml:12-13;
not vendored).  Everything derivable is derived here and re-checked by
tests/test_oracle_params.py (on-curve, order-r, root-of-unity order).
"""
from dataclasses import dataclass, field


@dataclass(frozen=True)
class Curve:
    name: str
    cid: int                 # id used by the C ABI (0 = BLS12-381, 1 = BN254)
    p: int                   # base field modulus
    r: int                   # scalar field modulus
    fq_limbs64: int          # arkworks BigInt<N> limb count for Fq
    fr_limbs64: int
    b1: int                  # G1: y^2 = x^3 + b1
    b2: tuple                # G2: y^2 = x^3 + b2   (Fq2, u^2 = -1)
    g1: tuple                # standard generators (affine)
    g2: tuple
    fr_generator: int        # F::GENERATOR (multiplicative generator, coset offset)
    two_adicity: int
    fr_bits: int             # MODULUS_BIT_SIZE of Fr

    @property
    def fq_R(self):
        return 1 << (64 * self.fq_limbs64)

    @property
    def fr_R(self):
        return 1 << (64 * self.fr_limbs64)

    @property
    def two_adic_root(self):
        # TWO_ADIC_ROOT_OF_UNITY = GENERATOR^((r-1)/2^s)
        return pow(self.fr_generator, (self.r - 1) >> self.two_adicity, self.r)


BLS12_381 = Curve(
    name="bls12_381",
    cid=0,
    p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    fq_limbs64=6,
    fr_limbs64=4,
    b1=4,
    b2=(4, 4),
    g1=(
        0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1,
    ),
    g2=(
        (
            0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
            0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e,
        ),
        (
            0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
            0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be,
        ),
    ),
    fr_generator=7,
    two_adicity=32,
    fr_bits=255,
)

_BN_P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_BN_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _bn254_b2():
    # b2 = 3 / (9 + u) in Fq2 = Fq[u]/(u^2+1)
    # 1/(9+u) = (9-u)/(81+1)
    inv82 = pow(82, -1, _BN_P)
    return (3 * 9 * inv82 % _BN_P, (-3 * inv82) % _BN_P)


BN254 = Curve(
    name="bn254",
    cid=1,
    p=_BN_P,
    r=_BN_R,
    fq_limbs64=4,
    fr_limbs64=4,
    b1=3,
    b2=_bn254_b2(),
    g1=(1, 2),
    g2=(
        (
            10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634,
        ),
        (
            8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531,
        ),
    ),
    fr_generator=5,
    two_adicity=28,
    fr_bits=254,
)

CURVES = {"bls12_381": BLS12_381, "bn254": BN254, 0: BLS12_381, 1: BN254}


def get_curve(c):
    if isinstance(c, Curve):
        return c
    return CURVES[c]
