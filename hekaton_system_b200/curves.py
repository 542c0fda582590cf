"""Curve constants the host side needs for marshalling (limb counts, moduli, generators).
ark-bls12-381 0.4 / ark-bn254 0.4 parameters; the device constants are generated from the
same numbers by tools/gen_field.py."""
from dataclasses import dataclass


@dataclass(frozen=True)
class CurveInfo:
    name: str
    cid: int
    p: int
    r: int
    fq_limbs: int          # u64 limbs per Fq element
    fr_limbs: int
    fr_bits: int
    two_adicity: int
    fr_generator: int
    g1: tuple
    g2: tuple

    @property
    def fq_R(self):
        return 1 << (64 * self.fq_limbs)

    @property
    def fr_R(self):
        return 1 << (64 * self.fr_limbs)


BLS12_381 = CurveInfo(
    name="bls12_381", cid=0,
    p=0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    r=0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
    fq_limbs=6, fr_limbs=4, fr_bits=255, two_adicity=32, fr_generator=7,
    g1=(0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
        0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    g2=((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
         0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
        (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
         0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
)

BN254 = CurveInfo(
    name="bn254", cid=1,
    p=21888242871839275222246405745257275088696311157297823662689037894645226208583,
    r=21888242871839275222246405745257275088548364400416034343698204186575808495617,
    fq_limbs=4, fr_limbs=4, fr_bits=254, two_adicity=28, fr_generator=5,
    g1=(1, 2),
    g2=((10857046999023057135944570762232829481370756359578518086990519993285655852781,
         11559732032986387107991004021392285783925812861821192530917403151452391805634),
        (8495653923123431417604973247489272438418190587263600148770280649306958101930,
         4082367875863433681332203403145435568316851327593401208105741076214120093531)),
)

_BY = {"bls12_381": BLS12_381, "bn254": BN254, 0: BLS12_381, 1: BN254}


def get_curve(c):
    if isinstance(c, CurveInfo):
        return c
    return _BY[c]
