"""hekaton_system_b200 -- B200-native (sm_100a) backend for Hekaton's per-subcircuit
CP-Groth16 hot path (MSM over G1/G2, Fr NTT, R1CS->QAP witness map, fused commit/prove).

The compute lives in lib/libhekaton_b200.so (CUDA, C ABI: include/hekaton_b200.h).  This
package is the ctypes binding plus host-side marshalling and a mirror of the reference's
cp-groth16 operator interface.  There is no CPU fallback: importing `backend` without the
built library, or creating a Context without an sm_100 GPU, raises.
"""
from .curves import BLS12_381, BN254, get_curve  # noqa: F401
