"""Marshalling between Python integers and the limb arrays that cross the C ABI
(little-endian u64 limbs, Montgomery form where the ABI says so)."""
import numpy as np


def ints_to_limbs(vals, nlimbs):
    """list of non-negative ints -> np.uint64 array (len, nlimbs), little-endian limbs."""
    nbytes = 8 * nlimbs
    buf = b"".join(int(v).to_bytes(nbytes, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(len(vals), nlimbs).copy()


def limbs_to_ints(arr):
    arr = np.ascontiguousarray(arr, dtype="<u8")
    nlimbs = arr.shape[-1]
    flat = arr.reshape(-1, nlimbs)
    raw = flat.tobytes()
    nbytes = 8 * nlimbs
    return [int.from_bytes(raw[i * nbytes:(i + 1) * nbytes], "little") for i in range(flat.shape[0])]


def to_mont(vals, mod, nlimbs):
    R = 1 << (64 * nlimbs)
    return ints_to_limbs([(v % mod) * R % mod for v in vals], nlimbs)


def from_mont(arr, mod):
    nlimbs = arr.shape[-1]
    Rinv = pow(1 << (64 * nlimbs), -1, mod)
    return [v * Rinv % mod for v in limbs_to_ints(arr)]


# ---- points -------------------------------------------------------------------------------
def g1_affine_to_limbs(points, curve):
    """[(x, y) | None] -> (array (n, 2*L) Montgomery, infinity flags uint8)."""
    L = curve.fq_limbs
    vals, inf = [], np.zeros(len(points), dtype=np.uint8)
    for i, P in enumerate(points):
        if P is None:
            vals += [0, 0]
            inf[i] = 1
        else:
            vals += [P[0], P[1]]
    return to_mont(vals, curve.p, L).reshape(len(points), 2 * L), inf


def g2_affine_to_limbs(points, curve):
    L = curve.fq_limbs
    vals, inf = [], np.zeros(len(points), dtype=np.uint8)
    for i, P in enumerate(points):
        if P is None:
            vals += [0, 0, 0, 0]
            inf[i] = 1
        else:
            vals += [P[0][0], P[0][1], P[1][0], P[1][1]]
    return to_mont(vals, curve.p, L).reshape(len(points), 4 * L), inf


def g1_affine_from_limbs(arr, curve, inf=None):
    L = curve.fq_limbs
    arr = np.asarray(arr).reshape(-1, 2, L)
    v = from_mont(arr.reshape(-1, L), curve.p)
    out = []
    for i in range(arr.shape[0]):
        x, y = v[2 * i], v[2 * i + 1]
        if (inf is not None and inf[i]) or (x == 0 and y == 0):
            out.append(None)
        else:
            out.append((x, y))
    return out


def g2_affine_from_limbs(arr, curve, inf=None):
    L = curve.fq_limbs
    arr = np.asarray(arr).reshape(-1, 4, L)
    v = from_mont(arr.reshape(-1, L), curve.p)
    out = []
    for i in range(arr.shape[0]):
        x0, x1, y0, y1 = v[4 * i:4 * i + 4]
        if (inf is not None and inf[i]) or (x0 == 0 and x1 == 0 and y0 == 0 and y1 == 0):
            out.append(None)
        else:
            out.append(((x0, x1), (y0, y1)))
    return out


def g1_jacobian_from_limbs(arr, curve):
    """(n, 3*L) Montgomery -> [(X, Y, Z)] canonical ints."""
    L = curve.fq_limbs
    v = from_mont(np.asarray(arr).reshape(-1, L), curve.p)
    return [tuple(v[3 * i:3 * i + 3]) for i in range(len(v) // 3)]


def g2_jacobian_from_limbs(arr, curve):
    L = curve.fq_limbs
    v = from_mont(np.asarray(arr).reshape(-1, L), curve.p)
    return [((v[6 * i], v[6 * i + 1]), (v[6 * i + 2], v[6 * i + 3]), (v[6 * i + 4], v[6 * i + 5]))
            for i in range(len(v) // 6)]
