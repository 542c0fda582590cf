"""ctypes binding of oracle/_build/liboracle.so (the C restatement; oracle/c/oracle.c).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py for who may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


class OrcPk(C.Structure):
    _fields_ = [
        ("n", C.c_size_t), ("l", C.c_size_t), ("nw", C.c_size_t),
        ("a_ptr", C.c_void_p), ("a_col", C.c_void_p), ("b_ptr", C.c_void_p), ("b_col", C.c_void_p),
        ("c_ptr", C.c_void_p), ("c_col", C.c_void_p),
        ("a_val", C.c_void_p), ("b_val", C.c_void_p), ("c_val", C.c_void_p),
        ("num_stages", C.c_uint32), ("stage_start", C.c_void_p),
        ("a_g", C.c_void_p), ("b_g", C.c_void_p), ("b_h", C.c_void_p), ("h_g", C.c_void_p),
        ("ck", C.POINTER(C.c_void_p)), ("deltas_g", C.c_void_p),
        ("alpha_g", C.c_void_p), ("beta_g", C.c_void_p), ("beta_h", C.c_void_p), ("last_delta_h", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(["make", "-C", os.path.join(HERE, "c")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(LIB)
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def ntt(curve_id, data, log_n, inverse=False, coset=False, threads=1):
    """In place on a (2^log_n, 4) u64 Montgomery array."""
    assert data.dtype == np.uint64 and data.flags["C_CONTIGUOUS"]
    rc = lib().orc_ntt(C.c_int(curve_id), _p(data), C.c_int(log_n), C.c_int(1 if inverse else 0),
                       C.c_int(1 if coset else 0), C.c_int(threads))
    assert rc == 0
    return data


def msm(curve_id, group, bases, inf, scalars, repr_=0, threads=1):
    """bases (n, 2L|4L) Montgomery, scalars (n, 4) -> (affine xy Montgomery array, inf flag)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    n = scalars.shape[0] if scalars.size else 0
    n = min(n, bases.shape[0]) if bases.size else 0
    width = bases.shape[1] if bases.ndim == 2 and bases.size else ((12 if curve_id == 0 else 8) * group)
    out = np.zeros(width, dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8)
    rc = lib().orc_msm(C.c_int(curve_id), C.c_int(group), _p(bases), _p(inf), _p(scalars), C.c_size_t(n),
                       C.c_int(repr_), _p(out), _p(oinf), C.c_int(threads))
    assert rc == 0
    return out, bool(oinf[0])


def field_mul(curve_id, field, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    lib().orc_field_mul(C.c_int(curve_id), C.c_int(field), _p(a), _p(b), _p(out), C.c_size_t(a.shape[0]))
    return out


def witness_map(curve_id, n, l, csr, z, threads=1):
    """csr = [(ptr u32, col u32, val u64[nnz,4])] * 3; z (m, 4) Montgomery -> h (N, 4)."""
    log_n = max(0, (n + l - 1).bit_length())
    h = np.zeros((1 << log_n, 4), dtype=np.uint64)
    args = []
    keep = []
    for ptr, col, val in csr:
        ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
        col = np.ascontiguousarray(col, dtype=np.uint32)
        val = np.ascontiguousarray(val, dtype=np.uint64)
        keep += [ptr, col, val]
        args += [_p(ptr), _p(col), _p(val)]
    z = np.ascontiguousarray(z, dtype=np.uint64)
    rc = lib().orc_witness_map(C.c_int(curve_id), C.c_size_t(n), C.c_size_t(l), *args, _p(z), _p(h), C.c_int(threads))
    assert rc == 0
    return h


class CpuProver:
    """Holds host copies of a proving key + matrices and runs the restated prover on the CPU."""

    def __init__(self, curve_id, n, l, nw, csr, stage_start, a_g, b_g, b_h, h_g, ck, deltas_g, alpha_g, beta_g,
                 beta_h, last_delta_h):
        self.curve_id = curve_id
        self.L = 6 if curve_id == 0 else 4
        u64 = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
        self.keep = []
        pk = OrcPk()
        pk.n, pk.l, pk.nw = n, l, nw
        names = [("a_ptr", "a_col", "a_val"), ("b_ptr", "b_col", "b_val"), ("c_ptr", "c_col", "c_val")]
        for (pn, cn, vn), (ptr, col, val) in zip(names, csr):
            ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
            col = np.ascontiguousarray(col, dtype=np.uint32)
            val = u64(val)
            self.keep += [ptr, col, val]
            setattr(pk, pn, ptr.ctypes.data)
            setattr(pk, cn, col.ctypes.data)
            setattr(pk, vn, val.ctypes.data)
        ss = np.ascontiguousarray(stage_start, dtype=np.uint32)
        self.keep.append(ss)
        pk.num_stages = len(stage_start) - 1
        pk.stage_start = ss.ctypes.data
        for nm, arr in (("a_g", a_g), ("b_g", b_g), ("b_h", b_h), ("h_g", h_g), ("deltas_g", deltas_g),
                        ("alpha_g", alpha_g), ("beta_g", beta_g), ("beta_h", beta_h), ("last_delta_h", last_delta_h)):
            arr = u64(arr)
            self.keep.append(arr)
            setattr(pk, nm, arr.ctypes.data)
        cks = [u64(c) for c in ck]
        self._ck = cks
        self._deltas = u64(deltas_g)
        self._stage_start = [int(x) for x in stage_start]
        self.keep += cks
        self.ck_arr = (C.c_void_p * len(cks))(*[c.ctypes.data for c in cks])
        pk.ck = C.cast(self.ck_arr, C.POINTER(C.c_void_p))
        self.pk = pk

    def commit(self, stage, z, kappa, threads=1):
        """committer.rs:85-91 on the CPU: MSM(ck_stage, w_stage) + kappa * last_delta_g, evaluated as one
        MSM over [ck.., delta_last] with scalars [w.., kappa] (same group element)."""
        L = self.L
        ss = self._stage_start
        l = self.pk.l
        ck = self._ck[stage]
        w = np.ascontiguousarray(z, dtype=np.uint64).reshape(-1, 4)[l + ss[stage]:l + ss[stage + 1]]
        bases = np.concatenate([ck.reshape(-1, 2 * L), self._deltas.reshape(-1, 2 * L)[-1:]])
        sc = np.concatenate([w, np.ascontiguousarray(kappa, dtype=np.uint64).reshape(1, 4)])
        return msm(self.curve_id, 1, bases, None, sc, repr_=1, threads=threads)

    def prove(self, z, r, s, kappas, threads=1):
        L = self.L
        z = np.ascontiguousarray(z, dtype=np.uint64)
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        kap = np.ascontiguousarray(kappas, dtype=np.uint64) if kappas is not None and len(kappas) else np.zeros((1, 4), dtype=np.uint64)
        a = np.zeros(2 * L, dtype=np.uint64)
        b = np.zeros(4 * L, dtype=np.uint64)
        c = np.zeros(2 * L, dtype=np.uint64)
        rc = lib().orc_prove(C.c_int(self.curve_id), C.byref(self.pk), _p(z), _p(r), _p(s), _p(kap), _p(a), _p(b),
                             _p(c), C.c_int(threads))
        assert rc == 0
        return a, b, c


def dot_fr(curve_id, a, b, threads=0):
    """sum_i a_i * b_i mod r for (n, 4) canonical limb arrays -> Python int."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    assert a.shape == b.shape and a.shape[1] == 4
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().orc_dot_fr(C.c_int(curve_id), _p(a), _p(b), C.c_size_t(a.shape[0]), _p(out),
                          C.c_int(threads or (os.cpu_count() or 1)))
    assert rc == 0
    return sum(int(out[i]) << (64 * i) for i in range(4))


def uses_asm():
    """True when the MULX/ADCX/ADOX field multiplication is in use on this host."""
    return bool(lib().orc_uses_asm())


def bench_field_mul(curve_id, field, iters=2_000_000):
    """ns per Montgomery multiplication (single thread, dependent chain); field 0 = Fr, 1 = Fq."""
    fn = lib().orc_bench_field_mul
    fn.restype = C.c_double
    return float(fn(C.c_int(curve_id), C.c_int(field), C.c_long(iters)))


def fixed_base_mul(curve_id, group, gen_xy, scalars, threads=0):
    """out[i] = scalars[i] * G on the CPU: gen_xy affine Montgomery (2L | 4L limbs), scalars (n, 4) canonical ->
    (n, 2L | 4L) affine Montgomery points, infinity as zeros.  (generator.rs:128-196 `FixedBase::msm`.)"""
    gen_xy = np.ascontiguousarray(gen_xy, dtype=np.uint64).ravel()
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((scalars.shape[0], gen_xy.size), dtype=np.uint64)
    rc = lib().orc_fixed_base_mul(C.c_int(curve_id), C.c_int(group), _p(gen_xy), _p(scalars), C.c_size_t(scalars.shape[0]),
                                  _p(out), C.c_int(threads or (os.cpu_count() or 1)))
    assert rc == 0
    return out
