"""ctypes binding of libhekaton_b200.so (C ABI: include/hekaton_b200.h).

Every method of `Context` is one C-ABI call; numpy arrays are host buffers (HK_HOST) and
integers are raw device pointers (HK_DEVICE, e.g. `tensor.data_ptr()`).  Failing to find
the library or a B200 raises -- there is no fallback path.
"""
import ctypes as C
import os

import numpy as np

from .curves import get_curve

HK_HOST, HK_DEVICE, HK_HOST_PINNED, HK_HOST_PINNED_ASYNC = 0, 1, 2, 3
SCALAR_BIGINT, SCALAR_MONT = 0, 1
NTT_FORWARD, NTT_INVERSE = 0, 1
OPT_ACCUM_WAVES, OPT_ACCUM_SEG_MIN, OPT_FOLD_SEG, OPT_REDUCE_CHUNK, OPT_SORT_GLOBAL_ATOMICS, OPT_G2_LANE_PAIR, OPT_QUAD_MAX_THREADS, OPT_NO_FOLD_RESOLVE = 1, 2, 3, 4, 5, 6, 7, 8
DOMAIN_PLAIN, DOMAIN_COSET = 0, 1
FIELD_FR, FIELD_FQ, FIELD_FQ2 = 0, 1, 2
(FOP_MUL, FOP_ADD, FOP_SUB, FOP_INV, FOP_TO_MONT, FOP_FROM_MONT, FOP_INV_GCD, FOP_SQR, FOP_MUL_KARA,
 FOP_MULSUB2) = range(10)

ERRORS = {0: "HK_OK", -1: "HK_ERR_INVALID", -2: "HK_ERR_CUDA", -3: "HK_ERR_NOMEM", -4: "HK_ERR_LENGTH",
          -5: "HK_ERR_HANDLE", -6: "HK_ERR_UNSUPPORTED", -7: "HK_ERR_NO_DEVICE"}

LIB_PATH = os.environ.get("HK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                    "libhekaton_b200.so")   # HK_LIB: A/B builds of the same ABI


class HkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class PkDesc(C.Structure):
    _fields_ = [
        ("a_g", C.c_uint64), ("b_g", C.c_uint64), ("h_g", C.c_uint64), ("b_h", C.c_uint64),
        ("ck", C.c_uint64 * 8), ("num_stages", C.c_uint32), ("deltas_g", C.c_uint64),
        ("alpha_g", C.c_void_p), ("beta_g", C.c_void_p), ("beta_h", C.c_void_p), ("last_delta_h", C.c_void_p),
        ("r1cs", C.c_uint64), ("stage_start", C.c_uint32 * 9), ("witness_density_permille", C.c_uint32),
    ]


_lib = None


def load_library(path=None):
    """Load the CUDA library.  Raises OSError with build instructions if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OSError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      f"(or `make -C hekaton_system_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(p)
    vp, u64p, u8p, u32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    sz, u32, u64, i32 = C.c_size_t, C.c_uint32, C.c_uint64, C.c_int32
    sigs = {
        "hk_ctx_create": [C.c_int, u32, vp, C.POINTER(vp)],
        "hk_ctx_destroy": [vp],
        "hk_ctx_fork": [vp, vp, C.POINTER(vp)],
        "hk_bases_precompute": [vp, u64, u32],
        "hk_ctx_sync": [vp],
        "hk_ctx_set_option": [vp, u32, u64],
        "hk_curve_info": [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)],
        "hk_bases_upload_g1": [vp, vp, u8p, sz, u32, C.POINTER(u64)],
        "hk_bases_upload_g2": [vp, vp, u8p, sz, u32, C.POINTER(u64)],
        "hk_bases_free": [vp, u64],
        "hk_bases_len": [vp, u64, C.POINTER(sz)],
        "hk_bases_download": [vp, u64, sz, sz, vp],
        "hk_bases_from_scalars_g1": [vp, vp, u64p, sz, u32, u32, C.POINTER(u64)],
        "hk_bases_from_scalars_g2": [vp, vp, u64p, sz, u32, u32, C.POINTER(u64)],
        "hk_msm_g1": [vp, u64, sz, u64p, sz, u32, u32, vp],
        "hk_msm_g2": [vp, u64, sz, u64p, sz, u32, u32, vp],
        "hk_msm_g1_checked": [vp, u64, sz, sz, u64p, sz, u32, u32, vp],
        "hk_msm_g1_batch": [vp, u64, sz, u64p, sz, sz, u32, u32, vp],
        "hk_msm_g2_batch": [vp, u64, sz, u64p, sz, sz, u32, u32, vp],
        "hk_scalar_mul_g1": [vp, vp, u8p, u64p, sz, u32, u32, vp, u8p],
        "hk_scalar_mul_g2": [vp, vp, u8p, u64p, sz, u32, u32, vp, u8p],
        "hk_g1_into_affine": [vp, vp, sz, vp, u8p],
        "hk_g2_into_affine": [vp, vp, sz, vp, u8p],
        "hk_ntt": [vp, u64p, u32, u32, u32, u32],
        "hk_ntt_batch": [vp, u64p, u32, sz, u32, u32, u32],
        "hk_qap_h": [vp, u64p, u64p, u64p, u32, u32, u64p],
        "hk_r1cs_upload": [vp, sz, sz, sz, u32p, u32p, u64p, u32p, u32p, u64p, u32p, u32p, u64p, C.POINTER(u64)],
        "hk_r1cs_free": [vp, u64],
        "hk_qap_witness_map": [vp, u64, u64p, u32, u64p],
        "hk_pk_create": [vp, C.POINTER(PkDesc), C.POINTER(u64)],
        "hk_pk_free": [vp, u64],
        "hk_commit": [vp, u64, u32, u64p, u64p, u32, vp, u8p],
        "hk_commit_batch": [vp, u64, u32, sz, u64p, sz, u64p, u32, vp, u8p],
        "hk_prove": [vp, u64, u64p, u64p, u64p, u64p, u32, vp, vp, vp, u8p],
        "hk_prove_batch": [vp, u64, sz, u64p, u64p, u64p, u64p, u32, vp, vp, vp, u8p],
        "hk_field_op": [vp, u32, u32, u64p, u64p, u64p, sz],
        "hk_bench_field_mul": [vp, u32, u32, u32, C.POINTER(C.c_float)],
        "hk_bench_imad_wide": [vp, u32, C.POINTER(C.c_double)],
        "hk_profile_enable": [vp, i32],
        "hk_profile_reset": [vp],
        "hk_profile_read": [vp, u32, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(u64)],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    lib.hk_build_flags.argtypes = []
    lib.hk_build_flags.restype = C.c_uint32
    lib.hk_last_error_string.argtypes = [vp]
    lib.hk_last_error_string.restype = C.c_char_p
    lib.hk_launch_count.argtypes = [vp]
    lib.hk_launch_count.restype = u64
    if path is None:
        _lib = lib
    return lib


def _ptr(x):
    """numpy array -> host pointer, int -> device pointer, None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "host buffers must be C-contiguous"
        return x.ctypes.data
    return int(x)


def _space(*xs):
    kinds = {isinstance(x, np.ndarray) for x in xs if x is not None}
    assert len(kinds) <= 1, "mixing host arrays and device pointers in one call"
    return HK_HOST if (not kinds or kinds.pop()) else HK_DEVICE


class Context:
    """One GPU, one stream.  `stream` = a cudaStream_t value (e.g.
    torch.cuda.current_stream().cuda_stream) or None for a private stream."""

    def __init__(self, curve="bls12_381", device=0, stream=None, _parent=None):
        self.lib = load_library()
        self.curve = get_curve(curve)
        h = C.c_void_p()
        self._parent = _parent          # keeps the parent alive as long as a fork exists
        # stream: None -> the library creates a private stream; 0 -> the legacy default stream
        # (cudaStreamLegacy = 0x1, which is what torch's default stream is); else a cudaStream_t
        if stream is None:
            sp = None
        elif int(stream) == 0:
            sp = C.c_void_p(1)
        else:
            sp = C.c_void_p(int(stream))
        if _parent is not None:
            rc = self.lib.hk_ctx_fork(_parent.h, sp, C.byref(h))
        else:
            rc = self.lib.hk_ctx_create(int(device), self.curve.cid, sp, C.byref(h))
        if rc != 0:
            raise HkError(rc, "hk_ctx_create failed (an sm_100 GPU is required; there is no CPU fallback)")
        self.h = h
        self.device = device
        self.L = self.curve.fq_limbs

    def fork(self, stream=None):
        """A context sharing this one's uploaded resources (same handles) with its own stream and
        scratch memory -- one per concurrently running proof / host thread."""
        return Context(self.curve, self.device, stream, _parent=self)

    def close(self):
        if getattr(self, "h", None):
            self.lib.hk_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise HkError(rc, self.lib.hk_last_error_string(self.h).decode())

    def set_option(self, option, value):
        """hk_ctx_set_option: OPT_ACCUM_WAVES, OPT_ACCUM_SEG_MIN, OPT_FOLD_SEG, OPT_REDUCE_CHUNK
        (MSM tuning; results never depend on them)."""
        self._ck(self.lib.hk_ctx_set_option(self.h, option, value))

    def sync(self):
        self._ck(self.lib.hk_ctx_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.hk_launch_count(self.h))

    # ---- bases ----
    def upload_g1(self, pts, inf=None):
        """pts: (n, 2L) u64 Montgomery host array or device pointer (then pass n via tuple (ptr, n))."""
        return self._upload(self.lib.hk_bases_upload_g1, pts, inf, 2 * self.L)

    def upload_g2(self, pts, inf=None):
        return self._upload(self.lib.hk_bases_upload_g2, pts, inf, 4 * self.L)

    def _upload(self, fn, pts, inf, width):
        if isinstance(pts, tuple):
            ptr, n = pts
            space = HK_DEVICE
        else:
            pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, width)
            ptr, n, space = pts.ctypes.data, pts.shape[0], HK_HOST
        if inf is not None and isinstance(inf, np.ndarray):
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
        h = C.c_uint64()
        self._ck(fn(self.h, ptr, _ptr(inf), n, space, C.byref(h)))
        return h.value

    def bases_from_scalars(self, group, gen_affine, scalars, repr_=SCALAR_BIGINT, n=None):
        gen = np.ascontiguousarray(gen_affine, dtype=np.uint64)
        if isinstance(scalars, np.ndarray):
            scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
            n = scalars.shape[0]
        h = C.c_uint64()
        fn = self.lib.hk_bases_from_scalars_g1 if group == 1 else self.lib.hk_bases_from_scalars_g2
        self._ck(fn(self.h, gen.ctypes.data, _ptr(scalars), n, repr_, _space(scalars), C.byref(h)))
        return h.value

    def bases_len(self, handle):
        n = C.c_size_t()
        self._ck(self.lib.hk_bases_len(self.h, handle, C.byref(n)))
        return n.value

    def bases_download(self, handle, group, offset=0, n=None):
        if n is None:
            n = self.bases_len(handle) - offset
        out = np.zeros((n, (2 if group == 1 else 4) * self.L), dtype=np.uint64)
        self._ck(self.lib.hk_bases_download(self.h, handle, offset, n, out.ctypes.data))
        return out

    def bases_precompute(self, handle, density=1.0):
        self._ck(self.lib.hk_bases_precompute(self.h, handle, int(round(density * 1000))))

    def bases_free(self, handle):
        self._ck(self.lib.hk_bases_free(self.h, handle))

    # ---- msm ----
    def msm(self, group, handle, scalars, repr_=SCALAR_BIGINT, offset=0, n=None, K=1, out=None):
        """Returns the Jacobian result(s) as (K, 3L | 6L) u64 (host) or writes to device `out`."""
        if isinstance(scalars, np.ndarray):
            scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
            if n is None:
                n = scalars.size // 4 // K
        width = (3 if group == 1 else 6) * self.L
        space = _space(scalars) if scalars is not None else HK_HOST
        if space == HK_HOST:
            out = np.zeros((K, width), dtype=np.uint64)
        fn = self.lib.hk_msm_g1_batch if group == 1 else self.lib.hk_msm_g2_batch
        self._ck(fn(self.h, handle, offset, _ptr(scalars), n or 0, K, repr_, space, _ptr(out)))
        return out

    def msm_checked(self, handle, n_bases_expected, scalars, repr_=SCALAR_BIGINT, offset=0):
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros((1, 3 * self.L), dtype=np.uint64)
        self._ck(self.lib.hk_msm_g1_checked(self.h, handle, offset, n_bases_expected, scalars.ctypes.data,
                                            scalars.shape[0], repr_, HK_HOST, out.ctypes.data))
        return out

    def into_affine(self, group, jac):
        jac = np.ascontiguousarray(jac, dtype=np.uint64)
        width = (3 if group == 1 else 6) * self.L
        n = jac.size // width
        out = np.zeros((n, width // 3 * 2), dtype=np.uint64)
        inf = np.zeros(n, dtype=np.uint8)
        fn = self.lib.hk_g1_into_affine if group == 1 else self.lib.hk_g2_into_affine
        self._ck(fn(self.h, jac.ctypes.data, n, out.ctypes.data, inf.ctypes.data))
        return out, inf

    def scalar_mul(self, group, points, scalars, inf=None, repr_=SCALAR_BIGINT):
        """out[i] = scalars[i] * points[i] (pairing_ops.rs:32-39 `scalar_pairing`): host arrays
        points (n, 2L | 4L) Montgomery, scalars (n, 4) -> (affine points, infinity flags)."""
        points = np.ascontiguousarray(points, dtype=np.uint64)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = scalars.shape[0]
        assert points.shape[0] == n
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
        out = np.zeros_like(points)
        oinf = np.zeros(n, dtype=np.uint8)
        fn = self.lib.hk_scalar_mul_g1 if group == 1 else self.lib.hk_scalar_mul_g2
        self._ck(fn(self.h, points.ctypes.data, _ptr(inf), scalars.ctypes.data, n, repr_, HK_HOST, out.ctypes.data,
                    oinf.ctypes.data))
        return out, oinf

    # ---- ntt / qap ----
    def ntt(self, data, log_n, inverse=False, coset=False, K=1):
        """In place on a host array (K * 2^log_n, 4) or a device pointer."""
        if isinstance(data, np.ndarray):
            assert data.dtype == np.uint64 and data.size == (K << log_n) * 4
        self._ck(self.lib.hk_ntt_batch(self.h, _ptr(data), log_n, K, NTT_INVERSE if inverse else NTT_FORWARD,
                                       DOMAIN_COSET if coset else DOMAIN_PLAIN, _space(data)))
        return data

    def qap_h(self, a, b, c, log_n, h=None):
        space = _space(a, b, c)
        if space == HK_HOST:
            h = np.zeros((1 << log_n, 4), dtype=np.uint64)
        self._ck(self.lib.hk_qap_h(self.h, _ptr(a), _ptr(b), _ptr(c), log_n, space, _ptr(h)))
        return h

    def r1cs_upload(self, n, l, nw, mats):
        """mats = [(row_ptr u32[n+1], col u32[nnz], val u64[nnz,4] Montgomery)] * 3 for A, B, C."""
        args, keep = [], []
        for ptr, col, val in mats:
            ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
            col = np.ascontiguousarray(col, dtype=np.uint32)
            val = np.ascontiguousarray(val, dtype=np.uint64)
            keep += [ptr, col, val]
            args += [ptr.ctypes.data, col.ctypes.data if col.size else None, val.ctypes.data if val.size else None]
        h = C.c_uint64()
        self._ck(self.lib.hk_r1cs_upload(self.h, n, l, nw, *args, C.byref(h)))
        return h.value

    def r1cs_free(self, handle):
        self._ck(self.lib.hk_r1cs_free(self.h, handle))

    def witness_map(self, r1cs, z, n_out=None, h=None):
        space = _space(z)
        if space == HK_HOST:
            z = np.ascontiguousarray(z, dtype=np.uint64)
            h = np.zeros((n_out, 4), dtype=np.uint64)
        self._ck(self.lib.hk_qap_witness_map(self.h, r1cs, _ptr(z), space, _ptr(h)))
        return h

    # ---- fused path ----
    def pk_create(self, a_g, b_g, h_g, b_h, ck, deltas_g, alpha_g, beta_g, beta_h, last_delta_h, r1cs, stage_start):
        d = PkDesc()
        d.a_g, d.b_g, d.h_g, d.b_h, d.deltas_g, d.r1cs = a_g, b_g, h_g, b_h, deltas_g, r1cs
        d.num_stages = len(ck)
        for i, c in enumerate(ck):
            d.ck[i] = c
        for i, s in enumerate(stage_start):
            d.stage_start[i] = s
        keep = [np.ascontiguousarray(x, dtype=np.uint64) for x in (alpha_g, beta_g, beta_h, last_delta_h)]
        d.alpha_g, d.beta_g, d.beta_h, d.last_delta_h = [k.ctypes.data for k in keep]
        h = C.c_uint64()
        self._ck(self.lib.hk_pk_create(self.h, C.byref(d), C.byref(h)))
        return h.value

    def pk_free(self, pk):
        self._ck(self.lib.hk_pk_free(self.h, pk))

    def commit(self, pk, stage, z, kappa):
        z = np.ascontiguousarray(z, dtype=np.uint64)
        kappa = np.ascontiguousarray(kappa, dtype=np.uint64)
        out = np.zeros(2 * self.L, dtype=np.uint64)
        inf = np.zeros(1, dtype=np.uint8)
        self._ck(self.lib.hk_commit(self.h, pk, stage, z.ctypes.data, kappa.ctypes.data, HK_HOST, out.ctypes.data,
                                    inf.ctypes.data))
        return out, bool(inf[0])

    def commit_batch(self, pk, stage, z, kappa):
        """K commitments: host arrays z (K, m, 4), kappa (K, 4) -> (points (K, 2L), inf (K,))."""
        z = np.ascontiguousarray(z, dtype=np.uint64)
        K = z.shape[0]
        kappa = np.ascontiguousarray(kappa, dtype=np.uint64).reshape(K, 4)
        out = np.zeros((K, 2 * self.L), dtype=np.uint64)
        inf = np.zeros(K, dtype=np.uint8)
        self._ck(self.lib.hk_commit_batch(self.h, pk, stage, K, z.ctypes.data, z.shape[1], kappa.ctypes.data, HK_HOST,
                                          out.ctypes.data, inf.ctypes.data))
        return out, inf

    def commit_batch_device(self, pk, stage, K, z, z_stride, kappa, out, inf):
        self._ck(self.lib.hk_commit_batch(self.h, pk, stage, K, z, z_stride, kappa, HK_DEVICE, out, inf))

    def prove(self, pk, z, r, s, kappas):
        """Host buffers in, host buffers out: (A (2L), B (4L), C (2L), inf[3])."""
        z = np.ascontiguousarray(z, dtype=np.uint64)
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        kappas = np.ascontiguousarray(kappas, dtype=np.uint64)
        a = np.zeros(2 * self.L, dtype=np.uint64)
        b = np.zeros(4 * self.L, dtype=np.uint64)
        c = np.zeros(2 * self.L, dtype=np.uint64)
        inf = np.zeros(3, dtype=np.uint8)
        self._ck(self.lib.hk_prove(self.h, pk, z.ctypes.data, r.ctypes.data, s.ctypes.data,
                                   kappas.ctypes.data if kappas.size else None, HK_HOST, a.ctypes.data, b.ctypes.data,
                                   c.ctypes.data, inf.ctypes.data))
        return a, b, c, inf

    def prove_batch(self, pk, z, r, s, kappas, pinned=False):
        """K proofs against one key.  Host arrays: z (K, m, 4), r (K, 4), s (K, 4), kappas (K, ns-1, 4)
        -> (A (K, 2L), B (K, 4L), C (K, 2L), inf (K, 3)).  pinned=True promises that the input arrays
        live in page-locked memory (e.g. numpy views of torch pinned tensors): no staging copy."""
        z = np.ascontiguousarray(z, dtype=np.uint64)
        K = z.shape[0]
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(K, 4)
        s = np.ascontiguousarray(s, dtype=np.uint64).reshape(K, 4)
        kappas = np.ascontiguousarray(kappas, dtype=np.uint64)
        a = np.zeros((K, 2 * self.L), dtype=np.uint64)
        b = np.zeros((K, 4 * self.L), dtype=np.uint64)
        c = np.zeros((K, 2 * self.L), dtype=np.uint64)
        inf = np.zeros((K, 3), dtype=np.uint8)
        self._ck(self.lib.hk_prove_batch(self.h, pk, K, z.ctypes.data, r.ctypes.data, s.ctypes.data,
                                         kappas.ctypes.data if kappas.size else None,
                                         HK_HOST_PINNED if pinned else HK_HOST, a.ctypes.data,
                                         b.ctypes.data, c.ctypes.data, inf.ctypes.data))
        return a, b, c, inf

    def prove_batch_async(self, pk, z, r, s, kappas, a, b, c, inf):
        """HK_HOST_PINNED_ASYNC: every argument is a numpy view of PAGE-LOCKED memory (inputs z (K, m, 4), r, s
        (K, 4), kappas (K, ns-1, 4); outputs a (K, 2L), b (K, 4L), c (K, 2L), inf (K, 3) uint8).  Only enqueues;
        the outputs are valid after `sync()`.  Several calls may be in flight."""
        K = z.shape[0]
        self._ck(self.lib.hk_prove_batch(self.h, pk, K, z.ctypes.data, r.ctypes.data, s.ctypes.data,
                                         kappas.ctypes.data if kappas.size else None, HK_HOST_PINNED_ASYNC,
                                         a.ctypes.data, b.ctypes.data, c.ctypes.data, inf.ctypes.data))

    def commit_batch_async(self, pk, stage, z, kappa, out, inf):
        """HK_HOST_PINNED_ASYNC commitments: page-locked z (K, m, 4), kappa (K, 4) in, out (K, 2L), inf (K,) out;
        valid after `sync()`."""
        self._ck(self.lib.hk_commit_batch(self.h, pk, stage, z.shape[0], z.ctypes.data, z.shape[1], kappa.ctypes.data,
                                          HK_HOST_PINNED_ASYNC, out.ctypes.data, inf.ctypes.data))

    def prove_batch_device(self, pk, K, z, r, s, kappas, a, b, c, inf):
        """Device pointers (z K*m, r K, s K, kappas K*(ns-1); outputs K each); asynchronous."""
        self._ck(self.lib.hk_prove_batch(self.h, pk, K, z, r, s, kappas, HK_DEVICE, a, b, c, inf))

    def prove_device(self, pk, z, r, s, kappas, a, b, c, inf):
        """All arguments are device pointers; asynchronous on the context stream."""
        self._ck(self.lib.hk_prove(self.h, pk, z, r, s, kappas, HK_DEVICE, a, b, c, inf))

    # ---- hooks ----
    def field_op(self, field, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.zeros_like(a)
        width = {FIELD_FR: 4, FIELD_FQ: self.L, FIELD_FQ2: 2 * self.L}[field]
        n = a.size // width
        self._ck(self.lib.hk_field_op(self.h, field, op, a.ctypes.data, b.ctypes.data if b is not None else None,
                                      out.ctypes.data, n))
        return out

    def profile(self, on=True, serial=False):
        """serial=True additionally maps every lane of commit/prove to the context stream so that each
        launch is timed alone (hk_profile_enable mode 2)."""
        self._ck(self.lib.hk_profile_enable(self.h, (2 if serial else 1) if on else 0))

    def profile_reset(self):
        self._ck(self.lib.hk_profile_reset(self.h))

    def profile_read(self, kernel_id):
        """-> (total_ms, launches, units) for PROF_MSM_ACCUM_G1 (0), _G2 (1), PROF_NTT_PASS (2)."""
        ms, n, u = C.c_double(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.hk_profile_read(self.h, kernel_id, C.byref(ms), C.byref(n), C.byref(u)))
        return ms.value, n.value, u.value

    def bench_field_mul(self, field, threads, iters):
        ms = C.c_float()
        self._ck(self.lib.hk_bench_field_mul(self.h, field, threads, iters, C.byref(ms)))
        return ms.value

    def bench_imad_wide(self, iters=4096):
        v = C.c_double()
        self._ck(self.lib.hk_bench_imad_wide(self.h, iters, C.byref(v)))
        return v.value
