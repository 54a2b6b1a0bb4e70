"""Thin numpy-facing wrapper over the C ABI (one Context per GPU).  Mirrors what the Julia `ccall` shim does:
marshal column-major Float64 arrays, check the int32 status, raise on failure."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ReachError, dptr, fmat, fvec, iptr


KEEP_ZERO_GENERATORS = 1     # REACH_GLGM06_KEEP_ZERO_GENERATORS
RECORD_SELECTIONS = 2        # REACH_GLGM06_RECORD_SELECTIONS


class Context:
    """`reach_ctx`: one CUDA device + stream.  Raises ReachError when no B200 is visible."""

    def __init__(self, device=0, stream: Optional[int] = None):
        """`device`: a CUDA device index, or a sequence of them for a multi-device context (reach_ctx_create_multi: the BFFPSV18
        box calls then shard their rows over all listed devices inside the library).
        `stream`: a cudaStream_t handle to borrow (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self.lib = _lib.load()
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple, np.ndarray)):
            assert stream is None, "a multi-device context owns its streams"
            devs = np.ascontiguousarray(device, dtype=np.int32)
            rc = self.lib.reach_ctx_create_multi(iptr(devs), len(devs), C.byref(self._h))
            self.devices = [int(d) for d in devs]
            device = self.devices[0] if len(devs) else 0
        else:
            rc = self.lib.reach_ctx_create_on_stream(int(device), C.c_void_p(stream), C.byref(self._h))
            self.devices = [int(device)]
        if rc != 0:
            raise ReachError(rc, (self.lib.reach_last_error(None) or b"").decode())
        self.device = device
        # borrowed cudaStream_t handle; None: the context owns its stream (a NULL handle -- torch's legacy default stream --
        # is NOT borrowed: reach_ctx_create_on_stream(dev, NULL) creates a private non-blocking stream)
        self.stream = int(stream) if stream else None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.reach_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise ReachError(rc, (self.lib.reach_last_error(self._h) or b"").decode())

    def sync(self):
        self._check(self.lib.reach_sync(self._h))

    def device_count(self) -> int:
        v = C.c_int32()
        self._check(self.lib.reach_ctx_device_count(self._h, C.byref(v)))
        return int(v.value)

    def launch_count(self) -> int:
        v = C.c_uint64()
        self._check(self.lib.reach_launch_count(self._h, C.byref(v)))
        return int(v.value)

    # ---- test hook -------------------------------------------------------------------------------------
    def dgemm(self, A, B, reps: int = 1):
        A, B = fmat(A), fmat(B)
        M, K = A.shape
        K2, N = B.shape
        assert K == K2
        Cm = np.zeros((M, N), order="F")
        ms = C.c_double()
        self._check(self.lib.reach_dgemm(self._h, M, N, K, dptr(A), M, dptr(B), K, dptr(Cm), M, int(reps), C.byref(ms)))
        return Cm, ms.value

    def dgemm_left(self, P, X, reps: int = 1):
        """Y = P @ X with the generator-layout (left multiplication) variant of the DMMA GEMM."""
        P, X = fmat(P), fmat(X)
        n, K = P.shape
        K2, Q = X.shape
        assert K == K2
        Y = np.zeros((n, Q), order="F")
        ms = C.c_double()
        self._check(self.lib.reach_dgemm_left(self._h, n, Q, K, dptr(P), n, dptr(X), K, dptr(Y), n, int(reps), C.byref(ms)))
        return Y, ms.value

    # ---- discretisation --------------------------------------------------------------------------------
    def expm(self, A, delta: float) -> np.ndarray:
        A = fmat(A)
        n = A.shape[0]
        Phi = np.zeros((n, n), order="F")
        self._check(self.lib.reach_expm(self._h, n, dptr(A), n, float(delta), dptr(Phi), n))
        return Phi

    def phi12(self, A, delta: float):
        A = fmat(A)
        n = A.shape[0]
        P1 = np.zeros((n, n), order="F")
        P2 = np.zeros((n, n), order="F")
        self._check(self.lib.reach_phi12(self._h, n, dptr(A), n, float(delta), dptr(P1), n, dptr(P2), n))
        return P1, P2

    def discretize_box(self, A, delta, c, r, B=None, cU=None, rU=None, algorithm="forward", want_phi2abs=False):
        """Returns dict(Phi, c0, r0, rE, rpsi, MU[, Phi2abs])."""
        A = fmat(A)
        n = A.shape[0]
        alg = {"forward": 0, "backward": 1, "firstorder": 2, "nobloating": 3}.get(algorithm)
        if alg is None:
            raise ValueError(f"the algorithm {algorithm} is unknown")
        c, r = fvec(c), fvec(r)
        m = 0
        Bm = None
        if cU is not None:
            cU, rU = fvec(cU), fvec(rU)
            m = len(cU)
            Bm = fmat(np.eye(n) if B is None else B)
            assert Bm.shape == (n, m)
        Phi = np.zeros((n, n), order="F")
        P2 = np.zeros((n, n), order="F") if want_phi2abs else None
        c0, r0, rE = np.zeros(n), np.zeros(n), np.zeros(n)
        rpsi = np.zeros(n) if m else None
        MU = np.zeros((n, m), order="F") if m else None
        self._check(self.lib.reach_discretize_box(
            self._h, n, dptr(A), n, float(delta), alg, dptr(c), dptr(r), m, dptr(Bm), n, dptr(cU), dptr(rU),
            dptr(Phi), n, dptr(P2), n, dptr(c0), dptr(r0), dptr(rE), dptr(rpsi), dptr(MU)))
        out = dict(Phi=Phi, c0=c0, r0=r0, rE=rE, rpsi=rpsi, MU=MU, cU=cU, rU=rU)
        if want_phi2abs:
            out["Phi2abs"] = P2
        return out

    def discretize_zonotope(self, A, delta, c, r, B=None, cU=None, rU=None, algorithm="forward", remove_zero_generators=True):
        """`discretize(...; set_operations="zonotope")` on the device (reach_discretize_zonotope).
        Returns dict(Phi, c0, G0, cV, GV) -- cV / GV are None for a homogeneous system."""
        A = fmat(A)
        n = A.shape[0]
        alg = {"forward": 0, "backward": 1, "firstorder": 2, "nobloating": 3}.get(algorithm)
        if alg is None:
            raise ValueError(f"the algorithm {algorithm} is unknown")
        if alg >= 2:        # ArgumentError, discretize.jl:105-115
            raise ValueError(f"the algorithm {algorithm} with set operations=zonotope is not implemented")
        c, r = fvec(c), fvec(r)
        m, Bm = 0, None
        if cU is not None:
            cU, rU = fvec(cU), fvec(rU)
            m = len(cU)
            Bm = None if B is None else fmat(np.asarray(B, dtype=np.float64).reshape(n, -1))
            assert Bm is None or Bm.shape == (n, m)
            assert Bm is not None or m == n, "B = None means the identity and needs an n-dimensional input set"
        Phi = np.zeros((n, n), order="F")
        c0 = np.zeros(n)
        G0 = np.zeros((n, 4 * n + m + 1), order="F")
        cV = np.zeros(n) if m else None
        GV = np.zeros((n, n + m), order="F") if m else None
        p0, pV = C.c_int64(), C.c_int64()
        flags = 0 if remove_zero_generators else KEEP_ZERO_GENERATORS
        self._check(self.lib.reach_discretize_zonotope(
            self._h, n, dptr(A), n, float(delta), alg, dptr(c), dptr(r), m, dptr(Bm), n, dptr(cU), dptr(rU), flags,
            dptr(Phi), n, dptr(c0), dptr(G0), n, C.byref(p0), dptr(cV), dptr(GV), n, C.byref(pV)))
        return dict(Phi=Phi, c0=c0, G0=np.asfortranarray(G0[:, :p0.value]), cV=cV,
                    GV=np.asfortranarray(GV[:, :pV.value]) if m else None)

    def discretize_inputs_box(self, A, delta, B, cUs, rUs):
        """`VaryingInput` discretisation (discretize.jl:690-696, 729-737): K input boxes (columns of cUs / rUs, m x K) through
        B -> dict(rpsi (n, K), MU = δB (n, m)); step k's discretised input is MU * Box(cUs[:, k], rUs[:, k]) ⊕ Box(0, rpsi[:, k])."""
        A = fmat(A)
        n = A.shape[0]
        cUs, rUs = fmat(cUs), fmat(rUs)
        m, K = cUs.shape
        assert rUs.shape == (m, K)
        Bm = None if B is None else fmat(np.asarray(B, dtype=np.float64).reshape(n, -1))
        assert (Bm is None and m == n) or Bm.shape == (n, m)
        rpsi = np.zeros((n, K), order="F")
        MU = np.zeros((n, m), order="F")
        self._check(self.lib.reach_discretize_inputs_box(self._h, n, dptr(A), n, float(delta), m, dptr(Bm), n, K, dptr(cUs), m,
                                                         dptr(rUs), m, dptr(rpsi), n, dptr(MU)))
        return dict(rpsi=rpsi, MU=MU)

    # ---- BFFPSV18 --------------------------------------------------------------------------------------
    @staticmethod
    def _box_args(Phi, c0, r0, MU, cU, rU, rpsi, rows, N):
        Phi = fmat(Phi)
        n = Phi.shape[0]
        c0, r0 = fvec(c0), fvec(r0)
        m = 0
        if MU is not None:
            MU = fmat(MU).reshape(n, -1, order="F")
            m = MU.shape[1]
            cU, rU, rpsi = fvec(cU), fvec(rU), fvec(rpsi)
        else:
            cU = rU = rpsi = None
        rows = np.arange(n, dtype=np.int32) if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        keep = (Phi, c0, r0, MU, cU, rU, rpsi, rows)
        args = (n, dptr(Phi), n, dptr(c0), dptr(r0), m, dptr(MU), n, dptr(cU), dptr(rU), dptr(rpsi), len(rows),
                iptr(rows), int(N))
        return keep, args, len(rows)

    def bffpsv18_boxes(self, Phi, c0, r0, N, MU=None, cU=None, rU=None, rpsi=None, rows=None):
        """One-shot host-buffer call.  Returns (centers, radii) as (nrows, N) Fortran arrays."""
        keep, args, nrows = self._box_args(Phi, c0, r0, MU, cU, rU, rpsi, rows, N)
        centers = np.zeros((nrows, N), order="F")
        radii = np.zeros((nrows, N), order="F")
        done = C.c_int64()
        self._check(self.lib.reach_bffpsv18_boxes(self._h, *args, dptr(centers), dptr(radii), C.byref(done)))
        return centers, radii

    def bffpsv18_boxes_into(self, Phi, c0, r0, N, MU, cU, rU, rpsi, rows, centers, radii):
        """One-shot call writing into caller-owned (e.g. pinned) (nrows, N) Fortran arrays; no host-side copies when
        the inputs are already column-major float64."""
        keep, args, nrows = self._box_args(Phi, c0, r0, MU, cU, rU, rpsi, rows, N)
        assert centers.shape == (nrows, N) and radii.shape == (nrows, N)
        assert centers.flags.f_contiguous and radii.flags.f_contiguous
        done = C.c_int64()
        self._check(self.lib.reach_bffpsv18_boxes(self._h, *args, dptr(centers), dptr(radii), C.byref(done)))
        return int(done.value)

    def bffpsv18_boxes_csc(self, Phi_sparse, c0, r0, N, MU=None, cU=None, rU=None, rpsi=None, rows=None):
        """Sparse twin of `bffpsv18_boxes` (reach_blocks.jl:47-132): Φ as a scipy.sparse matrix, sent as CSC fields."""
        n, colptr, rowval, nzval = _csc_fields(Phi_sparse)
        c0, r0 = fvec(c0), fvec(r0)
        m = 0
        if MU is not None:
            MU = fmat(MU).reshape(n, -1, order="F")
            m = MU.shape[1]
            cU, rU, rpsi = fvec(cU), fvec(rU), fvec(rpsi)
        else:
            cU = rU = rpsi = None
        rows = np.arange(n, dtype=np.int32) if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        centers = np.zeros((len(rows), int(N)), order="F")
        radii = np.zeros((len(rows), int(N)), order="F")
        done = C.c_int64()
        i64p = C.POINTER(C.c_int64)
        self._check(self.lib.reach_bffpsv18_boxes_csc(self._h, n, len(nzval), colptr.ctypes.data_as(i64p),
                                                      rowval.ctypes.data_as(i64p), dptr(nzval), dptr(c0), dptr(r0), m, dptr(MU),
                                                      n, dptr(cU), dptr(rU), dptr(rpsi), len(rows), iptr(rows), int(N),
                                                      dptr(centers), dptr(radii), C.byref(done)))
        return centers, radii

    def csc_info(self):
        """Which device path served the last `bffpsv18_boxes_csc` call: dict(max_component, sparse_path) --
        sparse_path is True when Φ was block-diagonal up to a permutation (components of <= 16 states) and the zero blocks
        were skipped (reach_bffpsv18_csc_info)."""
        a, b = C.c_int64(), C.c_int32()
        self._check(self.lib.reach_bffpsv18_csc_info(self._h, C.byref(a), C.byref(b)))
        return dict(max_component=int(a.value), sparse_path=bool(b.value))

    def boxplan(self, Phi, c0, r0, N, MU=None, cU=None, rU=None, rpsi=None, rows=None) -> "BoxPlan":
        return BoxPlan(self, Phi, c0, r0, N, MU, cU, rU, rpsi, rows)

    @staticmethod
    def _blocks(block_sizes, nrows):
        if block_sizes is None:
            return 0, None
        bs = np.ascontiguousarray(block_sizes, dtype=np.int32)
        assert int(bs.sum()) == nrows
        return len(bs), bs

    def bffpsv18_check(self, Phi, c0, r0, N, ell, bound, MU=None, cU=None, rU=None, rpsi=None, rows=None, eager=True,
                       block_sizes=None):
        """First violating step (1-based) of  ell . x <= bound, or 0.  `block_sizes` groups `rows` into blocks."""
        keep, args, nrows = self._box_args(Phi, c0, r0, MU, cU, rU, rpsi, rows, N)
        ell = fvec(ell)
        assert len(ell) == nrows
        nb, bs = self._blocks(block_sizes, nrows)
        v = C.c_int64()
        self._check(self.lib.reach_bffpsv18_check(self._h, *args, nb, iptr(bs), dptr(ell), float(bound), int(bool(eager)),
                                                  C.byref(v)))
        return int(v.value)

    def bffpsv18_output_map(self, Phi, c0, r0, N, Mout, MU=None, cU=None, rU=None, rpsi=None, rows=None, block_sizes=None):
        keep, args, nrows = self._box_args(Phi, c0, r0, MU, cU, rU, rpsi, rows, N)
        Mout = fmat(Mout)
        q = Mout.shape[0]
        assert Mout.shape[1] == nrows
        nb, bs = self._blocks(block_sizes, nrows)
        centers = np.zeros((q, N), order="F")
        radii = np.zeros((q, N), order="F")
        self._check(self.lib.reach_bffpsv18_output_map(self._h, *args, nb, iptr(bs), q, dptr(Mout), q, dptr(centers),
                                                       dptr(radii)))
        return centers, radii

    # ---- GLGM06 ----------------------------------------------------------------------------------------
    def glgm06(self, Phi, c0, G0, N, max_order=10, cV=None, GV=None, flags=0) -> "Flowpipe":
        return Flowpipe(self, Phi, c0, G0, N, max_order, cV, GV, flags)

    def glgm06_sharded(self, Phi, c0, G0, N, max_order, cV, GV, rank, nranks, flags=0) -> "ShardedFlowpipe":
        return ShardedFlowpipe(self, Phi, c0, G0, N, max_order, cV, GV, flags, rank, nranks)

    def reduce_order(self, c, G, r, flags=0):
        G = fmat(G)
        n, p = G.shape
        c = fvec(c)
        pmax = max(p, int(r) * n)
        Gout = np.zeros((n, pmax), order="F")
        pout = C.c_int64()
        ind = np.zeros(p, dtype=np.int32)
        self._check(self.lib.reach_reduce_order(self._h, n, p, dptr(c), dptr(G), n, int(r), int(flags), dptr(Gout), n,
                                                C.byref(pout), iptr(ind)))
        return np.asfortranarray(Gout[:, :pout.value]), ind

    def sparse(self, A) -> "SparseSystem":
        """Upload a sparse state matrix (scipy.sparse matrix, or dense array: its nonzeros) for the lazy-expm path."""
        return SparseSystem(self, A)

    def ensemble(self, Phi, c0s, G0s, N) -> "Ensemble":
        return Ensemble(self, Phi, c0s, G0s, N)


def _csc_fields(A):
    """(n, colptr, rowval, nzval) of A in Julia's SparseMatrixCSC convention (1-based Int64 indices)."""
    if hasattr(A, "tocsc"):
        M = A.tocsc()
        M.sort_indices()
        n = M.shape[0]
        assert M.shape[0] == M.shape[1], "the state matrix must be square"
        return (n, np.ascontiguousarray(M.indptr, dtype=np.int64) + 1, np.ascontiguousarray(M.indices, dtype=np.int64) + 1,
                np.ascontiguousarray(M.data, dtype=np.float64))
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    assert A.ndim == 2 and A.shape[1] == n, "the state matrix must be square"
    rows, cols = np.nonzero(A.T)[1], np.nonzero(A.T)[0]          # column-major order of the nonzeros
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr) + 1
    return n, colptr, np.ascontiguousarray(rows, dtype=np.int64) + 1, np.ascontiguousarray(A[rows, cols], dtype=np.float64)


class SparseSystem:
    """`reach_sparse`: a sparse state matrix on the device; Φ = e^{Aδ} is only ever applied, never formed
    (`SparseMatrixExp`, discretize.jl:153-154; BFFPSV18 `:lazy_expm`, reach_blocks.jl:227-397)."""

    def __init__(self, ctx: Context, A):
        self.ctx = ctx
        self.n, colptr, rowval, nzval = _csc_fields(A)
        self.nnz = len(nzval)
        self._h = C.c_void_p()
        i64p = C.POINTER(C.c_int64)
        ctx._check(ctx.lib.reach_sparse_create(ctx._h, self.n, self.nnz, colptr.ctypes.data_as(i64p),
                                               rowval.ctypes.data_as(i64p), dptr(nzval), C.byref(self._h)))

    def expmv(self, delta: float, X, transpose: bool = False) -> np.ndarray:
        """e^{δA} X (or e^{δAᵀ} X) for a block of vectors X (n x R)."""
        X = fmat(np.asarray(X, dtype=np.float64).reshape(self.n, -1))
        Y = np.zeros_like(X, order="F")
        self.ctx._check(self.ctx.lib.reach_expmv(self._h, float(delta), int(bool(transpose)), X.shape[1], dptr(X), self.n,
                                                 dptr(Y), self.n))
        return Y

    def discretize_box(self, delta, c, r, B=None, cU=None, rU=None, algorithm="forward"):
        """Returns dict(c0, r0, rE, rpsi, MU, cU, rU) -- `Context.discretize_box` without the dense matrices."""
        n = self.n
        alg = {"forward": 0, "backward": 1, "firstorder": 2, "nobloating": 3}.get(algorithm)
        if alg is None:
            raise ValueError(f"the algorithm {algorithm} is unknown")
        c, r = fvec(c), fvec(r)
        m, Bm = 0, None
        if cU is not None:
            cU, rU = fvec(cU), fvec(rU)
            m = len(cU)
            Bm = None if B is None else fmat(np.asarray(B, dtype=np.float64).reshape(n, -1))
            assert Bm is None or Bm.shape == (n, m)
        c0, r0, rE = np.zeros(n), np.zeros(n), np.zeros(n)
        rpsi = np.zeros(n) if m else None
        MU = np.zeros((n, m), order="F") if m else None
        self.ctx._check(self.ctx.lib.reach_discretize_box_sparse(self._h, float(delta), alg, dptr(c), dptr(r), m, dptr(Bm), n,
                                                                 dptr(cU), dptr(rU), dptr(c0), dptr(r0), dptr(rE), dptr(rpsi),
                                                                 dptr(MU)))
        return dict(c0=c0, r0=r0, rE=rE, rpsi=rpsi, MU=MU, cU=cU if m else None, rU=rU if m else None)

    def bffpsv18_boxes(self, delta, c0, r0, N, MU=None, cU=None, rU=None, rpsi=None, rows=None):
        """Lazy-expm `reach_blocks!`: (centers, radii), each (nrows, N) Fortran arrays."""
        n = self.n
        c0, r0 = fvec(c0), fvec(r0)
        m = 0
        if MU is not None:
            MU = fmat(MU).reshape(n, -1, order="F")
            m = MU.shape[1]
            cU, rU, rpsi = fvec(cU), fvec(rU), fvec(rpsi)
        else:
            cU = rU = rpsi = None
        rows = np.arange(n, dtype=np.int32) if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        centers = np.zeros((len(rows), int(N)), order="F")
        radii = np.zeros((len(rows), int(N)), order="F")
        done = C.c_int64()
        self.ctx._check(self.ctx.lib.reach_bffpsv18_boxes_lazy(self._h, float(delta), dptr(c0), dptr(r0), m, dptr(MU), n,
                                                               dptr(cU), dptr(rU), dptr(rpsi), len(rows), iptr(rows), int(N),
                                                               dptr(centers), dptr(radii), C.byref(done)))
        return centers, radii

    def _lazy_args(self, c0, r0, MU, cU, rU, rpsi, rows, block_sizes):
        n = self.n
        c0, r0 = fvec(c0), fvec(r0)
        m = 0
        if MU is not None:
            MU = fmat(MU).reshape(n, -1, order="F")
            m = MU.shape[1]
            cU, rU, rpsi = fvec(cU), fvec(rU), fvec(rpsi)
        else:
            cU = rU = rpsi = None
        rows = np.arange(n, dtype=np.int32) if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        nb, bs = Context._blocks(block_sizes, len(rows))
        keep = (c0, r0, MU, cU, rU, rpsi, rows, bs)
        return keep, (dptr(c0), dptr(r0), m, dptr(MU), n, dptr(cU), dptr(rU), dptr(rpsi), len(rows), iptr(rows)), nb, bs

    def bffpsv18_output_map(self, delta, c0, r0, N, Mout, MU=None, cU=None, rU=None, rpsi=None, rows=None, block_sizes=None):
        """`output_function` branch with a lazy exponential: (centers, radii), each (q, N)."""
        keep, args, nb, bs = self._lazy_args(c0, r0, MU, cU, rU, rpsi, rows, block_sizes)
        Mout = fmat(np.asarray(Mout, dtype=np.float64).reshape(-1, args[8]))
        q = Mout.shape[0]
        centers = np.zeros((q, int(N)), order="F")
        radii = np.zeros((q, int(N)), order="F")
        self.ctx._check(self.ctx.lib.reach_bffpsv18_output_map_lazy(self._h, float(delta), *args, int(N), nb, iptr(bs), q, dptr(Mout), q,
                                                                    dptr(centers), dptr(radii)))
        return centers, radii

    def bffpsv18_check(self, delta, c0, r0, N, ell, bound, MU=None, cU=None, rU=None, rpsi=None, rows=None, eager=True,
                       block_sizes=None) -> int:
        """Check mode with a lazy exponential: first violating step (1-based) or 0."""
        keep, args, nb, bs = self._lazy_args(c0, r0, MU, cU, rU, rpsi, rows, block_sizes)
        ell = fvec(ell)
        assert len(ell) == args[8]
        v = C.c_int64()
        self.ctx._check(self.ctx.lib.reach_bffpsv18_check_lazy(self._h, float(delta), *args, int(N), nb, iptr(bs), dptr(ell),
                                                               float(bound), int(bool(eager)), C.byref(v)))
        return int(v.value)

    def timing(self):
        ms, prod, b = C.c_double(), C.c_int64(), C.c_double()
        self.ctx._check(self.ctx.lib.reach_sparse_timing(self._h, C.byref(ms), C.byref(prod), C.byref(b)))
        return dict(run_ms=ms.value, products=prod.value, bytes=b.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.ctx.lib.reach_sparse_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BoxPlan:
    """`reach_boxplan`: BFFPSV18 problem resident on the device (upload once, run, fetch)."""

    def __init__(self, ctx: Context, Phi, c0, r0, N, MU, cU, rU, rpsi, rows):
        self.ctx = ctx
        keep, args, self.nrows = ctx._box_args(Phi, c0, r0, MU, cU, rU, rpsi, rows, N)
        self.N = int(N)
        self._h = C.c_void_p()
        ctx._check(ctx.lib.reach_boxplan_create(ctx._h, *args, C.byref(self._h)))

    def run(self):
        self.ctx._check(self.ctx.lib.reach_boxplan_run(self._h))

    def fetch(self):
        centers = np.zeros((self.nrows, self.N), order="F")
        radii = np.zeros((self.nrows, self.N), order="F")
        self.ctx._check(self.ctx.lib.reach_boxplan_fetch(self._h, dptr(centers), dptr(radii)))
        return centers, radii

    def device_results(self):
        """(centers_ptr, radii_ptr): device addresses of the nrows x N column-major results (zero-copy consumers)."""
        cp, rp = _lib.c_double_p(), _lib.c_double_p()
        self.ctx._check(self.ctx.lib.reach_boxplan_device_results(self._h, C.byref(cp), C.byref(rp)))
        return C.cast(cp, C.c_void_p).value, C.cast(rp, C.c_void_p).value

    def timing(self):
        run, gemm, fl = C.c_double(), C.c_double(), C.c_double()
        nl = C.c_int64()
        self.ctx._check(self.ctx.lib.reach_boxplan_timing(self._h, C.byref(run), C.byref(gemm), C.byref(nl), C.byref(fl)))
        return dict(run_ms=run.value, gemm_ms=gemm.value, gemm_launches=nl.value, gemm_flops=fl.value)

    def close(self):
        if self._h.value:
            self.ctx.lib.reach_boxplan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Flowpipe:
    """`reach_flowpipe`: GLGM06 zonotope flowpipe resident on the device."""

    def __init__(self, ctx: Context, Phi, c0, G0, N, max_order, cV, GV, flags):
        self.ctx = ctx
        Phi = fmat(Phi)
        n = Phi.shape[0]
        c0 = fvec(c0)
        G0 = fmat(G0).reshape(n, -1, order="F")
        p0 = G0.shape[1]
        pV = 0
        if cV is not None:
            cV = fvec(cV)
            GV = fmat(GV).reshape(n, -1, order="F")
            pV = GV.shape[1]
        self.n, self.N = n, int(N)
        self._h = C.c_void_p()
        ctx._check(ctx.lib.reach_glgm06_create(ctx._h, n, dptr(Phi), n, p0, dptr(c0), dptr(G0), n, pV, dptr(cV),
                                               dptr(GV) if pV else None, n, int(N), int(max_order), int(flags),
                                               C.byref(self._h)))

    def run(self):
        self.ctx._check(self.ctx.lib.reach_glgm06_run(self._h))
        return self

    def run_fetch(self, centers, G):
        """Run and stream every reach set into caller-owned Fortran arrays centers (n, N), G (n, pmax, N)."""
        assert centers.flags.f_contiguous and G.flags.f_contiguous
        assert centers.shape == (self.n, self.N) and G.shape[0] == self.n and G.shape[2] == self.N
        self.ctx._check(self.ctx.lib.reach_glgm06_run_fetch(self._h, dptr(centers), dptr(G), G.shape[1]))
        return self

    def map_info(self):
        """dict(sparse_levels, nnz_per_row): whether the generator map runs as an ELL sparse product (structurally sparse Φ)."""
        a, b = C.c_int32(), C.c_int32()
        self.ctx._check(self.ctx.lib.reach_flowpipe_map_info(self._h, C.byref(a), C.byref(b)))
        return dict(sparse_levels=int(a.value), nnz_per_row=int(b.value))

    def chain_info(self):
        """dict(parallel_batches, batches): how many batches of the last run took the parallel selection chain."""
        a, b = C.c_int64(), C.c_int64()
        self.ctx._check(self.ctx.lib.reach_flowpipe_chain_info(self._h, C.byref(a), C.byref(b)))
        return dict(parallel_batches=int(a.value), batches=int(b.value))

    def set_serial(self, serial: bool):
        """Diagnostics: run without stream overlap so that per-kernel times are those of each kernel alone."""
        self.ctx._check(self.ctx.lib.reach_flowpipe_set_serial(self._h, int(bool(serial))))
        return self

    def ngens(self) -> np.ndarray:
        p = np.zeros(self.N, dtype=np.int32)
        self.ctx._check(self.ctx.lib.reach_flowpipe_ngens(self._h, iptr(p)))
        return p

    def step(self, k: int, p: Optional[int] = None):
        """Reach set k (1-based) -> (c, G)."""
        if p is None:
            p = int(self.ngens()[k - 1])
        c = np.zeros(self.n)
        G = np.zeros((self.n, max(p, 1)), order="F")
        self.ctx._check(self.ctx.lib.reach_flowpipe_step(self._h, int(k), dptr(c), dptr(G), self.n))
        return c, np.asfortranarray(G[:, :p])

    def fetch(self, centers=None, G=None):
        """All reach sets: centers (n, N) and G (n, pmax, N) Fortran arrays (allocated when None)."""
        p = self.ngens()
        pmax = int(p.max()) if G is None else G.shape[1]
        if centers is None:
            centers = np.zeros((self.n, self.N), order="F")
        if G is None:
            G = np.zeros((self.n, pmax, self.N), order="F")
        assert centers.flags.f_contiguous and G.flags.f_contiguous and G.shape == (self.n, pmax, self.N)
        self.ctx._check(self.ctx.lib.reach_flowpipe_fetch(self._h, dptr(centers), dptr(G), pmax))
        return centers, G, p

    def box(self, k: int):
        lo, hi = np.zeros(self.n), np.zeros(self.n)
        self.ctx._check(self.ctx.lib.reach_flowpipe_box(self._h, int(k), dptr(lo), dptr(hi)))
        return lo, hi

    def project(self, M, generators: bool = False, Gp=None):
        """`project(rs, M)` of every reach set on the device (project_reach.jl:18-99): M is q x n with q <= 4.
        Returns (centers (q, N), radii (q, N)[, Gp (q, pmax, N)]): box of M * R_k and optionally M * G_k itself."""
        M = fmat(np.asarray(M, dtype=np.float64).reshape(-1, self.n))
        q = M.shape[0]
        centers = np.zeros((q, self.N), order="F")
        radii = np.zeros((q, self.N), order="F")
        pmax = 0
        if generators and Gp is None:
            pmax = int(self.ngens().max())
            Gp = np.zeros((q, max(pmax, 1), self.N), order="F")
        if Gp is not None:
            assert Gp.flags.f_contiguous and Gp.shape[0] == q and Gp.shape[2] == self.N
            pmax = Gp.shape[1]
        self.ctx._check(self.ctx.lib.reach_flowpipe_project(self._h, q, dptr(M), q, dptr(centers), dptr(radii), dptr(Gp), pmax))
        return (centers, radii, Gp) if Gp is not None else (centers, radii)

    def selection(self, k: int, which: int = 0, pmax: int = 1 << 16):
        pin, m = C.c_int32(), C.c_int32()
        self.ctx._check(self.ctx.lib.reach_flowpipe_selection(self._h, int(k), int(which), C.byref(pin), C.byref(m),
                                                              None, None))
        ind = np.zeros(max(pin.value, 1), dtype=np.int32)
        h = np.zeros(max(pin.value, 1))
        if pin.value:
            self.ctx._check(self.ctx.lib.reach_flowpipe_selection(self._h, int(k), int(which), C.byref(pin),
                                                                  C.byref(m), iptr(ind), dptr(h)))
        return dict(p_in=pin.value, m=m.value, ind=ind[:pin.value], h=h[:pin.value])

    def timing(self):
        vals = [C.c_double() for _ in range(5)]
        nl = C.c_int64()
        self.ctx._check(self.ctx.lib.reach_flowpipe_timing(self._h, C.byref(vals[0]), C.byref(vals[1]), C.byref(nl),
                                                           C.byref(vals[2]), C.byref(vals[3]), C.byref(vals[4])))
        ch, nu = C.c_double(), C.c_double()
        self.ctx._check(self.ctx.lib.reach_flowpipe_phase_ms(self._h, C.byref(ch), C.byref(nu)))
        return dict(run_ms=vals[0].value, gemm_ms=vals[1].value, gemm_launches=nl.value, gemm_flops=vals[2].value,
                    reduce_ms=vals[3].value, reduce_bytes=vals[4].value, chain_ms=ch.value, numerics_ms=nu.value)

    def close(self):
        if self._h.value:
            self.ctx.lib.reach_flowpipe_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedFlowpipe(Flowpipe):
    """Column-sharded `reach_flowpipe` of one rank (reach_glgm06_create_sharded): driven phase by phase, see
    `sharding.run_sharded`."""

    def __init__(self, ctx: Context, Phi, c0, G0, N, max_order, cV, GV, flags, rank, nranks):
        self.ctx = ctx
        Phi = fmat(Phi)
        n = Phi.shape[0]
        c0 = fvec(c0)
        G0 = fmat(G0).reshape(n, -1, order="F")
        cV = fvec(cV)
        GV = fmat(GV).reshape(n, -1, order="F")
        self.n, self.N, self.rank, self.nranks = n, int(N), int(rank), int(nranks)
        self._h = C.c_void_p()
        ctx._check(ctx.lib.reach_glgm06_create_sharded(ctx._h, n, dptr(Phi), n, G0.shape[1], dptr(c0), dptr(G0), n, GV.shape[1],
                                                       dptr(cV), dptr(GV) if GV.shape[1] else None, n, int(N), int(max_order),
                                                       int(flags), int(rank), int(nranks), C.byref(self._h)))

    def nbatches(self) -> int:
        v = C.c_int64()
        self.ctx._check(self.ctx.lib.reach_glgm06_shard_nbatches(self._h, C.byref(v)))
        return int(v.value)

    def phase(self, batch: int, phase: int, stream: Optional[int] = None):
        """One phase of one batch; `stream` (a raw cudaStream_t) issues it there instead of on the context's stream
        (reach_glgm06_shard_phase_on: phase 0 of batch b + 1 may overlap phases 1..3 of batch b)."""
        if stream is None:
            self.ctx._check(self.ctx.lib.reach_glgm06_shard_phase(self._h, int(batch), int(phase)))
        else:
            self.ctx._check(self.ctx.lib.reach_glgm06_shard_phase_on(self._h, int(batch), int(phase), C.c_void_p(int(stream))))

    def buffer(self, batch: int, which: int):
        """(device pointer, element count, numpy dtype string) of the buffer to sum-all-reduce, or None when empty."""
        ptr, cnt, dt = C.c_void_p(), C.c_int64(), C.c_int32()
        self.ctx._check(self.ctx.lib.reach_glgm06_shard_buffer(self._h, int(batch), int(which), C.byref(ptr), C.byref(cnt),
                                                               C.byref(dt)))
        if cnt.value == 0:
            return None
        return ptr.value, int(cnt.value), ("<f8" if dt.value == 0 else "<i4")

    def run(self):
        raise ReachError(1, "a column-sharded flowpipe is driven by sharding.run_sharded (collectives between the phases)")


class Ensemble:
    """`reach_ensemble`: batched homogeneous GLGM06 flowpipes (config C5)."""

    def __init__(self, ctx: Context, Phi, c0s, G0s, N):
        self.ctx = ctx
        Phi = fmat(Phi)
        n = Phi.shape[0]
        c0s = fmat(c0s)                       # n x batch
        batch = c0s.shape[1]
        G0s = np.asfortranarray(np.asarray(G0s, dtype=np.float64))   # n x p x batch
        assert G0s.shape[0] == n and G0s.shape[2] == batch
        p = G0s.shape[1]
        self.n, self.p, self.batch, self.N = n, p, batch, int(N)
        self._h = C.c_void_p()
        ctx._check(ctx.lib.reach_ensemble_create(ctx._h, n, dptr(Phi), n, batch, p, dptr(c0s), dptr(G0s), int(N),
                                                 C.byref(self._h)))

    def run(self):
        self.ctx._check(self.ctx.lib.reach_ensemble_run(self._h))
        return self

    def step(self, b: int, k: int):
        c = np.zeros(self.n)
        G = np.zeros((self.n, self.p), order="F")
        self.ctx._check(self.ctx.lib.reach_ensemble_step(self._h, int(b), int(k), dptr(c), dptr(G)))
        return c, G

    def boxes(self, k: int):
        lo = np.zeros((self.n, self.batch), order="F")
        hi = np.zeros((self.n, self.batch), order="F")
        self.ctx._check(self.ctx.lib.reach_ensemble_boxes(self._h, int(k), dptr(lo), dptr(hi)))
        return lo, hi

    def boxes_all(self, lo=None, hi=None):
        """Box hulls of every member at every step: (lo, hi), each (n, batch, N) Fortran arrays (caller-owned when given)."""
        shape = (self.n, self.batch, self.N)
        lo = np.zeros(shape, order="F") if lo is None else lo
        hi = np.zeros(shape, order="F") if hi is None else hi
        assert lo.shape == shape and hi.shape == shape and lo.flags.f_contiguous and hi.flags.f_contiguous
        self.ctx._check(self.ctx.lib.reach_ensemble_boxes_all(self._h, dptr(lo), dptr(hi)))
        return lo, hi

    def timing(self):
        a, b = C.c_double(), C.c_double()
        self.ctx._check(self.ctx.lib.reach_ensemble_timing(self._h, C.byref(a), C.byref(b)))
        return dict(run_ms=a.value, bytes_per_step=b.value)

    def close(self):
        if self._h.value:
            self.ctx.lib.reach_ensemble_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
