"""Host-side mirror of the SciMLOperators.jl operator interface for the hot path, over the C ABI.

Julia is not installed in this image, so the package extension that a SciMLOperators.jl maintainer would
add (julia/ext/SciMLOperatorsB200Ext.jl) cannot run here; this module is its executable
twin: the same operator types, constructors, algebra, `cache_operator` / `update_coefficients!` /
`mul!` protocol and error behaviour, with every `mul!` ending in a `ccall`-equivalent ctypes call into
libscimlop_b200.so.  Names follow the reference (`!` becomes a trailing underscore).

  reference                                             here
  ----------------------------------------------------  ----------------------------------------------
  MatrixOperator            src/matrix.jl:116-133       MatrixOperator
  DiagonalOperator          src/matrix.jl:438-459       DiagonalOperator (vector) -> batch.jl:30 for 2-D
  BatchedDiagonalOperator   src/batch.jl:11-40          BatchedDiagonalOperator
  IdentityOperator / NullOperator  src/basic.jl:28,185  IdentityOperator / NullOperator
  ScalarOperator            src/scalar.jl:124-127       ScalarOperator
  ScaledOperator            src/basic.jl:357-372        ScaledOperator
  AddedOperator             src/basic.jl:596-635        AddedOperator
  ComposedOperator          src/basic.jl:909-939        ComposedOperator
  TensorProductOperator     src/tensor.jl:88-142        TensorProductOperator, kron
  TensorSumOperator         src/tensor.jl:269-298       TensorSumOperator, kronsum
  InvertibleOperator        src/matrix.jl:501-511       InvertibleOperator, factorize
  InvertedOperator          src/basic.jl:1316-1336      InvertedOperator, inv
  ldiv!(w, L, v) / ldiv!(L, v)  src/tensor.jl:612-673   ldiv_(w, L, v) / ldiv_(L, v)
  BlockDiagonalOperator     src/block.jl:37-50          BlockDiagonalOperator
  WOperator                 src/woperator.jl:176-230    WOperator
  cache_operator            src/interface.jl:245-252    cache_operator
  update_coefficients!      src/interface.jl:150-154    update_coefficients_
  mul!(w, L, v[, α, β])                                 mul_(w, L, v[, alpha, beta])
  L(w, v, u, p, t[, α, β])  src/interface.jl:161-171    L(w, v, u, p, t[, alpha, beta])

No arithmetic happens on the host: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .array import DeviceArray

_DT = {np.dtype("float64"): _lib.F64, np.dtype("float32"): _lib.F32,
       np.dtype("complex64"): _lib.C64, np.dtype("complex128"): _lib.C128}
_CT = {np.dtype("float64"): C.c_double, np.dtype("float32"): C.c_float,
       np.dtype("complex64"): C.c_float, np.dtype("complex128"): C.c_double}


def _is_complex(dtype):
    return dtype is not None and np.dtype(dtype).kind == "c"


def _num(x):
    """A host scalar as Python float, or complex when it has an imaginary part (complex operators / λ = -1im)."""
    if isinstance(x, complex) or (hasattr(x, "imag") and not isinstance(x, (int, float)) and np.iscomplexobj(x)):
        x = complex(x)
        return x if x.imag != 0.0 else x.real
    return float(x)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _scalar(dtype, x):
    """Host scalar in the operand dtype, by reference (the ABI takes α/β as host pointers). None -> NULL."""
    if x is None:
        return None
    if _is_complex(dtype):            # (re, im) pair in the real base type
        z = complex(x)
        return (_CT[np.dtype(dtype)] * 2)(z.real, z.imag)
    return C.byref(_CT[np.dtype(dtype)](float(x)))


def _as_device(a):
    return a if isinstance(a, DeviceArray) else DeviceArray.from_numpy(np.asarray(a))


def _ncols(v):
    return v.shape[1] if v.ndim == 2 else 1


def _check_vw(L, w, v):
    m, n = L.size()
    if v.shape[0] != n or w.shape[0] != m or _ncols(v) != _ncols(w):
        raise AssertionError(f"DimensionMismatch: operator {m}x{n}, v {v.shape}, w {w.shape}")
    if v.dtype != w.dtype:
        raise TypeError("v and w must have the same element type")
    dt = L.dtype
    if dt is not None and dt != v.dtype:
        raise TypeError(f"operator eltype {dt} != array eltype {v.dtype}: the native ABI takes one dtype "
                        "(SURVEY.md §8 pitfalls); the Julia wrapper falls back to the generic path here")
    if w.t.data_ptr() == v.t.data_ptr() and w.t.numel() > 0:
        raise AssertionError("mul!: w and v must not alias")


def _check_ldiv(L, w, v):
    m, n = L.size()
    if v.shape[0] != m or w.shape[0] != n or _ncols(v) != _ncols(w):
        raise AssertionError(f"DimensionMismatch: ldiv! with operator {m}x{n}, v {v.shape}, w {w.shape}")
    if v.dtype != w.dtype or (L.dtype is not None and L.dtype != v.dtype):
        raise TypeError("ldiv!: operator, v and w must have one element type")


def _diag_div(diag, ldd, w, v, scale):
    h = _lib.handle(w.device_index)
    n, k = v.shape[0], _ncols(v)
    h.check(h.lib.scimlop_diag_div(h.ptr, _DT[w.dtype], n, k, C.c_void_p(diag.ptr), ldd, C.c_void_p(v.ptr), v.ld,
                                   C.c_void_p(w.ptr), w.ld, _stream()), "scimlop_diag_div")
    if scale is not None and float(scale) != 1.0:
        _lmul(w, float(scale), h)
    return w


class BandedMatrix:
    """Device storage of a banded n x n matrix (Tridiagonal, pentadiagonal, ...: what a finite-difference
    `MatrixOperator` holds).  `bands` is n x (2b+1): diagonal d in column d+b, aligned by row
    (bands[i, d+b] = A[i, i+d]) -- the SCIMLOP_BANDED layout of the C ABI."""

    def __init__(self, bands, b):
        self.bands = _as_device(bands)
        self.b = int(b)
        if self.bands.ndim != 2 or self.bands.shape[1] != 2 * self.b + 1:
            raise ValueError("BandedMatrix: bands must be n x (2b+1)")

    @staticmethod
    def from_dense(A, b):
        A = np.asarray(A)
        n = A.shape[0]
        bands = np.zeros((n, 2 * b + 1), dtype=A.dtype, order="F")
        for d in range(-b, b + 1):
            i = np.arange(max(0, -d), min(n, n - d))
            bands[i, d + b] = A[i, i + d]
        return BandedMatrix(bands, b)

    @property
    def shape(self):
        return (self.bands.shape[0], self.bands.shape[0])

    @property
    def dtype(self):
        return self.bands.dtype


class SparseMatrixCSR:
    """Device storage of a general sparse m x n matrix in CSR (0-based 32-bit indices): what a `MatrixOperator(sprand(...))`
    holds (test/basic.jl:405-426, test/Downstream/alloccheck.jl:37-45).  Built from a SciPy sparse matrix, a dense
    array, or the three CSR arrays."""

    def __init__(self, rowptr, colidx, values, shape):
        self.rowptr = _as_device(np.ascontiguousarray(rowptr, dtype=np.int32)) if not isinstance(rowptr, DeviceArray) else rowptr
        self.colidx = _as_device(np.ascontiguousarray(colidx, dtype=np.int32)) if not isinstance(colidx, DeviceArray) else colidx
        self.values = _as_device(values)
        self.shape = (int(shape[0]), int(shape[1]))
        if self.rowptr.shape[0] != self.shape[0] + 1 or self.colidx.shape[0] != self.values.shape[0]:
            raise ValueError("SparseMatrixCSR: rowptr needs rows + 1 entries, colidx and values nnz entries each")

    @staticmethod
    def from_scipy(A):
        A = A.tocsr()
        A.sum_duplicates()
        return SparseMatrixCSR(A.indptr, A.indices, np.asarray(A.data), A.shape)

    @staticmethod
    def from_dense(A):
        A = np.asarray(A)
        rows, cols = np.nonzero(A)                    # row-major order: already sorted by row
        rowptr = np.zeros(A.shape[0] + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows, minlength=A.shape[0]), out=rowptr[1:])
        return SparseMatrixCSR(rowptr, cols.astype(np.int32), np.ascontiguousarray(A[rows, cols]), A.shape)

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def nnz(self):
        return self.values.shape[0]

    @property
    def device_index(self):
        return self.values.device_index


def _csr_transpose(A):
    """CSR form of A^T (= the CSC form of A) on A's device: entries sorted by (column, row).  One-time set-up work of
    cache_operator for a lazy adjoint, not part of mul!."""
    m, n = A.shape
    rowptr, colidx, vals = A.rowptr.t, A.colidx.t, A.values.t
    dev = vals.device
    counts = (rowptr[1:] - rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(m, device=dev), counts)
    order = torch.argsort(colidx.long() * max(m, 1) + rows, stable=True)
    new_rowptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    if colidx.numel():
        new_rowptr[1:] = torch.cumsum(torch.bincount(colidx.long(), minlength=n), 0).to(torch.int32)
    nnz = int(colidx.numel())
    return SparseMatrixCSR(DeviceArray(new_rowptr, (n + 1,)), DeviceArray(rows[order].to(torch.int32).contiguous(), (nnz,)),
                           DeviceArray(vals[order].contiguous(), (nnz,)), (n, m))


class ScalarOperator:
    """src/scalar.jl:124-127 -- mutable scalar; `update_func(old, u, p, t)` gives the new value (:267-270)."""

    def __init__(self, val, update_func=None):
        self.val = val
        self.update_func = update_func

    def convert(self):
        return self.val  # convert(Number, λ), src/scalar.jl:219

    def update_coefficients_(self, u, p, t):
        if self.update_func is not None:
            self.val = self.update_func(self.val, u, p, t)


class AbstractSciMLOperator:
    precision = _lib.PREC_DEFAULT
    dtype = None  # np.dtype of the stored arrays; None for Identity/Null (Bool eltype in the reference)

    # ---- interface ---------------------------------------------------------------------------------
    def size(self):
        raise NotImplementedError

    def getops(self):
        return ()

    def iscached(self):  # src/interface.jl:199-204
        return all(op.iscached() for op in self.getops() if isinstance(op, AbstractSciMLOperator))

    def cache_operator(self, v):  # src/interface.jl:245-252
        return self

    def update_coefficients_(self, u=None, p=None, t=None):  # src/interface.jl:150-154
        for op in self.getops():
            op.update_coefficients_(u, p, t)
        return self

    def mul_(self, w, v, alpha=None, beta=None):
        raise NotImplementedError

    def _flat_terms(self):
        """[(scale, [leaf, ...]), ...] if this operator is a sum of products of native leaves, else None."""
        return None

    # ---- the solve side: has_ldiv / ldiv! (src/interface.jl has_ldiv, has_ldiv!) ---------------------------
    def has_ldiv(self):
        return False

    def _ldiv(self, w, v, scale=None):
        """w = scale * (L \\ v); w may be v itself for the two-argument form."""
        raise TypeError(f"{type(self).__name__} has no ldiv!: factorize(L) first (has_ldiv(L) is false)")

    def ldiv_(self, *args):
        """`ldiv!(w, L, v)` as L.ldiv_(w, v); `ldiv!(L, v)` as L.ldiv_(v).  Returns the destination."""
        if len(args) == 1:
            (v,) = args
            _check_ldiv(self, v, v)
            return self._ldiv(v, v)
        w, v = args
        _check_ldiv(self, w, v)
        if w.t.data_ptr() == v.t.data_ptr() and w.t.numel() > 0:
            raise AssertionError("ldiv!(w, L, v): w and v must not alias (use the two-argument form)")
        return self._ldiv(w, v)

    # ---- call forms: src/interface.jl:161-171 --------------------------------------------------------
    def __call__(self, *args):
        if len(args) == 4:      # L(v, u, p, t): out of place
            v, u, p, t = args
            self.update_coefficients_(u, p, t)
            return self * v
        if len(args) == 5:      # L(w, v, u, p, t)
            w, v, u, p, t = args
            self.update_coefficients_(u, p, t)
            return self.mul_(w, v)
        if len(args) == 7:      # L(w, v, u, p, t, α, β)
            w, v, u, p, t, a, b = args
            self.update_coefficients_(u, p, t)
            return self.mul_(w, v, a, b)
        raise TypeError("call forms: L(v,u,p,t), L(w,v,u,p,t), L(w,v,u,p,t,α,β)")

    # ---- algebra: src/basic.jl:375-390, 711-720, 936-939; src/tensor.jl:149-164 -----------------------
    def __add__(self, other):
        if isinstance(other, AbstractSciMLOperator):
            return AddedOperator(self, other)
        if np.isscalar(other):  # L + λ  ==  L + λ*I  (src/basic.jl:711-720)
            return AddedOperator(self, ScaledOperator(other, IdentityOperator(self.size()[0])))
        return NotImplemented

    __radd__ = __add__

    def __neg__(self):
        return ScaledOperator(-1.0, self)

    def __sub__(self, other):
        if isinstance(other, AbstractSciMLOperator):
            return AddedOperator(self, ScaledOperator(-1.0, other))
        if np.isscalar(other):
            return AddedOperator(self, ScaledOperator(-other, IdentityOperator(self.size()[0])))
        return NotImplemented

    def __mul__(self, other):
        if isinstance(other, AbstractSciMLOperator):
            return ComposedOperator(self, other)
        if isinstance(other, ScalarOperator) or np.isscalar(other):
            return ScaledOperator(other, self)
        if isinstance(other, DeviceArray):  # out-of-place L * v
            if not self.iscached():
                self.cache_operator(other)
            m = self.size()[0]
            w = DeviceArray.empty((m,) + other.shape[1:], other.dtype, other.device_index)
            return self.mul_(w, other)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, ScalarOperator) or np.isscalar(other):
            return ScaledOperator(other, self)
        return NotImplemented

    __matmul__ = __mul__

    def __truediv__(self, other):
        if np.isscalar(other):
            return ScaledOperator(1.0 / other, self)
        return NotImplemented

    def adjoint(self):
        raise NotImplementedError(f"adjoint of {type(self).__name__} is outside the B200 hot path")

    @property
    def T(self):
        return self.adjoint()  # real element types only: adjoint == transpose


def _beta_is_zero(beta):
    return beta is None or beta is False or beta == 0


def _lmul(w, beta, h):
    """lmul!(β, w) with a strong zero (src/basic.jl:251,263) -> scimlop_axpby with V == NULL."""
    n, k = w.shape[0], _ncols(w)
    if _is_complex(w.dtype):
        w.t.mul_(0 if _beta_is_zero(beta) else complex(beta))    # complex arrays: torch is the allocator AND this scale
        return w
    b = 0.0 if _beta_is_zero(beta) else float(beta)
    h.check(h.lib.scimlop_axpby(h.ptr, _DT[w.dtype], n, k, None, 0, C.c_void_p(w.ptr), w.ld, None,
                                _scalar(w.dtype, b), _stream()), "scimlop_axpby")
    return w


# ====================================================================================================
# leaves
# ====================================================================================================
_SPLIT_IMAGES = {}   # (device, ptr, ld, rows, cols) -> [image buffer, refcount]: attached Float32 hi / lo images


class MatrixOperator(AbstractSciMLOperator):
    """src/matrix.jl:116-133; `mul!` :337-350.  `trans=True` is the lazy Adjoint wrapper produced by
    `adjoint(::MatrixOperator)` on a constant operator (:174-175).  `update_func(A, u, p, t)` mutates the
    device matrix in place (:257-271)."""

    def __init__(self, A, trans=False, update_func=None):
        if hasattr(A, "tocsr"):                      # a SciPy sparse matrix
            A = SparseMatrixCSR.from_scipy(A)
        self.sparse = isinstance(A, SparseMatrixCSR)
        self.banded = isinstance(A, BandedMatrix)
        self.A = A if (self.banded or self.sparse) else _as_device(A)
        if not (self.banded or self.sparse) and self.A.ndim != 2:
            raise ValueError("MatrixOperator needs a matrix")
        if self.banded and trans:
            raise NotImplementedError("adjoint of a banded MatrixOperator: transpose the bands on the host")
        self.trans = bool(trans)
        self.update_func = update_func

    @property
    def dtype(self):
        return self.A.dtype

    def size(self):
        m, n = self.A.shape
        return (n, m) if self.trans else (m, n)


    def adjoint(self):
        if self.update_func is not None:
            raise NotImplementedError("adjoint of a time-dependent MatrixOperator builds an AdjointOperator (src/left.jl) -- out of scope")
        if self.sparse and _is_complex(self.A.dtype):
            # the CSR kernels' `trans` is the plain transpose: the adjoint of a complex sparse matrix conjugates a copy
            A = self.A
            return MatrixOperator(SparseMatrixCSR(A.rowptr, A.colidx, DeviceArray(A.values.t.conj().resolve_conj().clone(), A.values.shape),
                                                  A.shape), not self.trans)
        return MatrixOperator(self.A, not self.trans)

    # ---- Float32 matrices beyond the resident-factor kernel: hi / lo image prepared once (scimlop_split_attach) ----
    _split = None            # (key into _SPLIT_IMAGES, handle) once attached
    SPLIT_LIMIT_BYTES = 8 << 30

    def _attach_split(self):
        """`cache_operator` side of the no-allocation contract (test/Downstream/alloccheck.jl): the 3xTF32 halves of
        a large Float32 matrix are made once here, so that `mul!` is one launch with no scratch allocation."""
        if self.banded or self.sparse or self._split is not None or self.A.dtype != np.float32:
            return self
        m, n = self.A.shape
        if min(m, n) < 16 or max(m, n) <= 128:
            return self          # SIMT / skinny / resident-factor kernels: nothing to prepare
        lib = _lib.load()
        nbytes = C.c_size_t(0)
        lib.scimlop_split_bytes(m, n, C.byref(nbytes))
        if nbytes.value > self.SPLIT_LIMIT_BYTES:
            return self          # huge operator: its (smaller) input is split per call instead
        h = _lib.handle(self.A.device_index)
        key = (self.A.device_index, self.A.ptr, self.A.ld, m, n)
        ent = _SPLIT_IMAGES.get(key)             # operators that share one array (L and L') share one image
        if ent is None:
            buf = torch.empty(nbytes.value, dtype=torch.uint8, device=self.A.t.device)
            h.check(lib.scimlop_split_attach(h.ptr, C.c_void_p(self.A.ptr), self.A.ld, m, n, C.c_void_p(buf.data_ptr()),
                                             nbytes.value, _stream()), "scimlop_split_attach")
            ent = _SPLIT_IMAGES[key] = [buf, 0]
        ent[1] += 1
        self._split = (key, h)
        return self

    def cache_operator(self, v):
        if self.sparse:
            nbytes = C.c_size_t(0)
            A, trans = self.A, int(self.trans)
            if self.trans:
                # The lazy adjoint is constant (adjoint() refuses time-dependent matrices): build the CSR form of A^T
                # once -- what CUSPARSE's CSR -> CSC conversion gives a Julia caller -- so that mul! is the gather kernel
                # on it (deterministic, staged, coalesced) instead of the atomic scatter.
                self._csr_T = A = _csr_transpose(self.A)
                trans = 0
            _lib.load().scimlop_csr_workspace_bytes(_DT[v.dtype], trans, A.shape[0], A.shape[1], _ncols(v),
                                                    C.byref(nbytes))
            self._csr_ws = torch.empty(nbytes.value, dtype=torch.uint8, device=v.t.device) if nbytes.value else None
            return self
        return self._attach_split()

    def __del__(self):
        sp = getattr(self, "_split", None)
        if sp is not None:
            try:
                key, h = sp
                ent = _SPLIT_IMAGES.get(key)
                if ent is not None:
                    ent[1] -= 1
                    if ent[1] <= 0:
                        h.lib.scimlop_split_detach(h.ptr, C.c_void_p(key[1]))
                        del _SPLIT_IMAGES[key]
            except Exception:
                pass

    def update_coefficients_(self, u=None, p=None, t=None):
        if self.update_func is not None:
            self.update_func(self.A, u, p, t)
            if self._split is not None:      # the matrix changed in place: redo its hi / lo image (one launch)
                h = self._split[1]
                h.check(h.lib.scimlop_split_refresh(h.ptr, C.c_void_p(self.A.ptr), _stream()), "scimlop_split_refresh")
        return self

    def _flat_terms(self):
        return [(1.0, [self])]

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        h = _lib.handle(w.device_index)
        m, n = self.size()
        if _is_complex(w.dtype) and not self.sparse:   # ComplexF32 / ComplexF64: the one-factor case of the complex Kronecker planner
            if self.banded:
                raise NotImplementedError("complex banded MatrixOperator is outside the B200 backend")
            k = _ncols(v)
            sm, sn = self.A.shape
            dims, lds = (C.c_int64 * 2)(m, n), (C.c_int64 * 1)(self.A.ld)
            kinds = (C.c_int32 * 1)(_lib.DENSE | (_lib.TRANS if self.trans else 0))
            ptrs = (C.c_void_p * 1)(self.A.ptr)
            nbytes = C.c_size_t(0)
            h.check(h.lib.scimlop_kron_workspace_bytes(_DT[w.dtype], 1, dims, kinds, k, C.byref(nbytes)), "scimlop_kron_workspace_bytes")
            ws = getattr(self, "_cplx_ws", None)
            if ws is None or ws.numel() < nbytes.value:
                ws = self._cplx_ws = torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=w.t.device)
            h.check(h.lib.scimlop_kron_mul(h.ptr, _DT[w.dtype], self.precision, 1, ptrs, dims, lds, kinds, C.c_void_p(v.ptr), n,
                                           C.c_void_p(w.ptr), m, k, _scalar(w.dtype, alpha), _scalar(w.dtype, beta),
                                           C.c_void_p(ws.data_ptr()), ws.numel(), _stream()), "scimlop_kron_mul")
            return w
        if self.sparse:   # general sparse matrix: CSR gather (or, for the lazy adjoint, scatter) kernel
            A, trans = self.A, int(self.trans)
            if self.trans and getattr(self, "_csr_T", None) is not None:
                A, trans = self._csr_T, 0           # cached adjoint: gather on the CSR form of A^T
            ws = getattr(self, "_csr_ws", None)     # allocated by cache_operator (row-contiguous staging copy of V)
            h.check(h.lib.scimlop_csr_mul(h.ptr, _DT[w.dtype], trans, A.shape[0], A.shape[1], _ncols(v),
                                          C.c_void_p(A.rowptr.ptr), C.c_void_p(A.colidx.ptr), C.c_void_p(A.values.ptr), 0,
                                          C.c_void_p(v.ptr), v.ld, C.c_void_p(w.ptr), w.ld, _scalar(w.dtype, alpha),
                                          _scalar(w.dtype, beta), None if ws is None else C.c_void_p(ws.data_ptr()),
                                          0 if ws is None else ws.numel(), _stream()), "scimlop_csr_mul")
            return w
        if self.banded:   # a one-factor "Kronecker product": the banded mode-product kernel
            nb = self.A.shape[0]
            dims, lds = (C.c_int64 * 2)(nb, nb), (C.c_int64 * 1)(self.A.b)
            kinds, ptrs = (C.c_int32 * 1)(_lib.BANDED), (C.c_void_p * 1)(self.A.bands.ptr)
            h.check(h.lib.scimlop_kron_mul(h.ptr, _DT[w.dtype], self.precision, 1, ptrs, dims, lds, kinds,
                                           C.c_void_p(v.ptr), nb, C.c_void_p(w.ptr), nb, _ncols(v),
                                           _scalar(w.dtype, alpha), _scalar(w.dtype, beta), None, 0, _stream()),
                    "scimlop_kron_mul")
            return w
        h.check(h.lib.scimlop_matmul(h.ptr, _DT[w.dtype], self.precision, int(self.trans), m, n, _ncols(v),
                                     C.c_void_p(self.A.ptr), self.A.ld, C.c_void_p(v.ptr), v.ld,
                                     C.c_void_p(w.ptr), w.ld, _scalar(w.dtype, alpha), _scalar(w.dtype, beta),
                                     _stream()), "scimlop_matmul")
        return w


class _VectorDiagonalOperator(AbstractSciMLOperator):
    """`DiagonalOperator(::AbstractVector)` = MatrixOperator(Diagonal(d)), src/matrix.jl:438-459."""

    def __init__(self, d, update_func=None):
        self.diag = _as_device(d)
        if self.diag.ndim != 1:
            raise ValueError("vector diagonal expected")
        self.update_func = update_func

    @property
    def dtype(self):
        return self.diag.dtype

    def size(self):
        n = self.diag.shape[0]
        return (n, n)

    def adjoint(self):
        return self

    def update_coefficients_(self, u=None, p=None, t=None):
        if self.update_func is not None:
            self.update_func(self.diag, u, p, t)
        return self

    def _flat_terms(self):
        return [(1.0, [self])]

    def has_ldiv(self):
        return True

    def _ldiv(self, w, v, scale=None):  # stdlib ldiv!(::Diagonal, v) behind src/matrix.jl:363-366
        return _diag_div(self.diag, 0, w, v, scale)

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        h = _lib.handle(w.device_index)
        n = self.diag.shape[0]
        h.check(h.lib.scimlop_diag_mul(h.ptr, _DT[w.dtype], n, _ncols(v), C.c_void_p(self.diag.ptr), 0,
                                       C.c_void_p(v.ptr), v.ld, C.c_void_p(w.ptr), w.ld,
                                       _scalar(w.dtype, alpha), _scalar(w.dtype, beta), _stream()),
                "scimlop_diag_mul")
        return w


class BatchedDiagonalOperator(AbstractSciMLOperator):
    """src/batch.jl:11-28: the diagonal has the same shape as the (N x K) input; `mul!` :151-179."""

    def __init__(self, diag, update_func=None):
        self.diag = _as_device(diag)
        if self.diag.ndim != 2:
            raise ValueError("BatchedDiagonalOperator needs an N x K diagonal")
        self.update_func = update_func

    @property
    def dtype(self):
        return self.diag.dtype

    def size(self):
        n = self.diag.shape[0]
        return (n, n)

    def adjoint(self):
        return self  # src/batch.jl:49-50 (real eltype)

    def update_coefficients_(self, u=None, p=None, t=None):
        if self.update_func is not None:
            self.update_func(self.diag, u, p, t)
        return self

    def _flat_terms(self):
        return [(1.0, [self])]

    def has_ldiv(self):
        return True

    def _ldiv(self, w, v, scale=None):  # src/batch.jl:181-200
        if tuple(v.shape) != tuple(self.diag.shape):
            raise AssertionError(f"BatchedDiagonalOperator: v {v.shape} must match diag {self.diag.shape}")
        return _diag_div(self.diag, self.diag.ld, w, v, scale)

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if tuple(v.shape) != tuple(self.diag.shape):
            raise AssertionError(f"BatchedDiagonalOperator: v {v.shape} must match diag {self.diag.shape}")
        h = _lib.handle(w.device_index)
        n, k = self.diag.shape
        h.check(h.lib.scimlop_diag_mul(h.ptr, _DT[w.dtype], n, k, C.c_void_p(self.diag.ptr), self.diag.ld,
                                       C.c_void_p(v.ptr), v.ld, C.c_void_p(w.ptr), w.ld,
                                       _scalar(w.dtype, alpha), _scalar(w.dtype, beta), _stream()),
                "scimlop_diag_mul")
        return w


def DiagonalOperator(d, update_func=None):
    """Vector -> diagonal MatrixOperator (src/matrix.jl:438); array -> BatchedDiagonalOperator (src/batch.jl:30)."""
    d = _as_device(d)
    if d.ndim == 1:
        return _VectorDiagonalOperator(d, update_func)
    return BatchedDiagonalOperator(d, update_func)


class IdentityOperator(AbstractSciMLOperator):
    """src/basic.jl:28-30; `mul!` :74-90."""

    def __init__(self, n):
        self.len = int(n)

    def size(self):
        return (self.len, self.len)

    def adjoint(self):
        return self

    def _flat_terms(self):
        return [(1.0, [self])]

    def has_ldiv(self):
        return True

    def _ldiv(self, w, v, scale=None):  # src/basic.jl:92-99: a copy (nothing at all in place)
        h = _lib.handle(w.device_index)
        if w.t.data_ptr() == v.t.data_ptr():
            return w if scale is None or float(scale) == 1.0 else _lmul(w, float(scale), h)
        h.check(h.lib.scimlop_axpby(h.ptr, _DT[w.dtype], self.len, _ncols(v), C.c_void_p(v.ptr), v.ld,
                                    C.c_void_p(w.ptr), w.ld, _scalar(w.dtype, scale), None, _stream()), "scimlop_axpby")
        return w

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        h = _lib.handle(w.device_index)
        h.check(h.lib.scimlop_axpby(h.ptr, _DT[w.dtype], self.len, _ncols(v), C.c_void_p(v.ptr), v.ld,
                                    C.c_void_p(w.ptr), w.ld, _scalar(w.dtype, alpha), _scalar(w.dtype, beta),
                                    _stream()), "scimlop_axpby")
        return w


class NullOperator(AbstractSciMLOperator):
    """src/basic.jl:185-188; `mul!` :248-264 (`lmul!(false, w)` / `lmul!(β, w)`)."""

    def __init__(self, n):
        self.len = int(n)

    def size(self):
        return (self.len, self.len)

    def adjoint(self):
        return self

    def _flat_terms(self):
        return [(1.0, [self])]

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        return _lmul(w, beta, _lib.handle(w.device_index))


_LEAVES = (MatrixOperator, _VectorDiagonalOperator, BatchedDiagonalOperator, IdentityOperator, NullOperator)


# ====================================================================================================
# tree plan: an operator that flattens to Σ_t λ_t Π_f leaf  ->  ONE scimlop_tree_mul call
# ====================================================================================================
class _TreePlan:
    def __init__(self, owner, terms, v):
        self.owner = owner
        self.terms = terms  # [(scale_getter, [leaf...])]
        self.k = _ncols(v)
        self.dtype = v.dtype
        nt = len(terms)
        self.c_terms = (_lib.Term * nt)()
        self._keep = []
        self._promoted = []
        for i, (_, leaves) in enumerate(terms):
            arr = (_lib.Factor * len(leaves))()
            for j, leaf in enumerate(leaves):
                leaf = _promote_leaf(leaf, self.dtype) if isinstance(leaf, (MatrixOperator, _VectorDiagonalOperator)) else leaf
                self._promoted.append(leaf)
                m, n = leaf.size()
                f = arr[j]
                f.rows, f.cols, f.trans, f.data, f.ld = m, n, 0, None, 0
                if isinstance(leaf, MatrixOperator):
                    leaf._attach_split()
                    f.kind, f.trans, f.data, f.ld = _lib.DENSE, int(leaf.trans), leaf.A.ptr, leaf.A.ld
                elif isinstance(leaf, _VectorDiagonalOperator):
                    f.kind, f.data = _lib.DIAG, leaf.diag.ptr
                elif isinstance(leaf, BatchedDiagonalOperator):
                    if leaf.diag.shape[1] != self.k:
                        raise AssertionError(f"BatchedDiagonalOperator diag {leaf.diag.shape} does not match k={self.k}")
                    f.kind, f.data, f.ld = _lib.BDIAG, leaf.diag.ptr, leaf.diag.ld
                elif isinstance(leaf, IdentityOperator):
                    f.kind = _lib.IDENTITY
                else:
                    f.kind = _lib.NULL
                if leaf.dtype is not None and leaf.dtype != self.dtype:
                    raise TypeError(f"operator eltype {leaf.dtype} != array eltype {self.dtype}")
            self._keep.append(arr)
            self.c_terms[i].nfactors = len(leaves)
            self.c_terms[i].factors = arr
        self._refresh_scales()
        nbytes = C.c_size_t(0)
        lib = _lib.load()
        rc = lib.scimlop_tree_workspace_bytes(_DT[self.dtype], self.c_terms, nt, self.k, C.byref(nbytes))
        if rc != _lib.OK:
            raise _lib.ScimlopError(rc, "scimlop_tree_workspace_bytes")
        self.ws_bytes = nbytes.value
        self.ws = torch.empty(max(1, self.ws_bytes), dtype=torch.uint8, device=v.t.device)

    def _refresh_scales(self):
        for i, (scale, _) in enumerate(self.terms):
            self.c_terms[i].scale = float(scale())

    def mul_(self, w, v, alpha, beta, precision):
        if _ncols(v) != self.k or v.dtype != self.dtype:
            raise AssertionError("operator was cached for a different input shape / eltype: call cache_operator again")
        self._refresh_scales()  # ScalarOperators may have been updated (src/scalar.jl:267-270)
        h = _lib.handle(w.device_index)
        h.check(h.lib.scimlop_tree_mul(h.ptr, _DT[w.dtype], precision, self.c_terms, len(self.terms),
                                       C.c_void_p(v.ptr), v.ld, C.c_void_p(w.ptr), w.ld, self.k,
                                       _scalar(w.dtype, alpha), _scalar(w.dtype, beta),
                                       C.c_void_p(self.ws.data_ptr()), self.ws_bytes, _stream()),
                "scimlop_tree_mul")
        return w


def _scale_getter(scalars):
    def get():
        x = 1.0
        for s in scalars:
            x *= s.convert() if isinstance(s, ScalarOperator) else s
        return x
    return get


def _terms_with_scalars(L, scalars=()):
    """Flatten like src/basic.jl:619-635 (sums) and :936-939 (products); returns [(scalars, leaves)] or None."""
    if isinstance(L, InvertibleOperator):
        return _terms_with_scalars(L.L, scalars)      # mul! of an InvertibleOperator is mul! of L.L
    if isinstance(L, MatrixOperator) and (L.banded or L.sparse):
        return None    # banded / sparse leaves are applied by their own kernels, not by scimlop_tree_mul
    if isinstance(L, _LEAVES):
        return [(tuple(scalars), [L])]
    if isinstance(L, ScaledOperator):
        return _terms_with_scalars(L.L, tuple(scalars) + (L.lam,))
    if isinstance(L, AddedOperator):
        out = []
        for op in L.ops:
            t = _terms_with_scalars(op, scalars)
            if t is None:
                return None
            out.extend(t)
        return out
    if isinstance(L, ComposedOperator):
        sc, leaves = tuple(scalars), []
        for op in L.ops:
            t = _terms_with_scalars(op, ())
            if t is None or len(t) != 1:  # distributing a sum over a product would duplicate work
                return None
            sc += t[0][0]
            leaves.extend(t[0][1])
        return [(sc, leaves)]
    return None


class _TreeMixin:
    """Shared by Scaled/Added/Composed: try the one-call native plan, else recurse like the reference."""
    _plan = None

    def _try_plan(self, v):
        terms = None if _is_complex(v.dtype) or _is_complex(self.dtype) else _terms_with_scalars(self)
        if terms is None:
            self._plan = None
            return False
        self._plan = _TreePlan(self, [(_scale_getter(sc), leaves) for sc, leaves in terms], v)
        return True


# ====================================================================================================
# BlockDiagonalOperator and WOperator   
# ====================================================================================================
class BlockDiagonalOperator(AbstractSciMLOperator):
    """src/block.jl:37-50: block b maps its slice of v to its slice of w; `mul!` :145-179 loops over the blocks with
    row views.  Here the whole operator is ONE call (scimlop_blockdiag_mul): equal dense blocks run as a single
    strided-batched GEMM (a row slice of the column-major v / w is a batch stride), ragged blocks as one launch
    each.  `cache_operator` re-homes equal dense constant blocks into one contiguous allocation so that the batched
    form applies (the analogue of `cache_internals`, :110-119)."""

    def __init__(self, *ops):
        if len(ops) == 1 and isinstance(ops[0], (tuple, list)):
            ops = tuple(ops[0])
        if not ops:
            raise AssertionError("BlockDiagonalOperator needs at least one block")   # src/block.jl:45
        self.ops = tuple(op if isinstance(op, AbstractSciMLOperator) else MatrixOperator(op) for op in ops)
        for op in self.ops:
            if not isinstance(op, _LEAVES) or (isinstance(op, MatrixOperator) and (op.banded or op.sparse)):
                raise NotImplementedError(f"BlockDiagonalOperator block {type(op).__name__}: the B200 backend takes "
                                          "Matrix / Diagonal / BatchedDiagonal / Identity / Null blocks "
                                          "(the Julia wrapper keeps the generic loop for anything else)")
        self._blocks = None
        self._packed = None

    @property
    def dtype(self):   # reduce(_promote_operator_eltype, ops), src/interface.jl:444
        dts = [op.dtype for op in self.ops if op.dtype is not None]
        return None if not dts else np.dtype(np.result_type(*dts))

    def size(self):   # src/block.jl:66-69
        return (sum(op.size()[0] for op in self.ops), sum(op.size()[1] for op in self.ops))

    def getops(self):
        return self.ops

    def adjoint(self):   # src/block.jl:77-80
        return BlockDiagonalOperator(*[op.adjoint() for op in self.ops])

    def iscached(self):
        return self._blocks is not None

    def cache_operator(self, v):
        if v.shape[0] != self.size()[1]:
            raise AssertionError(f"cache_operator: v has {v.shape[0]} rows, operator has {self.size()[1]} columns")
        ops = self.ops
        dense = [op for op in ops if isinstance(op, MatrixOperator)]
        if len(ops) > 1 and len(dense) == len(ops) and all(
                op.update_func is None and op.A.shape == ops[0].A.shape and op.trans == ops[0].trans
                and op.A.dtype == ops[0].A.dtype for op in ops):
            r, c = ops[0].A.shape
            packed = DeviceArray.empty((r, c * len(ops)), ops[0].A.dtype, v.device_index)
            for b, op in enumerate(ops):
                packed.cols(b * c, (b + 1) * c).t.copy_(op.A.t)
            self._packed = packed
            self.ops = tuple(MatrixOperator(packed.cols(b * c, (b + 1) * c), op.trans) for b, op in enumerate(ops))
        k = _ncols(v)
        arr = (_lib.Factor * len(self.ops))()
        for b, op in enumerate(self.ops):
            f = arr[b]
            f.rows, f.cols = op.size()
            f.trans, f.data, f.ld = 0, None, 0
            if isinstance(op, MatrixOperator):
                f.kind, f.trans, f.data, f.ld = _lib.DENSE, int(op.trans), op.A.ptr, op.A.ld
            elif isinstance(op, _VectorDiagonalOperator):
                f.kind, f.data = _lib.DIAG, op.diag.ptr
            elif isinstance(op, BatchedDiagonalOperator):
                if op.diag.shape[1] != k:
                    raise AssertionError(f"BatchedDiagonalOperator block: diag {op.diag.shape} does not match k={k}")
                f.kind, f.data, f.ld = _lib.BDIAG, op.diag.ptr, op.diag.ld
            elif isinstance(op, IdentityOperator):
                f.kind = _lib.IDENTITY
            else:
                f.kind = _lib.NULL
            if op.dtype is not None and op.dtype != v.dtype:
                raise TypeError(f"block eltype {op.dtype} != array eltype {v.dtype}")
        self._blocks = arr
        return self

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if self._blocks is None:
            self.cache_operator(v)
        h = _lib.handle(w.device_index)
        h.check(h.lib.scimlop_blockdiag_mul(h.ptr, _DT[w.dtype], self.precision, len(self.ops), self._blocks,
                                            C.c_void_p(v.ptr), v.ld, C.c_void_p(w.ptr), w.ld, _ncols(v),
                                            _scalar(w.dtype, alpha), _scalar(w.dtype, beta), _stream()),
                "scimlop_blockdiag_mul")
        return w


class _NegInvGamma(ScalarOperator):
    """The -1/γ (or -λ/γ for a UniformScaling mass matrix) of a WOperator, read at every mul!."""

    def __init__(self, owner, lam=1.0):
        self.owner, self.lam = owner, float(lam)
        self.update_func = None

    @property
    def val(self):
        return -self.lam / float(self.owner.gamma)

    @val.setter
    def val(self, x):
        pass

    def convert(self):
        return self.val


class WOperator(AbstractSciMLOperator):
    """src/woperator.jl:176-230: `W = J - M/γ`.  The reference's `mul!` (:447-481) is `mul!(Y, M, B)`, `lmul!(-1/γ, Y)`,
    `mul!(cache, J, B)`, `Y .+= cache`; here it is one tree call: the J product writes Y, the M product accumulates with
    α = -1/γ in its epilogue (a scalar mass matrix λI rides in the elementwise pass).  `gamma` is mutable, as in the
    reference (`update_coefficients!(W, u, p, t; gamma)`, :372-388)."""

    def __init__(self, mass_matrix, gamma, J):
        self.J = J if isinstance(J, AbstractSciMLOperator) else MatrixOperator(J)
        n = self.J.size()[0]
        if np.isscalar(mass_matrix):                       # UniformScaling λ*I
            self.mass_matrix = IdentityOperator(n)
            self._neg = _NegInvGamma(self, mass_matrix)
        else:
            self.mass_matrix = mass_matrix if isinstance(mass_matrix, AbstractSciMLOperator) else MatrixOperator(mass_matrix)
            self._neg = _NegInvGamma(self)
        if self.mass_matrix.size() != self.J.size():
            raise AssertionError(f"WOperator: mass matrix {self.mass_matrix.size()} vs J {self.J.size()}")
        self.gamma = gamma
        self._sum = AddedOperator(self.J, ScaledOperator(self._neg, self.mass_matrix))
        self._inv, self._inv_key, self._version = None, None, 0

    @property
    def dtype(self):
        return self.J.dtype

    def size(self):
        return self.J.size()

    def getops(self):
        return (self.J, self.mass_matrix)

    def iscached(self):
        return self._sum.iscached()

    def cache_operator(self, v):
        self._sum = self._sum.cache_operator(v)
        return self

    def update_coefficients_(self, u=None, p=None, t=None, gamma=None):
        self.J.update_coefficients_(u, p, t)
        self.mass_matrix.update_coefficients_(u, p, t)
        if getattr(self.J, "update_func", None) is not None or getattr(self.mass_matrix, "update_func", None) is not None:
            self._version += 1        # the cached inverse of the concrete form is stale
        if gamma is not None:
            self.gamma = gamma
        return self

    def _flat_terms(self):
        return self._sum._flat_terms()

    def mul_(self, w, v, alpha=None, beta=None):
        if not self._sum.iscached():
            self._sum = self._sum.cache_operator(v)
        return self._sum.mul_(w, v, alpha, beta)

    # ---- W \\ x: the reference converts W to a dense matrix and solves (src/woperator.jl:431-445).  Here the concrete
    #      form J - M/γ is assembled on the device and inverted once per (J, M, γ); every solve is then a product. ----
    def has_ldiv(self):
        return isinstance(self.J, MatrixOperator) and not self.J.banded and not self.J.sparse and \
            isinstance(self.mass_matrix, (MatrixOperator, IdentityOperator))

    def _concrete_inverse(self):
        key = (float(self.gamma), self._version)
        if self._inv is not None and self._inv_key == key:
            return self._inv
        if not self.has_ldiv():
            raise TypeError("WOperator ldiv!: J (and a non-scalar mass matrix) must be dense MatrixOperators")
        n = self.J.size()[0]
        dt = self.J.dtype
        h = _lib.handle(self.J.A.device_index)
        dense = DeviceArray.empty((n, n), dt, self.J.A.device_index)
        one, coef = _scalar(dt, 1.0), _scalar(dt, self._neg.convert())
        src = self.J.A
        if self.J.trans:            # materialise op(J): dense = J' via a product with the identity is overkill; transpose on copy
            dense.t.copy_(src.t.view(n, n).t().contiguous().view(-1))
        else:
            dense.t.copy_(src.t)
        if isinstance(self.mass_matrix, IdentityOperator):
            ones = DeviceArray.full((n,), 1.0, dt, self.J.A.device_index)
            # the diagonal of `dense` as a 1 x n matrix with leading dimension n + 1
            h.check(h.lib.scimlop_axpby(h.ptr, _DT[dt], 1, n, C.c_void_p(ones.ptr), 1, C.c_void_p(dense.ptr), n + 1, coef, one,
                                        _stream()), "scimlop_axpby")
        else:
            Mm = self.mass_matrix.A
            if self.mass_matrix.trans:
                Mm = DeviceArray(Mm.t.view(n, n).t().contiguous().view(-1), (n, n))
            h.check(h.lib.scimlop_axpby(h.ptr, _DT[dt], n, n, C.c_void_p(Mm.ptr), n, C.c_void_p(dense.ptr), n, coef, one,
                                        _stream()), "scimlop_axpby")
        self._inv = MatrixOperator(_invert_into(dense, dense))
        self._inv_key = key
        return self._inv

    def _ldiv(self, w, v, scale=None):
        F = self._concrete_inverse()
        src = v
        if w.t.data_ptr() == v.t.data_ptr():
            src = v.copy()
        return F.mul_(w, src, scale, None if scale is None else False)


# ====================================================================================================
# the solve side: InvertibleOperator / factorize / ldiv!   
# ====================================================================================================
class InvertibleOperator(AbstractSciMLOperator):
    """src/matrix.jl:501-511: `L` answers mul!, `F` answers ldiv! (:632-639).  In the reference `F` is a LAPACK
    factorisation; here it is an operator that APPLIES the inverse (for a dense factor: the explicit inverse
    formed once on the device by scimlop_invert), so that ldiv! runs on the same tensor-core path as mul!.
    `refresh` re-forms F after `update_coefficients!` changed L (the reference updates both, :575-579)."""

    def __init__(self, L, F, refresh=None, lu=None):
        if L.size() != F.size()[::-1]:
            raise AssertionError(f"InvertibleOperator: L is {L.size()}, F is {F.size()}")
        self.L, self.F, self._refresh = L, F, refresh
        self.lu = lu                  # LUFactorization of a Float64 matrix leaf (None otherwise)
        self._stash = None
        self._lu_ws = None

    @property
    def stable_solve(self):
        """True when ldiv! runs the factorised (substitution) solve instead of multiplying by the explicit inverse:
        an ill-conditioned Float64 factor (cond_1 > COND_FAST).  The lazy adjoint solves with the same blob through the
        transposed sweeps (U^T, L^T, then P^T)."""
        return self.lu is not None and self.lu.stable_needed

    @property
    def _trans(self):
        return bool(getattr(self.L, "trans", False))

    @property
    def dtype(self):
        return self.L.dtype

    def size(self):
        return self.L.size()

    def getops(self):
        return (self.L, self.F)

    def adjoint(self):   # src/matrix.jl:565
        return InvertibleOperator(self.L.adjoint(), self.F.adjoint(), lu=self.lu)

    def iscached(self):
        return self.L.iscached() and self.F.iscached()

    def cache_operator(self, v):   # src/matrix.jl:610-615
        self.L = self.L.cache_operator(v)
        self.F = self.F.cache_operator(v)
        return self

    def update_coefficients_(self, u=None, p=None, t=None):
        self.L.update_coefficients_(u, p, t)
        if self._refresh is not None:
            self._refresh()
        else:
            self.F.update_coefficients_(u, p, t)
        return self

    def _flat_terms(self):
        return self.L._flat_terms()

    def mul_(self, w, v, alpha=None, beta=None):   # src/matrix.jl:619-631
        return self.L.mul_(w, v, alpha, beta)

    def has_ldiv(self):
        return True

    def _ldiv(self, w, v, scale=None):
        src = v
        if w.t.data_ptr() == v.t.data_ptr():
            if self._stash is None or self._stash.shape != tuple(v.shape) or self._stash.dtype != v.dtype:
                self._stash = DeviceArray.empty(tuple(v.shape), v.dtype, v.device_index)
            self._stash.t.copy_(v.t)
            src = self._stash
        if self.stable_solve:
            k = _ncols(v)
            need = self.lu.n * k * 8
            if self._lu_ws is None or self._lu_ws.numel() < need:
                self._lu_ws = torch.empty(need, dtype=torch.uint8, device=v.t.device)
            h = _lib.handle(w.device_index)
            _kron_ldiv(h, [self.lu], [self.lu.n], src, w, k, self._lu_ws, [self._trans])
            if scale is not None and float(scale) != 1.0:
                _lmul(w, float(scale), h)
            return w
        return self.F.mul_(w, src, scale, None if scale is None else False)


class InvertedOperator(AbstractSciMLOperator):
    """src/basic.jl:1316-1336: lazy inverse; `mul!` is the wrapped operator's ldiv! (:1420-1442) and vice versa."""

    def __init__(self, L):
        if not L.has_ldiv():
            raise AssertionError("InvertedOperator: the operator has no ldiv! (factorize it first)")
        self.L = L
        self._tmp = None

    @property
    def dtype(self):
        return self.L.dtype

    def size(self):
        return self.L.size()[::-1]

    def getops(self):
        return (self.L,)

    def iscached(self):
        return self.L.iscached()

    def cache_operator(self, v):
        m = self.L.size()[1]
        self.L = self.L.cache_operator(DeviceArray.zeros((m,) + tuple(v.shape[1:]), v.dtype, v.device_index))
        return self

    def has_ldiv(self):
        return True

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if alpha is None:
            return self.L._ldiv(w, v)
        if _beta_is_zero(beta):
            return self.L._ldiv(w, v, alpha)
        # 5-arg: w = α (L \\ v) + β w through one scratch (the reference's copy!/axpby!, src/basic.jl:1428-1442)
        if self._tmp is None or self._tmp.shape != tuple(w.shape) or self._tmp.dtype != w.dtype:
            self._tmp = DeviceArray.empty(tuple(w.shape), w.dtype, w.device_index)
        self.L._ldiv(self._tmp, v)
        h = _lib.handle(w.device_index)
        h.check(h.lib.scimlop_axpby(h.ptr, _DT[w.dtype], w.shape[0], _ncols(w), C.c_void_p(self._tmp.ptr), self._tmp.ld,
                                    C.c_void_p(w.ptr), w.ld, _scalar(w.dtype, alpha), _scalar(w.dtype, beta), _stream()),
                "scimlop_axpby")
        return w

    def _ldiv(self, w, v, scale=None):   # src/basic.jl:1444-1450
        if w.t.data_ptr() == v.t.data_ptr():
            raise NotImplementedError("two-argument ldiv! of an InvertedOperator (a mul! in place)")
        return self.L.mul_(w, v, scale, None if scale is None else False)


def invert_matrix(A):
    """inv(A) of a square device matrix via scimlop_invert (Gauss-Jordan, partial pivoting, on the device)."""
    A = _as_device(A)
    n, n2 = A.shape
    if n != n2:
        raise AssertionError("factorize: only square factors have an ldiv! on this path (test/matrix.jl:646)")
    out = DeviceArray.empty((n, n), A.dtype, A.device_index)
    _invert_into(A, out)
    return out


def _invert_into(A, out):
    h = _lib.handle(A.device_index)
    n = A.shape[0]
    nbytes = C.c_size_t(0)
    h.check(h.lib.scimlop_invert_workspace_bytes(_DT[A.dtype], n, C.byref(nbytes)), "scimlop_invert_workspace_bytes")
    ws = torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=A.t.device)
    h.check(h.lib.scimlop_invert(h.ptr, _DT[A.dtype], n, C.c_void_p(A.ptr), A.ld, C.c_void_p(out.ptr), out.ld,
                                 C.c_void_p(ws.data_ptr()), nbytes.value, _stream()), "scimlop_invert")
    return out


COND_FAST = 1.0e3   # cond_1 up to which ldiv! multiplies by the explicit inverse (residual cond * eps <= 1e-13)


class LUFactorization:
    """The device-side `lu(A)` of a Float64 matrix (scimlop_factorize): the opaque solve blob for scimlop_kron_ldiv, the
    explicit inverse and cond_1(A).  `refresh()` re-factorises after the matrix changed in place."""

    def __init__(self, A):
        self.A = A
        n = A.shape[0]
        if A.shape[1] != n:
            raise AssertionError("factorize: only square factors have an ldiv! on this path (test/matrix.jl:646)")
        self.n = n
        h = _lib.handle(A.device_index)
        fb, sbytes = C.c_size_t(0), C.c_size_t(0)
        h.check(h.lib.scimlop_factorize_bytes(_DT[A.dtype], n, C.byref(fb), C.byref(sbytes)), "scimlop_factorize_bytes")
        self.blob = torch.empty(max(1, fb.value), dtype=torch.uint8, device=A.t.device)
        self._fb, self._sb = fb.value, sbytes.value
        self.inv = DeviceArray.empty((n, n), A.dtype, A.device_index)
        self.cond = float("nan")
        self.refresh()

    def refresh(self):
        A = self.A
        h = _lib.handle(A.device_index)
        scratch = torch.empty(max(1, self._sb), dtype=torch.uint8, device=A.t.device)
        cond = C.c_double(0.0)
        h.check(h.lib.scimlop_factorize(h.ptr, _DT[A.dtype], self.n, C.c_void_p(A.ptr), A.ld, C.c_void_p(self.blob.data_ptr()),
                                        self._fb, C.c_void_p(self.inv.ptr), self.inv.ld, C.c_void_p(scratch.data_ptr()),
                                        self._sb, C.byref(cond), _stream()), "scimlop_factorize")
        self.cond = cond.value
        return self

    @property
    def stable_needed(self):
        return not (self.cond <= COND_FAST)       # NaN-safe: unknown conditioning takes the substitution path


def _kron_ldiv(h, facts, dims, v, w, k, ws, trans=None):
    """scimlop_kron_ldiv on LUFactorization blobs (None = identity factor; trans[i]: solve with the transposed factor)."""
    nf = len(dims)
    cf, cd, ck = (C.c_void_p * nf)(), (C.c_int64 * nf)(), (C.c_int32 * nf)()
    for i, (f, n) in enumerate(zip(facts, dims)):
        cd[i] = n
        ck[i] = _lib.IDENTITY if f is None else (_lib.DENSE | (_lib.TRANS if trans is not None and trans[i] else 0))
        cf[i] = None if f is None else f.blob.data_ptr()
    h.check(h.lib.scimlop_kron_ldiv(h.ptr, _lib.F64, nf, cf, cd, ck, C.c_void_p(v.ptr), C.c_void_p(w.ptr), k,
                                    C.c_void_p(ws.data_ptr()), ws.numel(), _stream()), "scimlop_kron_ldiv")
    return w


def factorize(L):
    """`LinearAlgebra.factorize(L)`: src/matrix.jl:513-516 (leaf), src/basic.jl:500 (Scaled), :1056 (Composed),
    src/tensor.jl:227 (every Kronecker factor).  Singular factors raise ScimlopError(ERR_SINGULAR).  Float64 matrices
    get a blocked LU on the device (+ the explicit inverse and cond_1 from it); Float32 keeps scimlop_invert."""
    if isinstance(L, InvertibleOperator):
        return L
    if isinstance(L, MatrixOperator):
        if L.banded or L.sparse:
            raise NotImplementedError("factorize of a banded / sparse MatrixOperator (a sparse solve) is not on this path")
        if L.A.dtype == np.float64:
            lu = LUFactorization(L.A)
            refresh = lu.refresh if L.update_func is not None else None
            return InvertibleOperator(L, MatrixOperator(lu.inv, L.trans), refresh, lu=lu)
        inv = invert_matrix(L.A)
        refresh = (lambda: _invert_into(L.A, inv)) if L.update_func is not None else None
        return InvertibleOperator(L, MatrixOperator(inv, L.trans), refresh)
    if isinstance(L, _VectorDiagonalOperator):
        rec = DeviceArray.empty(L.diag.shape, L.diag.dtype, L.diag.device_index)
        ones = DeviceArray.full(L.diag.shape, 1.0, L.diag.dtype, L.diag.device_index)
        fill = lambda: _diag_div(L.diag, 0, rec.reshape(rec.shape[0], 1), ones.reshape(ones.shape[0], 1), None)
        fill()
        return InvertibleOperator(L, _VectorDiagonalOperator(rec), fill if L.update_func is not None else None)
    if isinstance(L, (BatchedDiagonalOperator, IdentityOperator)):
        return L                                   # they answer ldiv! themselves
    if isinstance(L, ScaledOperator):
        return ScaledOperator(L.lam, factorize(L.L))
    if isinstance(L, ComposedOperator):
        return ComposedOperator(*[factorize(op) for op in L.ops])
    if isinstance(L, TensorProductOperator):
        return TensorProductOperator(*[factorize(op) for op in L.ops])
    raise TypeError(f"factorize: {type(L).__name__} cannot be factorised on this path "
                    "(the reference converts to a dense matrix, src/matrix.jl:513-516)")


def inv(L):
    """`inv(L)`: lazy InvertedOperator (src/basic.jl:1316-1336)."""
    return L.L if isinstance(L, InvertedOperator) else InvertedOperator(L)


def has_ldiv(L):
    return L.has_ldiv()


def ldiv_(*args):
    """`ldiv!(w, L, v)` or `ldiv!(L, v)`; returns the destination."""
    if len(args) == 3:
        w, L, v = args
        return L.ldiv_(w, v)
    if len(args) == 2:
        L, v = args
        return L.ldiv_(v)
    raise TypeError("ldiv!: ldiv_(w, L, v) or ldiv_(L, v)")


# ====================================================================================================
# lazy algebra
# ====================================================================================================
class ScaledOperator(_TreeMixin, AbstractSciMLOperator):
    """src/basic.jl:357-372; `mul!` :518-536: `mul!(w, L.L, v, λ·α, β)`, `iszero(λ)` short-circuits."""

    def __init__(self, lam, L):
        self.lam = lam if isinstance(lam, ScalarOperator) else ScalarOperator(lam)
        self.L = L

    @property
    def dtype(self):
        return self.L.dtype

    def size(self):
        return self.L.size()

    def getops(self):
        return (self.lam, self.L)

    def iscached(self):
        return self.L.iscached()

    def adjoint(self):
        return ScaledOperator(self.lam, self.L.adjoint())  # src/basic.jl:434-439 (real λ)

    def cache_operator(self, v):
        self.L = self.L.cache_operator(v)
        self._try_plan(v)
        return self

    def has_ldiv(self):
        return self.L.has_ldiv() and self.lam.convert() != 0

    def _ldiv(self, w, v, scale=None):  # src/basic.jl:538-546: ldiv!(w, L.L, v); ldiv!(λ, w) -- folded into one pass
        return self.L._ldiv(w, v, (1.0 if scale is None else float(scale)) / self.lam.convert())

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if self._plan is not None:
            return self._plan.mul_(w, v, alpha, beta, self.precision)
        lam = self.lam.convert()
        if lam == 0:
            return _lmul(w, beta, _lib.handle(w.device_index))
        a = lam if alpha is None else lam * _num(alpha)
        return self.L.mul_(w, v, a, False if beta is None else beta)


class AddedOperator(_TreeMixin, AbstractSciMLOperator):
    """src/basic.jl:596-610; nested sums are flattened (:619-635); `mul!` :834-854."""

    def __init__(self, *ops):
        flat = []
        for op in ops:
            flat.extend(op.ops if isinstance(op, AddedOperator) else [op])
        sizes = {op.size() for op in flat}
        if len(sizes) != 1:
            raise AssertionError(f"AddedOperator: size mismatch {sizes}")
        self.ops = tuple(flat)

    @property
    def dtype(self):   # reduce(_promote_operator_eltype, ops), src/interface.jl:444
        dts = [op.dtype for op in self.ops if op.dtype is not None]
        return None if not dts else np.dtype(np.result_type(*dts))

    def size(self):
        return self.ops[0].size()

    def getops(self):
        return self.ops

    def adjoint(self):
        return AddedOperator(*[op.adjoint() for op in self.ops])  # src/basic.jl:766-771

    def _as_kronsum(self):
        """`kron(A, I) + kron(I, B)` written as an AddedOperator of two TensorProductOperators (how config 5 and
        the reference's docs spell a Laplacian) is a Kronecker sum: one fused pass instead of two plus an accumulate."""
        if len(self.ops) != 2 or not all(isinstance(op, TensorProductOperator) for op in self.ops):
            return None
        (a1, b1), (a2, b2) = self.ops[0].ops, self.ops[1].ops
        for (A, I1), (I2, B) in (((a1, b1), (a2, b2)), ((a2, b2), (a1, b1))):
            if isinstance(I1, IdentityOperator) and isinstance(I2, IdentityOperator) and \
                    isinstance(A, MatrixOperator) and isinstance(B, MatrixOperator) and \
                    A.size() == (I2.len, I2.len) and B.size() == (I1.len, I1.len) and \
                    A.update_func is None and B.update_func is None:
                S = TensorSumOperator(A, B)
                if S._native:
                    return S
        return None

    def cache_operator(self, v):
        self._kronsum = self._as_kronsum()
        if self._kronsum is not None:
            self._kronsum = self._kronsum.cache_operator(v)
            self._plan = None
            return self
        self.ops = tuple(op.cache_operator(v) for op in self.ops)
        self._try_plan(v)
        return self

    def iscached(self):
        return getattr(self, "_kronsum", None) is not None or super().iscached()

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if getattr(self, "_kronsum", None) is not None:
            return self._kronsum.mul_(w, v, alpha, beta)
        if self._plan is not None:
            return self._plan.mul_(w, v, alpha, beta, self.precision)
        # summands that are not native leaves (e.g. TensorProductOperators): accumulate like
        # src/basic.jl:839-840 / :875-885 -- first summand carries β, the rest add with β = 1
        a = True if alpha is None else alpha
        self.ops[0].mul_(w, v, a, False if beta is None else beta)
        for op in self.ops[1:]:
            op.mul_(w, v, a, True)
        return w


class ComposedOperator(_TreeMixin, AbstractSciMLOperator):
    """src/basic.jl:909-927; `cache_self` :1104-1138; `mul!` :1160-1207."""

    def __init__(self, *ops):
        flat = []
        for op in ops:
            flat.extend(op.ops if isinstance(op, ComposedOperator) else [op])
        for a, b in zip(flat[:-1], flat[1:]):
            if a.size()[1] != b.size()[0]:
                raise AssertionError(f"ComposedOperator: size mismatch {a.size()} * {b.size()}")
        self.ops = tuple(flat)
        self.cache = None

    @property
    def dtype(self):   # reduce(_promote_operator_eltype, ops), src/interface.jl:444
        dts = [op.dtype for op in self.ops if op.dtype is not None]
        return None if not dts else np.dtype(np.result_type(*dts))

    def size(self):
        return (self.ops[0].size()[0], self.ops[-1].size()[1])

    def getops(self):
        return self.ops

    def adjoint(self):
        return ComposedOperator(*[op.adjoint() for op in reversed(self.ops)])  # src/basic.jl:1003-1011

    def iscached(self):
        return (self._plan is not None or self.cache is not None) and all(op.iscached() for op in self.ops)

    def cache_operator(self, v):
        if self._try_plan(v):
            self.ops = tuple(self.ops)
            # children are leaves here; nothing else to cache
            return self
        # generic chain: one scratch per intermediate (src/basic.jl:1104-1138), the last slot stashes w
        k = v.shape[1:]
        vecs = [v]
        for op in reversed(self.ops[1:]):
            vecs.insert(0, DeviceArray.zeros((op.size()[0],) + tuple(k), v.dtype, v.device_index))
        self.cache = vecs[:-1]  # no stash slot: α and β ride on the last stage instead of copy!/axpy! sweeps
        n = len(self.ops)
        inputs = vecs  # input of op i is vecs[i]
        self.ops = tuple(self.ops[i].cache_operator(inputs[i]) for i in range(n))
        return self

    def has_ldiv(self):
        return all(op.has_ldiv() for op in self.ops)

    def _ldiv(self, w, v, scale=None):
        """src/basic.jl:1209-1256: (A∘B∘C) \\ v = C \\ (B \\ (A \\ v)); intermediates ping-pong through two scratch
        arrays (all factors must be square here, as `factorize` requires)."""
        n = len(self.ops)
        if n == 1:
            return self.ops[0]._ldiv(w, v, scale)
        if getattr(self, "_ldiv_scratch", None) is None or self._ldiv_scratch[0].shape != tuple(v.shape) \
                or self._ldiv_scratch[0].dtype != v.dtype:
            self._ldiv_scratch = [DeviceArray.zeros(tuple(v.shape), v.dtype, v.device_index) for _ in range(2)]
        cur = v
        for i, op in enumerate(self.ops):
            last = i == n - 1
            out = w if last else self._ldiv_scratch[i & 1]
            op._ldiv(out, cur, scale if last else None)
            cur = out
        return w

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if not self.iscached():
            raise AssertionError("cache needs to be set up for operator of type ComposedOperator: "
                                 "call cache_operator(L, v) first")  # src/basic.jl:1183,1198
        if self._plan is not None:
            return self._plan.mul_(w, v, alpha, beta, self.precision)
        n = len(self.ops)
        if alpha is None:  # 3-arg: chain through the scratch, src/basic.jl:1172-1178
            vecs = [w] + self.cache[: n - 1] + [v]
            for i in reversed(range(n)):
                self.ops[i].mul_(vecs[i], vecs[i + 1])
            return w
        # 5-arg without the reference's copy!/lmul!/axpy! sweeps: α and β ride on the last stage
        vecs = [w] + self.cache[: n - 1] + [v]
        for i in reversed(range(1, n)):
            self.ops[i].mul_(vecs[i], vecs[i + 1])
        return self.ops[0].mul_(w, vecs[1], alpha, beta)


# ====================================================================================================
# tensor products
# ====================================================================================================
def _kron_leaf(op):
    """(kind, data_ptr, ld) if `op` can be a native Kronecker factor, else None."""
    while isinstance(op, InvertibleOperator):
        op = op.L
    if isinstance(op, MatrixOperator) and op.sparse:
        # SCIMLOP_CSR: factors[f] is a host descriptor of the device arrays (kept alive on the operator).  The lazy adjoint
        # (constant by construction) is converted once to the CSR form of A^T; complex sparse factors are not built.
        if _is_complex(op.A.dtype):
            raise NotImplementedError("a complex sparse MatrixOperator as a Kronecker factor is not supported by the B200 backend")
        A = op.A
        if op.trans:
            if getattr(op, "_csr_T", None) is None:
                op._csr_T = _csr_transpose(op.A)
            A = op._csr_T
        d = getattr(op, "_csr_desc", None)
        if d is None:                               # ONE struct per operator: native plans keep its address
            d = op._csr_desc = _lib.CsrFactor()
        d.rowptr, d.colidx, d.values, d.index_base, d.reserved = A.rowptr.ptr, A.colidx.ptr, A.values.ptr, 0, 0
        return (_lib.CSR, C.addressof(d), 0)
    if isinstance(op, MatrixOperator) and op.banded:
        return (_lib.BANDED, op.A.bands.ptr, op.A.b)      # lds[] carries the half-bandwidth for banded factors
    if isinstance(op, MatrixOperator):
        return (_lib.DENSE | (_lib.TRANS if op.trans else 0), op.A.ptr, op.A.ld)
    if isinstance(op, IdentityOperator):
        return (_lib.IDENTITY, None, 0)
    if isinstance(op, _VectorDiagonalOperator):
        return (_lib.DIAG, op.diag.ptr, 0)
    return None


def _promote_leaf(op, dtype):
    """Mixed element types (`_promote_operator_eltype`, src/interface.jl:444; src/tensor.jl:99): a constant dense /
    diagonal leaf whose eltype differs from the arrays' is applied through a converted copy made once at cache time
    (Float32 -> Float64 is exact; Float64 -> Float32 rounds the operator to the arrays' precision).  Time-dependent
    leaves keep the error: their copy would go stale."""
    if op.dtype is None or op.dtype == dtype:
        return op
    tdt = {np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32, np.dtype("complex64"): torch.complex64,
           np.dtype("complex128"): torch.complex128}[np.dtype(dtype)]
    if isinstance(op, MatrixOperator) and not (op.banded or op.sparse) and op.update_func is None:
        return MatrixOperator(DeviceArray(op.A.t.to(tdt), op.A.shape), op.trans)
    if isinstance(op, _VectorDiagonalOperator) and op.update_func is None:
        return _VectorDiagonalOperator(DeviceArray(op.diag.t.to(tdt), op.diag.shape))
    raise TypeError(f"operator eltype {op.dtype} != array eltype {np.dtype(dtype)} (only constant dense / diagonal "
                    "leaves are promoted)")


class TensorProductOperator(AbstractSciMLOperator):
    """src/tensor.jl:88-107.  `ops = (outer, inner)`; matrices are wrapped in MatrixOperator (:109-118); the
    N-ary constructor is a right fold (:121).  `mul!` :543-610 needs `cache_operator` first (:548,:583)."""

    def __init__(self, *ops):
        ops = [op if isinstance(op, AbstractSciMLOperator) else MatrixOperator(op) for op in ops]
        if len(ops) < 2:
            raise ValueError("TensorProductOperator needs at least two factors")
        if len(ops) > 2:
            ops = [ops[0], TensorProductOperator(*ops[1:])]
        self.ops = tuple(ops)
        self._native = None   # (factors..., workspace) when the whole product is native
        self._native_inv = None   # the same for ldiv!: the inverse factors of a factorised product
        self.cache = None     # c1 of the reference when falling back to the two-stage algorithm
        self._k = None

    @property
    def dtype(self):   # reduce(_promote_operator_eltype, ops), src/tensor.jl:99
        dts = [op.dtype for op in self.ops if op.dtype is not None]
        return None if not dts else np.dtype(np.result_type(*dts))

    def size(self):
        (mo, no), (mi, ni) = self.ops[0].size(), self.ops[1].size()
        return (mo * mi, no * ni)  # src/tensor.jl:179

    def getops(self):
        return self.ops

    def adjoint(self):
        return TensorProductOperator(*[op.adjoint() for op in self.ops])  # src/tensor.jl:181-191

    # ---- flattening into native factors (with the ScaledOperator λ's pulled out) ----------------------
    def _flatten(self):
        """([(leaf op)...], [scalars...]) outermost first, or None if some factor is not native."""
        leaves, scalars = [], []

        def walk(op):
            if isinstance(op, ScaledOperator):
                scalars.append(op.lam)
                return walk(op.L)
            if isinstance(op, TensorProductOperator):
                return all(walk(o) for o in op.ops)
            if _kron_leaf(op) is None:
                return False
            leaves.append(op)
            return True

        return (leaves, scalars) if walk(self) else None

    def iscached(self):
        return self._native is not None or (self.cache is not None and all(op.iscached() for op in self.ops))

    def cache_operator(self, v):
        """`cache_self` + `cache_internals` (src/tensor.jl:460-541).  Native path: ask the library for its
        workspace (<= 2 ping-pong buffers instead of the reference's 7 scratch tensors)."""
        k = _ncols(v)
        if v.shape[0] != self.size()[1]:
            raise AssertionError(f"cache_operator: v has {v.shape[0]} rows, operator has {self.size()[1]} columns")
        self._k = k
        flat = self._flatten()
        if flat is not None:
            leaves, scalars = flat
            leaves = [_promote_leaf(op, v.dtype) for op in leaves]
            self._promoted = leaves      # keeps converted copies alive
            nf = len(leaves)
            dims = (C.c_int64 * (2 * nf))()
            lds = (C.c_int64 * nf)()
            kinds = (C.c_int32 * nf)()
            ptrs = (C.c_void_p * nf)()
            for i, op in enumerate(leaves):
                m, n = op.size()
                dims[2 * i], dims[2 * i + 1] = m, n
                kinds[i], ptrs[i], lds[i] = _kron_leaf(op)
                if isinstance(op, MatrixOperator):
                    op._attach_split()
            nbytes = C.c_size_t(0)
            lib = _lib.load()
            rc = lib.scimlop_kron_workspace_bytes(_DT[v.dtype], nf, dims, kinds, k, C.byref(nbytes))
            if rc != _lib.OK:
                raise _lib.ScimlopError(rc, "scimlop_kron_workspace_bytes")
            ws = torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=v.t.device)
            self._native = dict(nf=nf, dims=dims, lds=lds, kinds=kinds, ptrs=ptrs, ws=ws, ws_bytes=nbytes.value,
                                scalars=scalars, dtype=v.dtype)
            return self
        # two-stage algorithm of the reference with the native outer stage (the `tensor_outer_mul_fast!` hook)
        outer, inner = self.ops
        mi, ni = inner.size()
        mo, no = outer.size()
        self.cache = DeviceArray.zeros((mi, no * k), v.dtype, v.device_index)     # c1, src/tensor.jl:460-476
        inner = inner.cache_operator(v.reshape(ni, no * k))                        # :538
        outer = outer.cache_operator(DeviceArray.zeros((no, mi * k), v.dtype, v.device_index))  # :536,:539
        self.ops = (outer, inner)
        return self

    # ---- ldiv!: src/tensor.jl:612-673 + outer_div! :864-933 ---------------------------------------------
    def _flatten_inverse(self):
        """Like _flatten, but every leaf is replaced by the operator that applies its inverse."""
        leaves, scalars = [], []
        self._inv_owners = []     # the InvertibleOperator (or None for an identity) behind every leaf

        def walk(op):
            if isinstance(op, ScaledOperator):
                scalars.append(op.lam)
                return walk(op.L)
            if isinstance(op, TensorProductOperator):
                return all(walk(o) for o in op.ops)
            if isinstance(op, IdentityOperator):
                leaves.append(op)
                self._inv_owners.append(None)
                return True
            if isinstance(op, InvertibleOperator) and _kron_leaf(op.F) is not None and not isinstance(op.F, InvertibleOperator):
                leaves.append(op.F)
                self._inv_owners.append(op)
                return True
            return False

        return (leaves, scalars) if walk(self) else None

    def has_ldiv(self):
        return self._flatten_inverse() is not None

    def _ldiv(self, w, v, scale=None):
        """(A ⊗ B) \\ v = (A⁻¹ ⊗ B⁻¹) v: with the inverse factors cached by `factorize` this is scimlop_kron_mul on
        them -- the same two tensor-core mode products as mul!, no permutedims!, no triangular sweeps."""
        if not self.iscached():
            raise AssertionError("cache needs to be set up for operator of type TensorProductOperator: "
                                 "call cache_operator(L, v) first")  # src/tensor.jl:617,653
        k = _ncols(v)
        if k != self._k:
            raise AssertionError("operator was cached for a different number of columns: call cache_operator again")
        if self._native_inv is None or self._native_inv["dtype"] != v.dtype:
            flat = self._flatten_inverse()
            if flat is None:
                raise TypeError("ldiv!(::TensorProductOperator): every factor must be factorised "
                                "(factorize(L), src/tensor.jl:227) or an identity")
            leaves, scalars = flat
            nf = len(leaves)
            dims, lds = (C.c_int64 * (2 * nf))(), (C.c_int64 * nf)()
            kinds, ptrs = (C.c_int32 * nf)(), (C.c_void_p * nf)()
            ndense = 0
            for i, op in enumerate(leaves):
                m, n = op.size()
                if m != n:
                    raise AssertionError("ldiv!(::TensorProductOperator) needs square factors")
                dims[2 * i], dims[2 * i + 1] = m, n
                kinds[i], ptrs[i], lds[i] = _kron_leaf(op)
                ndense += (kinds[i] & 0xFF) in (_lib.DENSE, _lib.BANDED)
            nbytes = C.c_size_t(0)
            rc = _lib.load().scimlop_kron_workspace_bytes(_DT[v.dtype], nf, dims, kinds, k, C.byref(nbytes))
            if rc != _lib.OK:
                raise _lib.ScimlopError(rc, "scimlop_kron_workspace_bytes")
            fwd = self._native
            ws = fwd["ws"] if fwd is not None and fwd["ws_bytes"] >= nbytes.value else \
                torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=v.t.device)
            self._native_inv = dict(nf=nf, dims=dims, lds=lds, kinds=kinds, ptrs=ptrs, ws=ws, ws_bytes=nbytes.value,
                                    scalars=scalars, dtype=v.dtype, nstages=sum(1 for i in range(nf) if kinds[i] != _lib.IDENTITY),
                                    ndense=ndense, stash=None, owners=list(self._inv_owners), lu_ws=None)
        nat = self._native_inv
        a = scale
        for s in nat["scalars"]:
            a = (1.0 if a is None else float(a)) / s.convert()
        src = v
        owners = nat["owners"]
        if any(o is not None and o.stable_solve for o in owners) and \
                all(o is None or o.lu is not None for o in owners):
            # an ill-conditioned factor: the factorised solve (substitution sweeps as GEMMs along every mode)
            need = v.t.numel() * 8
            if nat["lu_ws"] is None or nat["lu_ws"].numel() < need:
                nat["lu_ws"] = torch.empty(need, dtype=torch.uint8, device=v.t.device)
            if w.t.data_ptr() == v.t.data_ptr():
                if nat["stash"] is None or nat["stash"].shape != tuple(v.shape):
                    nat["stash"] = DeviceArray.empty(tuple(v.shape), v.dtype, v.device_index)
                nat["stash"].t.copy_(v.t)
                src = nat["stash"]
            h = _lib.handle(w.device_index)
            dims_n = [leaf.size()[0] for leaf in self._flatten_inverse()[0]]
            _kron_ldiv(h, [None if o is None else o.lu for o in owners], dims_n, src, w, k, nat["lu_ws"],
                       [o is not None and o._trans for o in owners])
            if a is not None and float(a) != 1.0:
                _lmul(w, float(a), h)
            return w
        if w.t.data_ptr() == v.t.data_ptr() and nat["nstages"] == 1 and nat["ndense"] == 1:
            # two-argument form with a single GEMM stage: a GEMM cannot run in place, stage v once
            if nat["stash"] is None or nat["stash"].shape != tuple(v.shape):
                nat["stash"] = DeviceArray.empty(tuple(v.shape), v.dtype, v.device_index)
            nat["stash"].t.copy_(v.t)
            src = nat["stash"]
        h = _lib.handle(w.device_index)
        n = self.size()[1]
        h.check(h.lib.scimlop_kron_mul(h.ptr, _DT[w.dtype], self.precision, nat["nf"], nat["ptrs"], nat["dims"],
                                       nat["lds"], nat["kinds"], C.c_void_p(src.ptr), n, C.c_void_p(w.ptr), n, k,
                                       _scalar(w.dtype, a), None, C.c_void_p(nat["ws"].data_ptr()), nat["ws_bytes"],
                                       _stream()), "scimlop_kron_mul")
        return w

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if not self.iscached():
            raise AssertionError("cache needs to be set up for operator of type TensorProductOperator: "
                                 "call cache_operator(L, v) first")  # src/tensor.jl:548,583
        k = _ncols(v)
        if k != self._k:
            raise AssertionError("operator was cached for a different number of columns: call cache_operator again")
        h = _lib.handle(w.device_index)
        if self._native is not None:
            nat = self._native
            a = alpha
            for s in nat["scalars"]:  # ScaledOperator factors fold into α (the intended src/tensor.jl:800-804)
                a = s.convert() * (1.0 if a is None else _num(a))
            m, n = self.size()
            h.check(h.lib.scimlop_kron_mul(h.ptr, _DT[w.dtype], self.precision, nat["nf"], nat["ptrs"], nat["dims"],
                                           nat["lds"], nat["kinds"], C.c_void_p(v.ptr), n, C.c_void_p(w.ptr), m, k,
                                           _scalar(w.dtype, a), _scalar(w.dtype, beta),
                                           C.c_void_p(nat["ws"].data_ptr()), nat["ws_bytes"], _stream()),
                    "scimlop_kron_mul")
            return w
        # ---- reference algorithm, src/tensor.jl:543-610 ----
        outer, inner = self.ops
        mi, ni = inner.size()
        mo, no = outer.size()
        U = v.reshape(ni, no * k)
        if isinstance(outer, IdentityOperator):
            inner.mul_(w.reshape(mi, no * k), U, alpha, beta)                      # :565 / :600
            return w
        C1 = self.cache
        inner.mul_(C1, U)                                                          # :568 / :603
        lam = 1.0
        while isinstance(outer, ScaledOperator):                                   # :755-759 / :800-804 (intended)
            lam *= outer.lam.convert()
            outer = outer.L
        if not isinstance(outer, MatrixOperator) or outer.banded or outer.sparse:
            raise NotImplementedError(f"outer factor {type(outer).__name__} is not supported by the B200 backend "
                                      "(the Julia wrapper keeps the generic path for it)")
        a = alpha if lam == 1.0 else lam * (1.0 if alpha is None else _num(alpha))
        if outer.trans:
            # W_j = C1_j * op(A)^T with op(A) = A^T  ->  plain (non-transposed B) batched GEMM
            h.check(h.lib.scimlop_gemm_strided_batched(
                h.ptr, _DT[w.dtype], self.precision, 0, 0, mi, mo, no, _scalar(w.dtype, a), C.c_void_p(C1.ptr), mi,
                mi * no, C.c_void_p(outer.A.ptr), outer.A.ld, 0, _scalar(w.dtype, beta), C.c_void_p(w.ptr), mi,
                mi * mo, k, _stream()), "scimlop_gemm_strided_batched")
        else:
            h.check(h.lib.scimlop_outer_mul_fast(h.ptr, _DT[w.dtype], self.precision, C.c_void_p(w.ptr),
                                                 C.c_void_p(outer.A.ptr), outer.A.ld, C.c_void_p(C1.ptr), mi, mo, no,
                                                 k, _scalar(w.dtype, a), _scalar(w.dtype, beta), _stream()),
                    "scimlop_outer_mul_fast")                                      # :773-776 / :816-819
        return w


class TensorSumOperator(AbstractSciMLOperator):
    """src/tensor.jl:269-298: `outer (x) I + I (x) inner`; `mul!` :387-403."""

    def __init__(self, outer, inner):
        outer = outer if isinstance(outer, AbstractSciMLOperator) else MatrixOperator(outer)
        inner = inner if isinstance(inner, AbstractSciMLOperator) else MatrixOperator(inner)
        for op in (outer, inner):
            m, n = op.size()
            if m != n:
                raise AssertionError("TensorSumOperator needs square factors")
        self.ops = (outer, inner)
        both = isinstance(outer, MatrixOperator) and isinstance(inner, MatrixOperator) and not (outer.sparse or inner.sparse)
        self._banded = both and outer.banded and inner.banded
        self._native = both and not outer.trans and not inner.trans and (self._banded or not (outer.banded or inner.banded))
        self.products = None
        if not self._native:
            self.products = (TensorProductOperator(outer, IdentityOperator(inner.size()[0])),
                             TensorProductOperator(IdentityOperator(outer.size()[0]), inner))
        self._cached = False

    @property
    def dtype(self):
        return self.ops[0].dtype or self.ops[1].dtype

    def size(self):
        n = self.ops[0].size()[0] * self.ops[1].size()[0]
        return (n, n)

    def getops(self):
        return self.ops

    def iscached(self):
        return self._cached

    def cache_operator(self, v):
        if not self._native:
            self.products = tuple(p.cache_operator(v) for p in self.products)
        self._cached = True
        return self

    def mul_(self, w, v, alpha=None, beta=None):
        _check_vw(self, w, v)
        if not self._native:
            a = True if alpha is None else alpha
            self.products[0].mul_(w, v, a, False if beta is None else beta)      # :388 / :400
            self.products[1].mul_(w, v, a, True)                                 # :389 / :401
            return w
        h = _lib.handle(w.device_index)
        A, B = self.ops[0].A, self.ops[1].A
        mo, mi = A.shape[0], B.shape[0]
        n = mo * mi
        if self._banded:   # finite-difference Laplacian: ONE HBM-bound pass
            h.check(h.lib.scimlop_kronsum_banded_mul(h.ptr, _DT[w.dtype], C.c_void_p(A.bands.ptr), A.b, mo,
                                                     C.c_void_p(B.bands.ptr), B.b, mi, C.c_void_p(v.ptr), n,
                                                     C.c_void_p(w.ptr), n, _ncols(v), _scalar(w.dtype, alpha),
                                                     _scalar(w.dtype, beta), _stream()), "scimlop_kronsum_banded_mul")
            return w
        h.check(h.lib.scimlop_kronsum_mul(h.ptr, _DT[w.dtype], self.precision, C.c_void_p(A.ptr), A.ld, mo,
                                          C.c_void_p(B.ptr), B.ld, mi, C.c_void_p(v.ptr), n, C.c_void_p(w.ptr), n,
                                          _ncols(v), _scalar(w.dtype, alpha), _scalar(w.dtype, beta), _stream()),
                "scimlop_kronsum_mul")
        return w


# ====================================================================================================
# module-level protocol (the reference's free functions)
# ====================================================================================================
def kron(*ops):
    """`kron(A, B, ...)` / `⊗` over operators or matrices: src/tensor.jl:149-164."""
    return TensorProductOperator(*ops)


def kronsum(outer, inner):
    """src/tensor.jl:334."""
    return TensorSumOperator(outer, inner)


def cache_operator(L, v):
    L = L.cache_operator(v)
    try:
        L._cached_for = (tuple(v.shape), v.dtype, v.device_index)
    except AttributeError:
        pass
    return L


def copy(L):
    """`Base.copy(L)`: a deep copy that shares no storage with L -- operator arrays AND caches
    (src/tensor.jl:210-215 copies the ops and deep-copies the cache; the leaves copy their arrays).  A cached operator
    yields a cached copy (its workspace is allocated anew for the same input shape)."""
    new = adapt(L, to=lambda a: DeviceArray(a.t.clone(), a.shape))
    for attr in ("precision",):
        if attr in getattr(L, "__dict__", {}):
            setattr(new, attr, getattr(L, attr))
    cf = getattr(L, "_cached_for", None)
    if cf is not None:
        shape, dtype, dev = cf
        new = cache_operator(new, DeviceArray.empty(shape, dtype, dev))
    return new


def adapt(L, device=None, to=None):
    """`Adapt.adapt(to, L)` (src/adapt.jl:12-115): rebuild the operator tree with every stored array re-homed
    by `to` (default: a copy on CUDA device `device`).  Structure, update functions and ScalarOperators are
    kept; caches are dropped (call `cache_operator` again), exactly what replicating the factors onto another
    GPU needs.  `to` may be any callable DeviceArray -> DeviceArray (test/adapt.jl's fake
    backend pattern)."""
    if to is None:
        import torch
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        to = lambda a: DeviceArray(a.t.to(dev, copy=True), a.shape)
    if isinstance(L, MatrixOperator):
        if L.banded:
            return MatrixOperator(BandedMatrix(to(L.A.bands), L.A.b), L.trans, L.update_func)
        if L.sparse:
            return MatrixOperator(SparseMatrixCSR(to(L.A.rowptr), to(L.A.colidx), to(L.A.values), L.A.shape), L.trans, L.update_func)
        return MatrixOperator(to(L.A), L.trans, L.update_func)
    if isinstance(L, _VectorDiagonalOperator):
        return _VectorDiagonalOperator(to(L.diag), L.update_func)
    if isinstance(L, BatchedDiagonalOperator):
        return BatchedDiagonalOperator(to(L.diag), L.update_func)
    if isinstance(L, (IdentityOperator, NullOperator)):
        return type(L)(L.len)
    if isinstance(L, InvertibleOperator):
        if L.lu is not None:                      # re-factorise on the new home (the blob is device-local)
            return factorize(adapt(L.L, to=to))
        return InvertibleOperator(adapt(L.L, to=to), adapt(L.F, to=to))
    if isinstance(L, BlockDiagonalOperator):
        return BlockDiagonalOperator(*[adapt(op, to=to) for op in L.ops])
    if isinstance(L, WOperator):
        lam = L._neg.lam
        return WOperator(lam if isinstance(L.mass_matrix, IdentityOperator) and lam != 1.0 else adapt(L.mass_matrix, to=to),
                         L.gamma, adapt(L.J, to=to))
    if isinstance(L, ScaledOperator):
        return ScaledOperator(L.lam, adapt(L.L, to=to))
    if isinstance(L, AddedOperator):
        return AddedOperator(*[adapt(op, to=to) for op in L.ops])
    if isinstance(L, ComposedOperator):
        return ComposedOperator(*[adapt(op, to=to) for op in L.ops])
    if isinstance(L, TensorProductOperator):
        return TensorProductOperator(*[adapt(op, to=to) for op in L.ops])
    if isinstance(L, TensorSumOperator):
        return TensorSumOperator(*[adapt(op, to=to) for op in L.ops])
    raise TypeError(f"adapt: {type(L).__name__} is outside the B200 hot path")


def iscached(L):
    return L.iscached()


def update_coefficients_(L, u=None, p=None, t=None):
    return L.update_coefficients_(u, p, t)


def mul_(w, L, v, alpha=None, beta=None):
    """`mul!(w, L, v)` / `mul!(w, L, v, α, β)`; returns w."""
    if (alpha is None) != (beta is None):
        raise TypeError("mul!: give both α and β or neither")
    return L.mul_(w, v, alpha, beta)
