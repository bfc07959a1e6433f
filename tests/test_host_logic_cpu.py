"""Host-side logic of the operator mirror that never reaches the device (CPU only): sizes, flattening of trees and
Kronecker products into the native programs the C ABI takes, adjoint structure, `iscached` / dimension / eltype
errors (the reference's @assert's).  DeviceArray storage is kept on the CPU for these tests; any attempt to launch
would fail in `handle()` -- there is no CPU fallback to fall into."""
import numpy as np
import pytest
import torch

import scimlop_b200 as sb
from scimlop_b200 import _lib, operators as ops
from scimlop_b200.array import DeviceArray


@pytest.fixture(autouse=True)
def cpu_storage(monkeypatch):
    def from_numpy(a, device=None):
        a = np.asarray(a)
        return DeviceArray(torch.from_numpy(np.ascontiguousarray(a.reshape(-1, order="F"))), a.shape)
    monkeypatch.setattr(DeviceArray, "from_numpy", staticmethod(from_numpy))


def M(rng, m, n):
    return sb.MatrixOperator(np.asfortranarray(rng.random((m, n))))


def test_sizes_and_nary_kron_fold():
    rng = np.random.default_rng(0)
    A, B, Cm = M(rng, 3, 5), M(rng, 7, 11), M(rng, 13, 17)
    L = sb.kron(A, B, Cm)
    assert L.size() == (3 * 7 * 13, 5 * 11 * 17)                      # src/tensor.jl:179
    assert isinstance(L.ops[1], sb.TensorProductOperator)              # right fold, src/tensor.jl:121
    leaves, scalars = L._flatten()
    assert [l.size() for l in leaves] == [(3, 5), (7, 11), (13, 17)] and scalars == []
    La = L.adjoint()
    assert La.size() == L.size()[::-1] and all(l.trans for l in La._flatten()[0])


def test_scaled_factors_are_pulled_out_of_a_kron():
    rng = np.random.default_rng(1)
    L = sb.kron(2.0 * M(rng, 4, 4), sb.IdentityOperator(3), -1.5 * M(rng, 2, 2))
    leaves, scalars = L._flatten()
    assert len(leaves) == 3 and sorted(s.convert() for s in scalars) == [-1.5, 2.0]
    assert ops._kron_leaf(leaves[1])[0] == _lib.IDENTITY


def test_tree_flattening_matches_the_reference_rules():
    """Sums flatten (src/basic.jl:619-635), products concatenate, a sum inside a product is not distributed."""
    rng = np.random.default_rng(2)
    N = 6
    Mo, D = M(rng, N, N), sb.DiagonalOperator(rng.random(N))
    tree = D * Mo.adjoint() + 2.0 * Mo + sb.IdentityOperator(N) - 0.5 * (D + Mo)
    terms = ops._terms_with_scalars(tree)
    assert len(terms) == 5
    coeffs = sorted(float(np.prod([s.convert() for s in sc])) if sc else 1.0 for sc, _ in terms)
    assert coeffs == [-0.5, -0.5, 1.0, 1.0, 2.0]
    assert [len(lv) for _, lv in terms] == [2, 1, 1, 1, 1]
    assert ops._terms_with_scalars((D + Mo) * Mo) is None


def test_mul_requires_a_cached_operator_and_matching_shapes():
    rng = np.random.default_rng(3)
    L = sb.kron(M(rng, 3, 3), M(rng, 4, 4))
    v = DeviceArray.from_numpy(np.zeros((12, 2)))
    w = DeviceArray.from_numpy(np.zeros((12, 2)))
    with pytest.raises(AssertionError, match="cache needs to be set up"):   # src/tensor.jl:548
        sb.mul_(w, L, v)
    with pytest.raises(AssertionError, match="DimensionMismatch"):
        sb.mul_(DeviceArray.from_numpy(np.zeros((11, 2))), L, v)
    with pytest.raises(TypeError):                                           # one element type at the ABI
        sb.mul_(DeviceArray.from_numpy(np.zeros((12, 2), dtype=np.float32)), L, v)
    with pytest.raises(TypeError, match="both"):
        sb.mul_(w, L, v, 0.7)
    with pytest.raises(AssertionError, match="alias"):
        L.mul_(v, v)


def test_block_diagonal_and_woperator_structure():
    rng = np.random.default_rng(4)
    L = sb.BlockDiagonalOperator(M(rng, 3, 5), sb.DiagonalOperator(rng.random(4)), sb.IdentityOperator(2))
    assert L.size() == (9, 11) and L.adjoint().size() == (11, 9)             # src/block.jl:66-80
    with pytest.raises(NotImplementedError):
        sb.BlockDiagonalOperator(2.0 * M(rng, 2, 2))
    W = sb.WOperator(M(rng, 5, 5), 0.25, M(rng, 5, 5))
    (c1, l1), (c2, l2) = ops._terms_with_scalars(W._sum)
    assert l1[0] is W.J and l2[0] is W.mass_matrix and c2[0].convert() == -4.0
    W.update_coefficients_(None, None, None, gamma=0.5)
    assert c2[0].convert() == -2.0                                            # γ is read at call time
    with pytest.raises(AssertionError):
        sb.WOperator(M(rng, 4, 4), 1.0, M(rng, 5, 5))


def test_ldiv_needs_factorised_operators():
    rng = np.random.default_rng(5)
    L = sb.kron(M(rng, 3, 3), M(rng, 4, 4))
    assert not sb.has_ldiv(L) and not sb.has_ldiv(M(rng, 3, 3))
    assert sb.has_ldiv(sb.DiagonalOperator(rng.random(3))) and sb.has_ldiv(sb.IdentityOperator(3))
    v = DeviceArray.from_numpy(np.zeros((12, 2)))
    with pytest.raises(TypeError, match="factorize"):
        sb.ldiv_(DeviceArray.from_numpy(np.zeros((12, 2))), sb.MatrixOperator(np.eye(12)), v)
    with pytest.raises(AssertionError, match="square"):
        sb.invert_matrix(np.zeros((3, 4)))


def test_column_shards_cover_every_column_once():
    for k, world in [(256, 8), (10, 4), (3, 8), (1, 2)]:
        seen = []
        for r in range(world):
            s = sb.ColumnShard(k, world, r)
            seen.extend(range(s.first, s.first + s.count))
        assert seen == list(range(k))


def test_sparse_matrix_csr_from_dense_layout():
    """SparseMatrixCSR.from_dense builds 0-based CSR arrays (rows + 1 pointers, column indices sorted within a row)."""
    import importlib
    ops = importlib.import_module("scimlop_b200.operators")

    class FakeDev:      # stands in for DeviceArray on a CPU-only box: keeps the numpy array
        def __init__(self, a):
            self.a = np.asarray(a)
            self.shape = self.a.shape
            self.ndim = self.a.ndim
            self.dtype = self.a.dtype
    orig = ops._as_device
    ops._as_device = lambda a: FakeDev(a)
    try:
        A = np.array([[0.0, 2.0, 0.0], [1.0, 0.0, 3.0], [0.0, 0.0, 0.0], [4.0, 5.0, 6.0]])
        S = ops.SparseMatrixCSR.from_dense(A)
        assert S.shape == (4, 3) and S.nnz == 6
        assert S.rowptr.a.tolist() == [0, 1, 3, 3, 6]
        assert S.colidx.a.tolist() == [1, 0, 2, 0, 1, 2]
        assert S.values.a.tolist() == [2.0, 1.0, 3.0, 4.0, 5.0, 6.0]
        with pytest.raises(ValueError):
            ops.SparseMatrixCSR(np.zeros(3, dtype=np.int32), np.zeros(2, dtype=np.int32), np.zeros(2), (4, 3))
    finally:
        ops._as_device = orig


def test_csr_transpose_builds_the_csc_form():
    """operators._csr_transpose (the one-time conversion behind cached sparse adjoints and adjoint sparse Kronecker
    factors): the CSR arrays of A^T, columns sorted within each row, empty rows / columns kept, duplicates summing the
    same -- checked against the dense transpose on CPU tensors (the function is pure tensor bookkeeping, no launch)."""
    from scimlop_b200 import operators as ops
    rng = np.random.default_rng(12)
    for m, n, density in ((7, 5, 0.4), (1, 9, 0.5), (6, 6, 0.0), (40, 33, 0.1)):
        A = rng.standard_normal((m, n)) * (rng.random((m, n)) < density)
        if m > 3:
            A[2, :] = 0.0                       # an empty row -> an empty column of A^T
        rows, cols = np.nonzero(A)
        rowptr = np.zeros(m + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows, minlength=m), out=rowptr[1:])
        wrap = lambda a: DeviceArray(torch.from_numpy(np.ascontiguousarray(a)), a.shape)
        S = ops.SparseMatrixCSR(wrap(rowptr), wrap(cols.astype(np.int32)), wrap(A[rows, cols]), (m, n))
        T = ops._csr_transpose(S)
        assert T.shape == (n, m) and T.nnz == S.nnz
        rp, ci, va = T.rowptr.t.numpy(), T.colidx.t.numpy(), T.values.t.numpy()
        assert rp.dtype == np.int32 and ci.dtype == np.int32 and rp[0] == 0 and rp[-1] == S.nnz
        back = np.zeros((n, m))
        for r in range(n):
            seg = ci[rp[r]:rp[r + 1]]
            assert np.all(np.diff(seg) > 0), "column indices of a row must be strictly increasing"
            back[r, seg] = va[rp[r]:rp[r + 1]]
        assert np.array_equal(back, A.T)


def test_factorisation_and_option_entry_points_are_declared_and_exported():
    """The round-2 entry points are in the header, the binding table and the library (no GPU needed to resolve them)."""
    from scimlop_b200 import _lib
    lib = _lib.load()
    for name in ("scimlop_factorize", "scimlop_factorize_bytes", "scimlop_kron_ldiv", "scimlop_csr_mul",
                 "scimlop_csr_workspace_bytes", "scimlop_set_option"):
        assert name in _lib.SIGNATURES and hasattr(lib, name)
    import ctypes as C
    fb, sbytes = C.c_size_t(0), C.c_size_t(0)
    assert lib.scimlop_factorize_bytes(_lib.F64, 512, C.byref(fb), C.byref(sbytes)) == _lib.OK
    assert fb.value >= 2 * 512 * 512 * 8 and sbytes.value >= 2 * 512 * 512 * 8
    assert lib.scimlop_factorize_bytes(_lib.F32, 512, C.byref(fb), C.byref(sbytes)) == _lib.ERR_UNSUPPORTED
    nb = C.c_size_t(1)
    assert lib.scimlop_csr_workspace_bytes(_lib.F64, 0, 100, 200, 16, C.byref(nb)) == _lib.OK and nb.value == 200 * 16 * 8
    assert lib.scimlop_csr_workspace_bytes(_lib.F64, 1, 100, 200, 16, C.byref(nb)) == _lib.OK and nb.value == 0
    assert lib.scimlop_set_option(None, 1, 0) == _lib.ERR_BAD_HANDLE


def test_mixed_eltype_products_report_the_promoted_eltype():
    rng = np.random.default_rng(8)
    A32 = DeviceArray.from_numpy(np.asfortranarray(rng.random((3, 3)).astype(np.float32)))
    B64 = DeviceArray.from_numpy(np.asfortranarray(rng.random((4, 4))))
    L = sb.TensorProductOperator(sb.MatrixOperator(A32), sb.MatrixOperator(B64))
    assert L.dtype == np.float64
    assert (sb.MatrixOperator(A32) + sb.MatrixOperator(A32)).dtype == np.float32
