"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/scimlop_b200.h declares; entry points that need no GPU behave; and there is no CPU fallback (asking
for a handle without an sm_100 device fails loudly)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import scimlop_b200 as sb
    return sb.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "scimlop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scimlop_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_what_the_binding_binds(lib):
    from scimlop_b200 import _lib
    names = header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and include/scimlop_b200.h disagree"


def header_constants():
    """Every `SCIMLOP_X = n` enumerator and `#define SCIMLOP_X n` of the header."""
    src = open(os.path.join(ROOT, "include", "scimlop_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"\b(SCIMLOP_[A-Z0-9_]+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", src)}
    out.update({m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+(SCIMLOP_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|\d+)\b", src)})
    return out


def test_header_constants_match_the_binding():
    """The numeric codes a binding hard-codes (dtypes, precisions, factor kinds, options, SCIMLOP_TRANS, the sparse
    factor descriptor) are the header's: a renumbered enum would otherwise pass every symbol check and corrupt calls."""
    from scimlop_b200 import _lib
    hc = header_constants()
    pairs = {"SCIMLOP_F64": _lib.F64, "SCIMLOP_F32": _lib.F32, "SCIMLOP_C64": _lib.C64, "SCIMLOP_C128": _lib.C128,
             "SCIMLOP_PREC_DEFAULT": _lib.PREC_DEFAULT, "SCIMLOP_PREC_TF32X3": _lib.PREC_TF32X3, "SCIMLOP_PREC_TF32": _lib.PREC_TF32,
             "SCIMLOP_PREC_FP32_EXACT": _lib.PREC_FP32_EXACT, "SCIMLOP_PREC_BF16": _lib.PREC_BF16,
             "SCIMLOP_DENSE": _lib.DENSE, "SCIMLOP_IDENTITY": _lib.IDENTITY, "SCIMLOP_DIAG": _lib.DIAG, "SCIMLOP_BDIAG": _lib.BDIAG,
             "SCIMLOP_NULL": _lib.NULL, "SCIMLOP_BANDED": _lib.BANDED, "SCIMLOP_CSR": _lib.CSR, "SCIMLOP_TRANS": _lib.TRANS,
             "SCIMLOP_OPT_TAIL_SPLIT": _lib.OPT_TAIL_SPLIT, "SCIMLOP_OPT_PEER_PIPELINE": _lib.OPT_PEER_PIPELINE,
             "SCIMLOP_OPT_PEER_CHUNKS": _lib.OPT_PEER_CHUNKS, "SCIMLOP_OPT_PEER_STAGE2_SMS": _lib.OPT_PEER_STAGE2_SMS,
             "SCIMLOP_OPT_MODE_PIPELINE": _lib.OPT_MODE_PIPELINE,
             "SCIMLOP_ERR_SINGULAR": _lib.ERR_SINGULAR, "SCIMLOP_ERR_UNSUPPORTED": _lib.ERR_UNSUPPORTED}
    missing = [k for k in pairs if k not in hc]
    assert not missing, f"not found in the header: {missing}"
    wrong = {k: (hc[k], v) for k, v in pairs.items() if hc[k] != v}
    assert not wrong, f"header vs binding: {wrong}"
    # scimlop_csr_factor: three pointers + two int32 (include/scimlop_b200.h)
    assert C.sizeof(_lib.CsrFactor) == 3 * C.sizeof(C.c_void_p) + 8
    assert [f[0] for f in _lib.CsrFactor._fields_] == ["rowptr", "colidx", "values", "index_base", "reserved"]


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), f"libscimlop_b200.so does not export {name}"


def test_version_and_status_strings(lib):
    assert lib.scimlop_version() == 100
    seen = {lib.scimlop_status_string(s).decode() for s in range(0, 9)}
    assert len(seen) == 9 and "ok" in seen


def test_shard_columns_partitions_the_batch(lib):
    import scimlop_b200 as sb
    for k in (0, 1, 7, 256, 1000):
        for n in (1, 2, 3, 8):
            spans = [sb.shard_columns(k, n, r) for r in range(n)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == k
            for (f0, c0), (f1, _) in zip(spans[:-1], spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    first, count = C.c_int64(), C.c_int64()
    assert lib.scimlop_shard_columns(8, 2, 2, C.byref(first), C.byref(count)) == 2  # bad rank -> BAD_SHAPE
    assert lib.scimlop_shard_columns(8, 2, 0, None, C.byref(count)) == 1            # NULL -> NULL_POINTER


def test_workspace_queries_need_no_gpu(lib):
    from scimlop_b200 import _lib
    nbytes = C.c_size_t(123)
    # 512x512 (x) 512x512, k = 256 -> one intermediate of 512*512*256 doubles (the reference keeps 4: c1..c4)
    dims = (C.c_int64 * 4)(512, 512, 512, 512)
    kinds = (C.c_int32 * 2)(_lib.DENSE, _lib.DENSE)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dims, kinds, 256, C.byref(nbytes)) == 0
    assert nbytes.value == 512 * 512 * 256 * 8
    # identity outer: no scratch at all (src/tensor.jl:502-506 allocates `nothing` there too)
    kinds = (C.c_int32 * 2)(_lib.IDENTITY, _lib.DENSE)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dims, kinds, 256, C.byref(nbytes)) == 0
    assert nbytes.value == 0
    # three dense factors: two ping-pong buffers
    dims3 = (C.c_int64 * 6)(128, 128, 128, 128, 128, 128)
    kinds3 = (C.c_int32 * 3)(0, 0, 0)
    assert lib.scimlop_kron_workspace_bytes(_lib.F32, 3, dims3, kinds3, 4, C.byref(nbytes)) == 0
    assert nbytes.value == 2 * 128 ** 3 * 4 * 4
    # rectangular: intermediates are sized by the partially transformed tensor
    dimsr = (C.c_int64 * 4)(3, 5, 7, 11)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dimsr, (C.c_int32 * 2)(0, 0), 19, C.byref(nbytes)) == 0
    assert nbytes.value >= 7 * 5 * 19 * 8
    # errors
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 0, dims, kinds, 1, C.byref(nbytes)) == _lib.ERR_BAD_SHAPE
    assert lib.scimlop_kron_workspace_bytes(5, 2, dims, kinds, 1, C.byref(nbytes)) == _lib.ERR_UNSUPPORTED
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dims, kinds, 1, None) == _lib.ERR_NULL_POINTER
    bad = (C.c_int32 * 2)(_lib.BDIAG, _lib.DENSE)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dims, bad, 1, C.byref(nbytes)) == _lib.ERR_UNSUPPORTED
    # a sparse innermost factor (rectangular 40 x 33) next to a dense 7 x 9 one, k = 5: the intermediate (40 * 9 * 5) plus
    # the row-contiguous staging copy of the input (33 * 9 * 5); a sparse OUTER factor needs no staging
    dimss = (C.c_int64 * 4)(7, 9, 40, 33)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dimss, (C.c_int32 * 2)(_lib.DENSE, _lib.CSR), 5, C.byref(nbytes)) == 0
    up = lambda b: (b + 255) // 256 * 256
    assert nbytes.value == up(40 * 9 * 5 * 8) + up(33 * 9 * 5 * 8)
    dimso = (C.c_int64 * 4)(40, 33, 7, 9)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dimso, (C.c_int32 * 2)(_lib.CSR, _lib.DENSE), 5, C.byref(nbytes)) == 0
    assert nbytes.value == up(7 * 33 * 5 * 8)
    assert lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dimss, (C.c_int32 * 2)(_lib.DENSE, _lib.CSR | _lib.TRANS), 5,
                                            C.byref(nbytes)) == _lib.ERR_UNSUPPORTED


def test_tree_workspace_query(lib):
    from scimlop_b200 import _lib
    N, k = 64, 8

    def term(kinds):
        arr = (_lib.Factor * len(kinds))()
        for f, kd in zip(arr, kinds):
            f.kind, f.rows, f.cols, f.ld = kd, N, N, N
        return arr

    keep = [term([_lib.DIAG, _lib.DENSE]), term([_lib.DENSE]), term([_lib.IDENTITY])]
    terms = (_lib.Term * 3)()
    for t, arr in zip(terms, keep):
        t.scale, t.nfactors, t.factors = 1.0, len(arr), arr
    nbytes = C.c_size_t(1)
    assert lib.scimlop_tree_workspace_bytes(_lib.F64, terms, 3, k, C.byref(nbytes)) == 0
    assert nbytes.value == 0          # D*M' + 2M + I needs no scratch: everything rides in epilogues
    keep2 = [term([_lib.DENSE, _lib.DIAG, _lib.DENSE, _lib.BDIAG])]
    terms2 = (_lib.Term * 1)()
    terms2[0].scale, terms2[0].nfactors, terms2[0].factors = 1.0, 4, keep2[0]
    assert lib.scimlop_tree_workspace_bytes(_lib.F64, terms2, 1, k, C.byref(nbytes)) == 0
    assert nbytes.value == 2 * N * k * 8
    # mismatched summand sizes are rejected like the reference's AddedOperator constructor
    keep[1][0].rows = N + 1
    assert lib.scimlop_tree_workspace_bytes(_lib.F64, terms, 3, k, C.byref(nbytes)) == _lib.ERR_BAD_SHAPE


def test_no_cpu_fallback(lib):
    """Without an sm_100 GPU the product refuses to run instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from scimlop_b200 import _lib
    h = C.c_void_p()
    assert lib.scimlop_create(C.byref(h), 0) == _lib.ERR_NO_DEVICE
    assert not h.value
    with pytest.raises(_lib.ScimlopError):
        _lib.Handle(0)
    # a NULL handle is rejected by every compute entry point
    assert lib.scimlop_axpby(None, 0, 4, 1, None, 4, None, 4, None, None, None) == _lib.ERR_BAD_HANDLE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scimloperators.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu fallback", ""), f"{f} mentions the oracle"


def test_host_mirror_algebra_structure():
    """Operator algebra builds the same lazy structures as the reference (no GPU needed: sizes only)."""
    import scimlop_b200 as sb
    I4, I3, N5 = sb.IdentityOperator(4), sb.IdentityOperator(3), sb.NullOperator(4)
    S = sb.AddedOperator(I4, sb.AddedOperator(N5, I4))
    assert len(S.ops) == 3                                    # nested sums flatten, src/basic.jl:619-635
    T = sb.TensorProductOperator(I3, I4, I3)                  # right fold, src/tensor.jl:121
    assert isinstance(T.ops[1], sb.TensorProductOperator) and T.size() == (36, 36)
    assert (2.0 * I4).size() == (4, 4) and isinstance(I4 * I4, sb.ComposedOperator)
    with pytest.raises(AssertionError):
        sb.AddedOperator(I4, I3)
    with pytest.raises(AssertionError):
        sb.ComposedOperator(I4, I3)
    with pytest.raises(TypeError):
        sb.mul_(None, I4, None, 1.0)                          # α without β
