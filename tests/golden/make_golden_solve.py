"""Generates tests/golden/solve_blocks_golden.npz: frozen inputs and outputs of the DENSE definitions
(`kron(A,B[,C]) \\ v`, `blockdiag(...) * v`, `(J - M/γ) * v`, Float64) for the widened rows of SURVEY.md §8(f):
`ldiv!` of a factorised TensorProductOperator (test/matrix.jl:646-671, :801-806), BlockDiagonalOperator (src/block.jl)
and WOperator (src/woperator.jl:447-481).  Checked by tests/test_golden.py against the oracle restatements (CPU)
and against the CUDA path (GPU).

    python tests/golden/make_golden_solve.py
"""
import os

import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
ALPHA, BETA = 0.7, 0.3


def cases():
    rng = np.random.default_rng(20241020)
    f = lambda *s: np.asfortranarray(rng.random(s))
    out = {}
    # square factors of the reference's shape matrix, shifted to be well conditioned (the solve's forward error is κ·ε)
    n1, n2, n3 = 3, 7, 13
    A, B, C = f(n1, n1) + n1 * np.eye(n1), f(n2, n2) + n2 * np.eye(n2), f(n3, n3) + n3 * np.eye(n3)
    out.update(ldiv_A=A, ldiv_B=B, ldiv_C=C)
    for nf, mats in ((2, (A, B)), (3, (A, B, C))):
        dense = mats[0]
        for M in mats[1:]:
            dense = np.kron(dense, M)
        v, vec = f(dense.shape[0], 19), rng.random(dense.shape[0])
        out[f"ldiv{nf}_v"], out[f"ldiv{nf}_vec"] = v, vec
        out[f"ldiv{nf}_out"], out[f"ldiv{nf}_outvec"] = np.linalg.solve(dense, v), np.linalg.solve(dense, vec)
    # block diagonal: dense 3x5, diagonal 4, dense 6x2, identity 3 (ragged, rectangular blocks)
    b0, d1, b2 = f(3, 5), rng.random(4) + 0.5, f(6, 2)
    dense = sla.block_diag(b0, np.diag(d1), b2, np.eye(3))
    v, w0 = f(dense.shape[1], 12), f(dense.shape[0], 12)
    out.update(blk_b0=b0, blk_d1=d1, blk_b2=b2, blk_v=v, blk_w0=w0, blk_out3=dense @ v,
               blk_out5=ALPHA * (dense @ v) + BETA * w0)
    # W = J - M/γ on one state vector
    n = 16
    J, M, gamma = f(n, n) - 0.5, f(n, n), 0.37
    b = rng.random(n)
    out.update(w_J=J, w_M=M, w_gamma=np.float64(gamma), w_b=b, w_out=(J - M / gamma) @ b,
               w_out_uniform=(J - 2.0 * np.eye(n) / gamma) @ b)
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "solve_blocks_golden.npz")
    np.savez_compressed(path, **cases())
    print(path, os.path.getsize(path), "bytes")
