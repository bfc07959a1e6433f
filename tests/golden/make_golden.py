"""Generates tests/golden/kron_tree_golden.npz from the pinned CPU oracle (oracle/ref_numpy.py).

The reference holds no golden files for this path; its tests compare against the dense
definition.  These fixtures freeze inputs AND outputs of that definition (np.kron / dense tree x v, evaluated
in Float64) for the reference's own test shapes, so that both the oracle restatements (CPU suite) and the CUDA
path (GPU suite) are checked against stored numbers, independently of any later change to the oracle.

    python tests/golden/make_golden.py      # rewrites the .npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_numpy as R  # noqa: E402

ALPHA, BETA = 0.7, 0.3


def cases():
    rng = np.random.default_rng(20240607)
    f = lambda *s: np.asfortranarray(rng.random(s))
    out = {}
    # TPO shape matrix of test/matrix.jl:510-517 (rectangular and square), K = 19 (test/matrix.jl:12)
    for tag, (m1, n1, m2, n2, m3, n3) in {"rect": (3, 5, 7, 11, 13, 17), "square": (3, 3, 7, 7, 13, 13)}.items():
        A, B, C = f(m1, n1), f(m2, n2), f(m3, n3)
        for nf, mats in ((2, (A, B)), (3, (A, B, C))):
            dense = mats[0]
            for M in mats[1:]:
                dense = np.kron(dense, M)
            v, w0 = f(dense.shape[1], 19), f(dense.shape[0], 19)
            key = f"tpo{nf}_{tag}"
            for i, M in enumerate(mats):
                out[f"{key}_F{i}"] = M
            out[f"{key}_v"], out[f"{key}_w0"] = v, w0
            out[f"{key}_out3"] = dense @ v
            out[f"{key}_out5"] = ALPHA * (dense @ v) + BETA * w0
    # tree of README.md:69-73 / config 4: alpha*(D*M' + 2M + I)*v + beta*w, N = 8, K = 12 (test/basic.jl:16-18)
    N, K = 8, 12
    Mm, d, Dk = f(N, N), f(N), f(N, K)
    v, w0 = f(N, K), f(N, K)
    dense = np.diag(d) @ Mm.T + 2 * Mm + np.eye(N)
    out.update(tree_M=Mm, tree_d=d, tree_Dk=Dk, tree_v=v, tree_w0=w0, tree_out3=dense @ v,
               tree_out5=ALPHA * (dense @ v) + BETA * w0)
    # batched-diagonal tree: (Dk .* (M' v)) + 2 M v + v, one dense operator per column
    Lref = R.AddedOperator(R.ComposedOperator(R.BatchedDiagonalOperator(Dk), R.MatrixOperator(Mm).adjoint()),
                           R.ScaledOperator(2.0, R.MatrixOperator(Mm)), R.IdentityOperator(N))
    bd = R.dense_apply(Lref, v)
    out.update(btree_out3=bd, btree_out5=ALPHA * bd + BETA * w0)
    # Kronecker sum (test/matrix.jl:809-867)
    A, B = f(5, 5), f(7, 7)
    v, w0 = f(35, 9), f(35, 9)
    dense = np.kron(A, np.eye(7)) + np.kron(np.eye(5), B)
    out.update(ksum_A=A, ksum_B=B, ksum_v=v, ksum_w0=w0, ksum_out3=dense @ v, ksum_out5=ALPHA * (dense @ v) + BETA * w0)
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "kron_tree_golden.npz")
    np.savez_compressed(path, **cases())
    print(path, os.path.getsize(path), "bytes")
