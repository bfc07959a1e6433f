"""Golden fixtures (tests/golden/kron_tree_golden.npz, made by tests/golden/make_golden.py): the CPU oracle's
restated reference algorithms on the CPU suite, and the CUDA path through the C ABI on the GPU suite, both
against the same stored numbers."""
import os

import numpy as np
import pytest

from oracle import ref_numpy as R

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kron_tree_golden.npz"))
ALPHA, BETA = 0.7, 0.3
F = np.asfortranarray


@pytest.mark.parametrize("key", ["tpo2_rect", "tpo3_rect", "tpo2_square", "tpo3_square"])
@pytest.mark.parametrize("lv", [False, True])
def test_oracle_tpo_against_golden(key, lv):
    nf = int(key[3])
    mats = [G[f"{key}_F{i}"] for i in range(nf)]
    v, w0 = F(G[f"{key}_v"]), F(G[f"{key}_w0"])
    L = R.TensorProductOperator(*mats, lv=lv).cache_operator(v)
    w = np.zeros_like(w0, order="F")
    assert R.rel_err(L.mul3(w, v), G[f"{key}_out3"]) < 1e-13
    w = w0.copy(order="F")
    assert R.rel_err(L.mul5(w, v, ALPHA, BETA), G[f"{key}_out5"]) < 1e-13
    assert R.rel_err(R.kron_apply(mats, v, ALPHA, BETA, w0), G[f"{key}_out5"]) < 1e-13


def test_oracle_trees_against_golden(oracle_c):
    Mm, d, Dk, v, w0 = (F(G[k]) for k in ("tree_M", "tree_d", "tree_Dk", "tree_v", "tree_w0"))
    N = Mm.shape[0]
    L = R.AddedOperator(R.ComposedOperator(R.DiagonalOperator(d), R.MatrixOperator(Mm).adjoint()),
                        R.ScaledOperator(2.0, R.MatrixOperator(Mm)), R.IdentityOperator(N)).cache_operator(v)
    assert R.rel_err(L.mul3(np.zeros_like(w0, order="F"), v), G["tree_out3"]) < 1e-13
    assert R.rel_err(L.mul5(w0.copy(order="F"), v, ALPHA, BETA), G["tree_out5"]) < 1e-13
    Lb = R.AddedOperator(R.ComposedOperator(R.BatchedDiagonalOperator(Dk), R.MatrixOperator(Mm).adjoint()),
                         R.ScaledOperator(2.0, R.MatrixOperator(Mm)), R.IdentityOperator(N)).cache_operator(v)
    assert R.rel_err(Lb.mul5(w0.copy(order="F"), v, ALPHA, BETA), G["btree_out5"]) < 1e-13
    S = R.TensorSumOperator(F(G["ksum_A"]), F(G["ksum_B"])).cache_operator(F(G["ksum_v"]))
    assert R.rel_err(S.mul5(F(G["ksum_w0"]).copy(order="F"), F(G["ksum_v"]), ALPHA, BETA), G["ksum_out5"]) < 1e-13
    # C restatement of the 2-factor algorithms
    for lv in (True, False):
        got = oracle_c.tpo2_mul(F(G["tpo2_rect_w0"]).copy(order="F"), F(G["tpo2_rect_F0"]), F(G["tpo2_rect_F1"]),
                                F(G["tpo2_rect_v"]), ALPHA, BETA, lv=lv)
        assert R.rel_err(got, G["tpo2_rect_out5"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_cuda_path_against_golden(dtype, tol):
    import scimlop_b200 as sb
    dev = lambda a: sb.DeviceArray.from_numpy(F(np.asarray(a, dtype=dtype)))
    for key in ("tpo2_rect", "tpo3_rect", "tpo2_square", "tpo3_square"):
        nf = int(key[3])
        mats = [F(G[f"{key}_F{i}"].astype(dtype)) for i in range(nf)]
        v = dev(G[f"{key}_v"])
        L = sb.cache_operator(sb.TensorProductOperator(*mats), v)
        w = sb.DeviceArray.full(G[f"{key}_w0"].shape, float("nan"), dtype)
        assert R.rel_err(sb.mul_(w, L, v).numpy(), G[f"{key}_out3"]) <= tol, key
        w = dev(G[f"{key}_w0"])
        assert R.rel_err(sb.mul_(w, L, v, ALPHA, BETA).numpy(), G[f"{key}_out5"]) <= tol, key
    Mm, d, Dk = (F(G[k].astype(dtype)) for k in ("tree_M", "tree_d", "tree_Dk"))
    v = dev(G["tree_v"])
    N = Mm.shape[0]
    Mop = sb.MatrixOperator(Mm)
    L = sb.cache_operator(sb.DiagonalOperator(d) * Mop.adjoint() + 2.0 * Mop + sb.IdentityOperator(N), v)
    assert R.rel_err(sb.mul_(dev(G["tree_w0"]), L, v, ALPHA, BETA).numpy(), G["tree_out5"]) <= tol
    Lb = sb.cache_operator(sb.DiagonalOperator(Dk) * Mop.adjoint() + 2.0 * Mop + sb.IdentityOperator(N), v)
    assert R.rel_err(sb.mul_(dev(G["tree_w0"]), Lb, v, ALPHA, BETA).numpy(), G["btree_out5"]) <= tol
    S = sb.cache_operator(sb.kronsum(F(G["ksum_A"].astype(dtype)), F(G["ksum_B"].astype(dtype))), dev(G["ksum_v"]))
    assert R.rel_err(sb.mul_(dev(G["ksum_w0"]), S, dev(G["ksum_v"]), ALPHA, BETA).numpy(), G["ksum_out5"]) <= tol


# ---- ldiv!, BlockDiagonalOperator, WOperator (tests/golden/make_golden_solve.py) ----
GS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_blocks_golden.npz"))


def test_oracle_solve_and_blocks_against_golden():
    A, B, Cm = F(GS["ldiv_A"]), F(GS["ldiv_B"]), F(GS["ldiv_C"])
    assert R.rel_err(R.tpo_ldiv(A, B, F(GS["ldiv2_v"])), GS["ldiv2_out"]) < 1e-12
    assert R.rel_err(R.tpo_ldiv(A, B, GS["ldiv2_vec"]), GS["ldiv2_outvec"]) < 1e-12
    assert R.rel_err(R.tpo_ldiv(A, np.kron(B, Cm), F(GS["ldiv3_v"])), GS["ldiv3_out"]) < 1e-12
    assert R.rel_err(R.kron_solve([A, B, Cm], GS["ldiv3_vec"]), GS["ldiv3_outvec"]) < 1e-12
    blocks = [GS["blk_b0"], np.diag(GS["blk_d1"]), GS["blk_b2"], np.eye(3)]
    assert R.rel_err(R.block_diag_mul(np.full_like(GS["blk_w0"], np.nan), blocks, GS["blk_v"]), GS["blk_out3"]) < 1e-13
    assert R.rel_err(R.block_diag_mul(GS["blk_w0"].copy(), blocks, GS["blk_v"], ALPHA, BETA), GS["blk_out5"]) < 1e-13
    g = float(GS["w_gamma"])
    assert R.rel_err(R.woperator_mul(np.empty(16), GS["w_M"], g, GS["w_J"], GS["w_b"]), GS["w_out"]) < 1e-13
    assert R.rel_err(R.woperator_mul(np.empty(16), 2.0, g, GS["w_J"], GS["w_b"]), GS["w_out_uniform"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_cuda_solve_and_blocks_against_golden(dtype, tol):
    import scimlop_b200 as sb
    dev = lambda a: sb.DeviceArray.from_numpy(F(np.asarray(a, dtype=dtype)))
    A, B, Cm = (F(GS[k].astype(dtype)) for k in ("ldiv_A", "ldiv_B", "ldiv_C"))
    for nf, mats in ((2, (A, B)), (3, (A, B, Cm))):
        for vk, ok in ((f"ldiv{nf}_v", f"ldiv{nf}_out"), (f"ldiv{nf}_vec", f"ldiv{nf}_outvec")):
            v = dev(GS[vk])
            L = sb.cache_operator(sb.factorize(sb.kron(*mats)), v)
            w = sb.DeviceArray.full(GS[vk].shape, float("nan"), dtype)
            assert R.rel_err(sb.ldiv_(w, L, v).numpy(), GS[ok]) <= tol, vk
            u = dev(GS[vk])
            assert R.rel_err(sb.ldiv_(L, u).numpy(), GS[ok]) <= tol, vk + " in place"
    v = dev(GS["blk_v"])
    L = sb.cache_operator(sb.BlockDiagonalOperator(F(GS["blk_b0"].astype(dtype)), sb.DiagonalOperator(GS["blk_d1"].astype(dtype)),
                                                   F(GS["blk_b2"].astype(dtype)), sb.IdentityOperator(3)), v)
    w = sb.DeviceArray.full(GS["blk_w0"].shape, float("nan"), dtype)
    assert R.rel_err(sb.mul_(w, L, v).numpy(), GS["blk_out3"]) <= tol
    assert R.rel_err(sb.mul_(dev(GS["blk_w0"]), L, v, ALPHA, BETA).numpy(), GS["blk_out5"]) <= tol
    b = dev(GS["w_b"])
    g = float(GS["w_gamma"])
    y = sb.DeviceArray.full((16,), float("nan"), dtype)
    W = sb.cache_operator(sb.WOperator(F(GS["w_M"].astype(dtype)), g, F(GS["w_J"].astype(dtype))), b)
    assert R.rel_err(sb.mul_(y, W, b).numpy(), GS["w_out"]) <= 4 * tol
    Wu = sb.cache_operator(sb.WOperator(2.0, g, F(GS["w_J"].astype(dtype))), b)
    assert R.rel_err(sb.mul_(y, Wu, b).numpy(), GS["w_out_uniform"]) <= 4 * tol
