import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
rng = np.random.default_rng(0)
for (mo, mi, k) in ((40, 300, 3), (16, 64, 1), (600, 520, 2)):
    A = np.asfortranarray(rng.random((mo, 3))); B = np.asfortranarray(rng.random((mi, 3)))
    A[0, 0] = A[-1, 2] = B[0, 0] = B[-1, 2] = 0
    v = np.asfortranarray(rng.random((mo * mi, k)))
    L = sb.kronsum(sb.MatrixOperator(sb.BandedMatrix(A, 1)), sb.MatrixOperator(sb.BandedMatrix(B, 1)))
    dv = sb.DeviceArray.from_numpy(v)
    L = sb.cache_operator(L, dv)
    w = sb.DeviceArray.full((mo * mi, k), float("nan"), np.float64)
    try:
        sb.mul_(w, L, dv); torch.cuda.synchronize()
        got = w.numpy()
        want = np.empty_like(v)
        for j in range(k):
            Vj = v[:, j].reshape((mi, mo), order="F")
            Wj = np.zeros_like(Vj)
            for d in (-1, 0, 1):
                for i in range(mi):
                    if 0 <= i + d < mi: Wj[i, :] += B[i, d + 1] * Vj[i + d, :]
                for o in range(mo):
                    if 0 <= o + d < mo: Wj[:, o] += A[o, d + 1] * Vj[:, o + d]
            want[:, j] = Wj.reshape(-1, order="F")
        print(mo, mi, k, "max abs err", np.nanmax(np.abs(got - want)), "nan count", np.isnan(got).sum())
    except Exception as e:
        print(mo, mi, k, "FAILED", str(e)[:200]); break
