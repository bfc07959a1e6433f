import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
h = sb.handle(0)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(M, N, K, batch, ta, tb, A, B, prec, shared_b=False):
    dA, dB = sb.DeviceArray.from_numpy(A), sb.DeviceArray.from_numpy(B)
    dC = sb.DeviceArray.full((M, N, batch), -7.0, np.float32)
    rc = h.lib.scimlop_gemm_strided_batched(h.ptr, 1, prec, ta, tb, M, N, K, None, C.c_void_p(dA.ptr), A.shape[0], A.shape[0]*A.shape[1],
        C.c_void_p(dB.ptr), B.shape[0], 0 if shared_b else B.shape[0]*B.shape[1], None, C.c_void_p(dC.ptr), M, M*N, batch, st())
    h.check(rc, "gemm"); torch.cuda.synchronize()
    return dC.numpy()
np.set_printoptions(linewidth=200, precision=3, suppress=True)
rng = np.random.default_rng(0)
# path (a): factor A = identity -> C = B
M, N, K = 128, 512, 128
A = np.asfortranarray(np.eye(M, K, dtype=np.float32))[:, :, None]
B = np.asfortranarray(np.arange(K * N, dtype=np.float32).reshape(K, N, order="F") % 97)[:, :, None]
for prec in (2, 0):
    got = run(M, N, K, 1, 0, 0, A, B, prec)[:, :, 0]
    want = A[:, :, 0] @ B[:, :, 0]
    print("path a identity prec", prec, "max abs diff", np.abs(got - want).max(), "unwritten", (got == -7).sum(), "zeros", (got == 0).sum())
    print(got[:6, :6]); print(want[:6, :6])
    print("got row sums", got.sum(1)[:8], "want", want.sum(1)[:8])
# A = single nonzero entries to see the index mapping
for (i, j) in ((0, 0), (1, 0), (0, 1), (5, 9), (40, 70)):
    A2 = np.zeros((M, K, 1), np.float32, order="F"); A2[i, j, 0] = 1
    got = run(M, N, K, 1, 0, 0, A2, B, 2)[:, :, 0]
    nz = np.argwhere(np.abs(got) > 0)
    rows = sorted(set(nz[:, 0].tolist()))
    print(f"A[{i},{j}]=1: nonzero rows {rows[:8]} count {len(nz)}; got[{i},:6]={got[i,:6]} want={B[j,:6,0]}")
