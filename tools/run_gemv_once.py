"""One Float64 matrix-vector and one 8-column product on a 16384^2 matrix (target of an ncu capture)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
n = 16384
A = sb.DeviceArray(torch.rand(n * n, dtype=torch.float64, device="cuda"), (n, n))
for k in (1, 8):
    v = sb.DeviceArray(torch.rand(n * k, dtype=torch.float64, device="cuda"), (n, k))
    w = sb.DeviceArray.zeros((n, k), np.float64)
    for trans in (False, True):
        L = sb.MatrixOperator(A, trans=trans)
        for _ in range(3):
            sb.mul_(w, L, v)
torch.cuda.synchronize()
print("ok")
