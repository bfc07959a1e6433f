"""One large Float32 MatrixOperator product (8192^2 x 1024 columns) on the streamed tcgen05 kernel (ncu target)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
n, k = 8192, 1024
A = sb.DeviceArray(torch.rand(n * n, dtype=torch.float32, device="cuda"), (n, n))
v = sb.DeviceArray(torch.rand(n * k, dtype=torch.float32, device="cuda"), (n, k))
w = sb.DeviceArray.zeros((n, k), np.float32)
L = sb.MatrixOperator(A)
for _ in range(4):
    sb.mul_(w, L, v)
torch.cuda.synchronize()
print("ok")
