"""Sparse MatrixOperator (CSR, 65536 x 65536, 16 nnz/row) on 1024 Float64 columns, cached (target of an ncu capture)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
N, K, per_row = 65536, 1024, 16
g = torch.Generator(device="cuda").manual_seed(3)
cols, _ = torch.sort(torch.randint(0, N, (N, per_row), device="cuda", generator=g, dtype=torch.int32), dim=1)
rowptr = torch.arange(0, N * per_row + 1, per_row, device="cuda", dtype=torch.int32)
vals = torch.rand(N * per_row, device="cuda", generator=g, dtype=torch.float64)
A = sb.SparseMatrixCSR(sb.DeviceArray(rowptr, (N + 1,)), sb.DeviceArray(cols.reshape(-1).contiguous(), (N * per_row,)),
                       sb.DeviceArray(vals, (N * per_row,)), (N, N))
V = sb.DeviceArray(torch.rand(N * K, dtype=torch.float64, device="cuda", generator=g), (N, K))
W = sb.DeviceArray.empty((N, K), np.float64)
L = sb.cache_operator(sb.MatrixOperator(A), V)
for _ in range(3):
    sb.mul_(W, L, V)
torch.cuda.synchronize()
print("ok")
