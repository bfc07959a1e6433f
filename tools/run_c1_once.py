"""configs[0] (100 x 100 (x) 100 x 100 Float64, 10^4 x 100 batch): a few mul! calls (target of an ncu capture)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
g = torch.Generator(device="cuda").manual_seed(0)
A = sb.DeviceArray(torch.rand(100 * 100, dtype=torch.float64, device="cuda", generator=g), (100, 100))
B = sb.DeviceArray(torch.rand(100 * 100, dtype=torch.float64, device="cuda", generator=g), (100, 100))
V = sb.DeviceArray(torch.rand(10 ** 4 * 100, dtype=torch.float64, device="cuda", generator=g), (10 ** 4, 100))
W = sb.DeviceArray.empty((10 ** 4, 100), np.float64)
L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
for _ in range(4):
    sb.mul_(W, L, V)
torch.cuda.synchronize()
print("ok")
