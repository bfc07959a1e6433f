import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import scimlop_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device="cuda").manual_seed(0)
A = sb.DeviceArray(torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g) - 0.5, (n, n))
lu = sb.LUFactorization(A)
torch.cuda.synchronize()
print("ok", lu.cond)
