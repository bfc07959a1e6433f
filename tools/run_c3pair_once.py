"""One fused two-mode Float32 Kronecker product, 128 x 128 (x) 128 x 128 on a 16384 x k batch (target of an ncu capture).
    python tools/run_c3pair_once.py [k] [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
k = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(0)
F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda", generator=g), (128, 128)) for _ in range(2)]
V = sb.DeviceArray(torch.rand(128 * 128 * k, dtype=torch.float32, device="cuda", generator=g), (128 * 128, k))
W = sb.DeviceArray.empty((128 * 128, k), np.float32)
L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
for _ in range(reps):
    sb.mul_(W, L, V)
torch.cuda.synchronize()
print("ok")
