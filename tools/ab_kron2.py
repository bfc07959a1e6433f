import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from importlib import import_module
variant = sys.argv[1]
import scimlop_b200 as sb
from scimlop_b200 import _lib
_lib.LIB_PATH = f"/root/repo/scimloperators.jl_b200/libscimlop_b200_{variant}.so"
h = sb.handle(0)
g = torch.Generator(device="cuda").manual_seed(0)
k = 32768
F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda", generator=g), (128, 128)) for _ in range(2)]
V = sb.DeviceArray(torch.rand(128 * 128 * k, dtype=torch.float32, device="cuda", generator=g), (128 * 128, k))
W = sb.DeviceArray.empty((128 * 128, k), np.float32)
L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)
for _ in range(5): sb.mul_(W, L, V)
torch.cuda.synchronize()
ts = []
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): sb.mul_(W, L, V)
    e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
print(variant, [round(t, 4) for t in ts])
