import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import scimlop_b200 as sb
h = sb.handle(0)
g = torch.Generator(device="cuda").manual_seed(0)
A = sb.DeviceArray(torch.rand(512*512, dtype=torch.float64, device="cuda", generator=g), (512, 512))
B = sb.DeviceArray(torch.rand(512*512, dtype=torch.float64, device="cuda", generator=g), (512, 512))
res = {}
for k in (256, 128, 64, 32):
    V = sb.DeviceArray(torch.rand(512*512*k, dtype=torch.float64, device="cuda", generator=g), (512*512, k))
    W = sb.DeviceArray.empty((512*512, k), np.float64)
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)
    for opt in (0, 1):
        h.set_option(1, opt)
        for _ in range(3): sb.mul_(W, L, V)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): sb.mul_(W, L, V)
        e1.record(); e1.synchronize()
        res[f"k={k} tail_split={opt}"] = e0.elapsed_time(e1) / 20
base = res["k=256 tail_split=0"]
for kk, v in res.items():
    k = int(kk.split()[0][2:])
    print(json.dumps({"case": kk, "ms": round(v, 4), "strong_efficiency_vs_k256": round(base / v / (256 / k), 3)}))
