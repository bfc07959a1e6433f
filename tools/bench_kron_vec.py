"""TensorProductOperator mul! with vector / few-column input (the ODE-solver shape): 512x512 (x) 512x512 Float64."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb

def ev_time(fn, reps=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

rng = np.random.default_rng(0)
for n in (128, 512, 1024):
    A, B = (np.asfortranarray(rng.random((n, n))) for _ in range(2))
    for k in (1, 2, 4, 8, 16):
        V = sb.DeviceArray(torch.rand(n * n * k, dtype=torch.float64, device="cuda"), (n * n, k))
        W = sb.DeviceArray.zeros((n * n, k), np.float64)
        L = sb.cache_operator(sb.kron(A, B), V)
        t = ev_time(lambda: sb.mul_(W, L, V))
        print(json.dumps({"n": n, "k": k, "ms": round(t, 4), "TFLOPs": round(4.0 * n ** 3 * k / t / 1e9, 2)}), flush=True)
