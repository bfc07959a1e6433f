"""Config-2 shape in Float32 (512x512 (x) 512x512, 262144 x 256): streamed tcgen05 path vs exact FFMA."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
from scimlop_b200 import _lib
def ev_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps
rng = np.random.default_rng(0)
n, k = 512, 256
A, B = (np.asfortranarray(rng.random((n, n)).astype(np.float32)) for _ in range(2))
V = sb.DeviceArray(torch.rand(n * n * k, dtype=torch.float32, device="cuda"), (n * n, k))
W = sb.DeviceArray.zeros((n * n, k), np.float32)
L = sb.cache_operator(sb.kron(A, B), V)
for prec, name in ((_lib.PREC_DEFAULT, "default (3xTF32)"), (_lib.PREC_FP32_EXACT, "fp32_exact (FFMA)")):
    L.precision = prec
    t = ev_time(lambda: sb.mul_(W, L, V))
    print(json.dumps({"config": "c2 shape in Float32", "precision": name, "ms": round(t, 3), "TFLOPs": round(4.0 * n ** 3 * k / t / 1e9, 1)}), flush=True)
