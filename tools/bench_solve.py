"""Solve-side timings: scimlop_invert (setup) and ldiv! of the factorised config-2 operator vs its mul!."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb

def ev_time(fn, reps=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

rng = np.random.default_rng(0)
out = {}
for n in (512, 2048, 4096):
    A = sb.DeviceArray.from_numpy(np.asfortranarray(rng.random((n, n)) + n * np.eye(n)))
    sb.invert_matrix(A); torch.cuda.synchronize()
    t0 = time.perf_counter(); sb.invert_matrix(A); torch.cuda.synchronize()
    out[f"invert_{n}_ms"] = (time.perf_counter() - t0) * 1e3
A, B = (np.asfortranarray(rng.random((512, 512)) + 512 * np.eye(512)) for _ in range(2))
k = 256
V = sb.DeviceArray(torch.rand(512 * 512 * k, dtype=torch.float64, device="cuda"), (512 * 512, k))
W = sb.DeviceArray.empty((512 * 512, k), np.float64)
L = sb.cache_operator(sb.factorize(sb.kron(A, B)), V)
out["c2_mul_ms"] = ev_time(lambda: sb.mul_(W, L, V))
out["c2_ldiv_ms"] = ev_time(lambda: sb.ldiv_(W, L, V))
out["c2_ldiv_inplace_ms"] = ev_time(lambda: sb.ldiv_(L, W))
out["c2_ldiv_tflops"] = 2 * 2 * 512 ** 3 * k / (out["c2_ldiv_ms"] * 1e-3) / 1e12
print(json.dumps(out))
