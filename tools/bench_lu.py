"""Set-up time of `factorize` (blocked LU on the device + solve structures + explicit inverse + cond_1) and the
throughput of the two ldiv! paths at the config-2 shape (512 (x) 512 Float64 on 256 columns) against mul!.
    python tools/bench_lu.py"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb

def ev(fn, reps=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

g = torch.Generator(device="cuda").manual_seed(0)
for n in (512, 1024, 2048, 4096):
    A = sb.DeviceArray(torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g) - 0.5, (n, n))
    lu = sb.LUFactorization(A)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); lu.refresh(); torch.cuda.synchronize(); t1 = time.perf_counter()
    old = ev(lambda: sb.invert_matrix(A), reps=2, warm=1) if n <= 2048 else None
    print(json.dumps({"factorize": n, "ms": round((t1 - t0) * 1e3, 2), "cond1": lu.cond, "round1_gauss_jordan_ms": old}), flush=True)

n, k = 512, 256
def mk(shift):
    A = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    if shift: A += shift * torch.eye(n, dtype=torch.float64, device="cuda").t().reshape(-1)
    return sb.DeviceArray(A, (n, n))
V = sb.DeviceArray(torch.rand(n * n * k, dtype=torch.float64, device="cuda", generator=g), (n * n, k))
W = sb.DeviceArray.empty((n * n, k), np.float64)
for tag, shift in (("well-conditioned factors (inverse path)", float(n)), ("general factors (substitution path)", 0.0)):
    A, B = mk(shift), mk(shift)
    L = sb.cache_operator(sb.factorize(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B))), V)
    t_mul = ev(lambda: sb.mul_(W, L, V))
    t_div = ev(lambda: sb.ldiv_(W, L, V))
    stable = [op.stable_solve for op in L.ops]
    print(json.dumps({"config2 shape": tag, "stable_solve": stable, "mul_ms": round(t_mul, 3), "ldiv_ms": round(t_div, 3),
                      "ldiv_over_mul": round(t_mul / t_div, 3)}), flush=True)
