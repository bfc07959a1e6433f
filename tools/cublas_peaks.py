"""cuBLAS cross-check of the FP64 / FP32 GEMM ceilings: torch.matmul f64 and f32
(TF32 off and on) at 4096^3 and 8192^3, burst (best of N) and sustained (~2 s loop). Library calls are
used ONLY as a roofline sanity ceiling, never on the product path."""
import json, sys, torch

def bench(dtype, n, allow_tf32=False, secs=2.0):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.rand(n, n, device="cuda", dtype=dtype); b = torch.rand(n, n, device="cuda", dtype=dtype)
    c = torch.empty_like(a)
    for _ in range(3): torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(3, int(secs * 1000 / best))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    sus = e0.elapsed_time(e1) / reps
    fl = 2.0 * n ** 3
    return {"dtype": str(dtype), "n": n, "tf32": allow_tf32, "tflops_burst": fl / best / 1e9,
            "tflops_sustained": fl / sus / 1e9}

if __name__ == "__main__":
    out = []
    for n in (4096, 8192):
        out.append(bench(torch.float64, n))
        out.append(bench(torch.float32, n, False))
        out.append(bench(torch.float32, n, True))
    print(json.dumps({"cublas": out}))
