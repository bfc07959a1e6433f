"""MatrixOperator mul! with few columns (k = 1..16): GB/s of the matrix read vs measured HBM."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb

def ev_time(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

for dtype in (np.float64, np.float32):
    es = np.dtype(dtype).itemsize
    for n in (4096, 16384):
        A = sb.DeviceArray(torch.rand(n * n, dtype=torch.float64 if es == 8 else torch.float32, device="cuda"), (n, n))
        for trans in (False, True):
            L = sb.MatrixOperator(A, trans=trans)
            for k in (1, 2, 4, 8, 16):
                v = sb.DeviceArray(torch.rand(n * k, dtype=A.t.dtype, device="cuda"), (n, k))
                w = sb.DeviceArray.zeros((n, k), dtype)
                # rotate over several matrices? one 2 GB / 1 GB matrix is already >> L2 at n = 16384
                t3 = ev_time(lambda: sb.mul_(w, L, v))
                t5 = ev_time(lambda: sb.mul_(w, L, v, 0.7, 0.3))
                print(json.dumps({"dtype": np.dtype(dtype).name, "n": n, "trans": trans, "k": k, "ms_3arg": round(t3, 4),
                                  "ms_5arg": round(t5, 4), "GBps_3arg": round(n * n * es / t3 / 1e6, 1)}), flush=True)
