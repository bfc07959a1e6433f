"""Large MatrixOperator products (config-4-dense style): TFLOP/s per dtype / precision mode."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
from scimlop_b200 import _lib

def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

for dtype, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
    for (n, k) in ((8192, 1024), (16384, 1024), (4096, 4096)):
        A = sb.DeviceArray(torch.rand(n * n, dtype=tdt, device="cuda"), (n, n))
        v = sb.DeviceArray(torch.rand(n * k, dtype=tdt, device="cuda"), (n, k))
        w = sb.DeviceArray.zeros((n, k), dtype)
        for trans in (False, True):
            L = sb.MatrixOperator(A, trans=trans)
            for prec, name in ((_lib.PREC_DEFAULT, "default"),) + (((_lib.PREC_FP32_EXACT, "fp32_exact"), (_lib.PREC_TF32, "tf32_single_pass_opt_in")) if dtype == np.float32 else ()):
                L.precision = prec
                t = ev_time(lambda: sb.mul_(w, L, v))
                print(json.dumps({"dtype": np.dtype(dtype).name, "n": n, "k": k, "trans": trans, "precision": name,
                                  "ms": round(t, 3), "TFLOPs": round(2.0 * n * n * k / t / 1e9, 2)}), flush=True)
