"""configs[2] (three 128 x 128 Float32 factors, 2097152 x 512) with the two-pass column-chunk pipeline
(SCIMLOP_OPT_MODE_PIPELINE; SM share of the second pass through SCIMLOP_OPT_PEER_STAGE2_SMS): every shape is compared
bit for bit with the unchunked call and timed.

    gpurun -- python tools/bench_c3_pipeline.py > profiles/r02_c3_pipeline_sweep.jsonl
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scimlop_b200 as sb

h = sb.handle(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device="cuda").manual_seed(0)
F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda", generator=g) - 0.5, (128, 128)) for _ in range(3)]
n = 128 ** 3
V = sb.DeviceArray(torch.rand(n * K, dtype=torch.float32, device="cuda", generator=g), (n, K))
W = sb.DeviceArray.empty((n, K), np.float32)
L = sb.cache_operator(sb.TensorProductOperator(*[sb.MatrixOperator(f) for f in F]), V)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


for five in (False, True):
    call = (lambda: sb.mul_(W, L, V, 0.5, 0.25)) if five else (lambda: sb.mul_(W, L, V))
    h.set_option(sb.OPT_MODE_PIPELINE, 0)
    W.t.fill_(0.125)
    call()
    torch.cuda.synchronize()
    ref = W.t.clone()
    tmin, tmed = timed(call)
    print(json.dumps({"config3": "5-arg" if five else "3-arg", "K": K, "chunks": 0, "pass2_sms": 0, "ms_min": round(tmin, 4),
                      "ms_median": round(tmed, 4)}), flush=True)
    for chunks in (2, 4, 8):
        for sms in (40, 48, 56, 64, 74):
            h.set_option(sb.OPT_MODE_PIPELINE, chunks)
            h.set_option(sb.OPT_PEER_STAGE2_SMS, sms)
            W.t.fill_(0.125)
            call()
            torch.cuda.synchronize()
            same = bool(torch.equal(W.t, ref))
            tmin, tmed = timed(call)
            print(json.dumps({"config3": "5-arg" if five else "3-arg", "K": K, "chunks": chunks, "pass2_sms": sms, "ms_min": round(tmin, 4),
                              "ms_median": round(tmed, 4), "bit_identical_to_unchunked": same}), flush=True)
    h.set_option(sb.OPT_MODE_PIPELINE, 0)
    h.set_option(sb.OPT_PEER_STAGE2_SMS, 0)
