"""Experiment: can the Kronecker GEMM pair stream its operands straight from / to PINNED HOST memory (TMA loads and
stores over PCIe, no staging copies)?  Calls scimlop_kron_mul (the device entry) on configs[1] with V and / or W being
pinned host buffers (device-addressable under UVA) and times it against the all-device call.

    gpurun -- python tools/exp_zero_copy.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scimlop_b200 as sb
from scimlop_b200 import _lib

h = sb.handle(0)
MO = NO = MI = NI = 512
n = NO * NI
rng = np.random.default_rng(0)
A = sb.DeviceArray.from_numpy(np.asfortranarray(rng.random((MO, NO))))
B = sb.DeviceArray.from_numpy(np.asfortranarray(rng.random((MI, NI))))
ptrs = (C.c_void_p * 2)(A.ptr, B.ptr)
dims = (C.c_int64 * 4)(MO, NO, MI, NI)
lds = (C.c_int64 * 2)(MO, MI)
kinds = (C.c_int32 * 2)(0, 0)


def run(K, v_host, w_host, reps=5):
    Vh = torch.rand(n * K, dtype=torch.float64).pin_memory()
    Wh = torch.zeros(MO * MI * K, dtype=torch.float64).pin_memory()
    Vd = Vh.cuda()
    Wd = torch.zeros(MO * MI * K, dtype=torch.float64, device="cuda")
    nb = C.c_size_t(0)
    h.check(h.lib.scimlop_kron_workspace_bytes(_lib.F64, 2, dims, kinds, K, C.byref(nb)), "ws")
    ws = torch.empty(max(1, nb.value), dtype=torch.uint8, device="cuda")
    V = Vh if v_host else Vd
    W = Wh if w_host else Wd
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        h.check(h.lib.scimlop_kron_mul(h.ptr, _lib.F64, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(V.data_ptr()), n,
                                       C.c_void_p(W.data_ptr()), MO * MI, K, None, None, C.c_void_p(ws.data_ptr()), nb.value, st),
                "scimlop_kron_mul")
    call()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    # correctness against the all-device result
    if v_host or w_host:
        h.check(h.lib.scimlop_kron_mul(h.ptr, _lib.F64, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vd.data_ptr()), n,
                                       C.c_void_p(Wd.data_ptr()), MO * MI, K, None, None, C.c_void_p(ws.data_ptr()), nb.value, st), "ref")
        torch.cuda.synchronize()
        same = bool(torch.equal((Wh if w_host else Wd).cpu(), Wd.cpu())) if w_host else True
    else:
        same = True
    gb = n * K * 8 / 1e9
    print(json.dumps({"K": K, "V": "pinned host" if v_host else "device", "W": "pinned host" if w_host else "device",
                      "ms_min": round(min(ts), 3), "ms_median": round(sorted(ts)[len(ts) // 2], 3),
                      "GBs_if_one_direction": round(gb / (min(ts) * 1e-3), 1), "bit_identical": same}), flush=True)


for K in (32, 256):
    for v_host, w_host in ((False, False), (True, False), (False, True), (True, True)):
        try:
            run(K, v_host, w_host)
        except Exception as e:   # a refusal (tensor map over host memory, ...) is a result too
            print(json.dumps({"K": K, "V_host": v_host, "W_host": w_host, "error": str(e)[:200]}), flush=True)
