"""1-GPU probe: kron_mul_peers with 'peer' targets that are local buffers (isolates the kernel from NVLink)."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
from scimlop_b200.operators import _DT, _scalar
h = sb.handle(0)
kper = int(os.environ.get("KPER", 32)); n = 512 * 512
rng = np.random.default_rng(0)
A, B = np.asfortranarray(rng.random((512, 512))), np.asfortranarray(rng.random((512, 512)))
V = sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda"), (n, kper))
L = sb.cache_operator(sb.kron(A, B), V)
r1 = sb.DeviceArray.zeros((n, kper), np.float64)
sb.mul_(r1, L, V); torch.cuda.synchronize()
nat = L._native
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = {}
for npeer in (1, 3, 7):
    W = sb.DeviceArray.zeros((n, kper), np.float64)
    P = [sb.DeviceArray.zeros((n, kper), np.float64) for _ in range(npeer)]
    arr = (C.c_void_p * npeer)(*[p.ptr for p in P])
    bad_local = bad_peer = 0
    for rep in range(20):
        W.t.zero_()
        for p in P: p.t.zero_()
        h.check(h.lib.scimlop_kron_mul_peers(h.ptr, _DT[V.dtype], L.precision, nat["nf"], nat["ptrs"], nat["dims"],
                nat["lds"], nat["kinds"], C.c_void_p(V.ptr), kper, C.c_void_p(W.ptr), arr, npeer,
                _scalar(V.dtype, None), C.c_void_p(nat["ws"].data_ptr()), nat["ws_bytes"], st), "peers")
        torch.cuda.synchronize()
        dl = float((W.t - r1.t).abs().max()); dp = max(float((p.t - r1.t).abs().max()) for p in P)
        bad_local += dl > 0; bad_peer += dp > 0
        if dl > 0 and "detail" not in out:
            d = (W.t - r1.t).reshape(kper, 512, 512)  # [k, mo, l]
            idx = torch.nonzero(d != 0)
            out["detail"] = {"npeer": npeer, "rep": rep, "count": int(idx.shape[0]),
                             "k": sorted(set(idx[:, 0].tolist())), "mo": sorted(set(idx[:, 1].tolist())),
                             "l": sorted(set(idx[:, 2].tolist())),
                             "vals": d[idx[:8, 0], idx[:8, 1], idx[:8, 2]].tolist()}
    out[f"npeer{npeer}"] = {"bad_local": int(bad_local), "bad_peer": int(bad_peer)}
print(json.dumps(out))
