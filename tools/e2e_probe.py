"""e2e chunk-size sweep for scimlop_kron_mul_host on config 2 + raw PCIe copy bandwidths (debug aid)."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
h = sb.handle(0)
n, k = 512 * 512, 256
A = np.asfortranarray(np.random.default_rng(0).random((512, 512))); B = np.asfortranarray(np.random.default_rng(1).random((512, 512)))
Vh = torch.rand(n * k, dtype=torch.float64).pin_memory(); Wh = torch.zeros(n * k, dtype=torch.float64).pin_memory()
ptrs = (C.c_void_p * 2)(A.ctypes.data, B.ctypes.data); dims = (C.c_int64 * 4)(512, 512, 512, 512)
lds = (C.c_int64 * 2)(512, 512); kinds = (C.c_int32 * 2)(0, 0)
def call(chunk):
    h.check(h.lib.scimlop_kron_mul_host(h.ptr, 0, 0, 2, ptrs, dims, lds, kinds, C.c_void_p(Vh.data_ptr()), n, C.c_void_p(Wh.data_ptr()), n, k, None, None, chunk), "host")
for chunk in (0, 2, 4, 6, 8, 12, 16, 32):
    call(chunk); call(chunk); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): call(chunk)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"chunk_cols={chunk:3d}: {dt*1e3:.2f} ms  {1.374e11/dt/1e12:.2f} TFLOP/s")
# raw copies
d = torch.empty(n * k, dtype=torch.float64, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(Vh, non_blocking=True)), ("D2H", lambda: Wh.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n*k*8/dt/1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): d.copy_(Vh, non_blocking=True)
    with torch.cuda.stream(s2): Wh.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"bidirectional: {n*k*8/dt/1e9:.1f} GB/s per direction")
