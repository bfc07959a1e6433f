import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
h = sb.handle(0)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(name, M, N, K, batch, ta, tb, shared_b, beta):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.rand((K*M if True else 0) * batch, device="cuda", generator=g)
    B = torch.rand(K*N * (1 if shared_b else batch), device="cuda", generator=g)
    Cc = torch.rand(M*N*batch, device="cuda", generator=g)
    pb = C.byref(C.c_float(beta)) if beta is not None else None
    lda = K if ta else M; ldb = N if tb else K
    rc = h.lib.scimlop_gemm_strided_batched(h.ptr, 1, 0, ta, tb, M, N, K, None, C.c_void_p(A.data_ptr()), lda, M*K,
        C.c_void_p(B.data_ptr()), ldb, 0 if shared_b else K*N, pb, C.c_void_p(Cc.data_ptr()), M, M*N, batch, st())
    try:
        h.check(rc, name); torch.cuda.synchronize()
        # spot check one column block
        print(name, "OK", float(Cc[:4].sum()), hex(A.data_ptr()), hex(B.data_ptr()), hex(Cc.data_ptr()))
    except Exception as e:
        print(name, "FAILED", str(e)[:150]); sys.exit(1)
mode = sys.argv[1] if len(sys.argv) > 1 else "rep"
if mode == "rep":
    for i in range(6):
        run(f"rep {i} (a) rows=524288 beta=None", 128, 524288, 128, 1, 0, 0, False, None)
elif mode == "mix":
    run("(a) rows=524288 beta=0.3", 128, 524288, 128, 1, 0, 0, False, 0.3)
    run("(a) rows=524288 beta=None", 128, 524288, 128, 1, 0, 0, False, None)
elif mode == "small_then_big":
    run("(a) rows=4096 beta=None", 128, 4096, 128, 1, 0, 0, False, None)
    run("(a) rows=524288 beta=None", 128, 524288, 128, 1, 0, 0, False, None)
