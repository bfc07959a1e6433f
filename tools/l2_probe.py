"""Round-2 probe: how fast is the tcgen05 resident-factor kernel when its operands come from L2 instead of HBM,
and do column-chunked intermediates really stay in L2?

  (a) pure copy bandwidth at L2-resident and HBM sizes (torch copy_, read + write bytes);
  (b) one Float32 mode product (factor 128x128) on a tile count that is a multiple of the SM count, ping-ponging
      between two buffers that fit L2, N launches in one CUDA graph -> time per tile without the HBM bound;
  (c) config 3 (128^3, k columns) applied in column chunks through the existing three-launch path, all chunks
      in one CUDA graph, against the plain three-launch path.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb  # noqa: E402


def ev(fn, reps=5, warm=2):
    for _ in range(warm):
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
    return float(np.median(ts))


def graph_of(fn, n):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    return g


def copy_bw():
    for mb in (8, 16, 32, 48, 64, 256, 1024):
        n = mb * (1 << 20) // 4
        a = torch.rand(n, dtype=torch.float32, device="cuda")
        b = torch.empty_like(a)
        g = graph_of(lambda: b.copy_(a), 50)
        t = ev(g.replay) / 50
        print(json.dumps({"probe": "copy", "MB_each": mb, "us": round(t * 1e3, 2), "TBps_rw": round(2 * mb * (1 << 20) / t / 1e9, 2)}), flush=True)


def mode_l2():
    F = sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda") / 64, (128, 128))
    Mop = sb.MatrixOperator(F)
    for tiles_per_sm in (1, 2, 4, 8, 16):
        tiles = 148 * tiles_per_sm
        cols = tiles * 128
        X = sb.DeviceArray(torch.rand(128 * cols, dtype=torch.float32, device="cuda"), (128, cols))
        Y = sb.DeviceArray.empty((128, cols), np.float32)
        # innermost-mode geometry (XK): Y(128 x cols) = F * X
        def pp():
            sb.mul_(Y, Mop, X)
            sb.mul_(X, Mop, Y)
        g = graph_of(pp, 25)
        X.t.uniform_(0, 1e-3)
        t = ev(g.replay) / 50
        print(json.dumps({"probe": "mode XK", "tiles": tiles, "MB_in": round(cols * 512 / 2 ** 20, 1), "us_per_launch": round(t * 1e3, 2),
                          "us_per_tile_per_sm": round(t * 1e3 / tiles_per_sm, 3)}), flush=True)
        # outer-mode geometry (X_MN, batched): kron(F, I_128) on (128*128) x tiles columns
        k = tiles
        V = sb.DeviceArray(torch.rand(128 * 128 * k, dtype=torch.float32, device="cuda") * 1e-2, (128 * 128, k))
        W = sb.DeviceArray.empty((128 * 128, k), np.float32)
        L = sb.cache_operator(sb.kron(Mop, sb.IdentityOperator(128)), V)
        def pp2():
            sb.mul_(W, L, V)
            sb.mul_(V, L, W)
        g = graph_of(pp2, 25)
        V.t.uniform_(0, 1e-3)
        t = ev(g.replay) / 50
        print(json.dumps({"probe": "mode XMN", "tiles": tiles, "MB_in": round(k * 65536 / 2 ** 20, 1), "us_per_launch": round(t * 1e3, 2),
                          "us_per_tile_per_sm": round(t * 1e3 / tiles_per_sm, 3)}), flush=True)
        del X, Y, V, W, L, g
        torch.cuda.empty_cache()


def c3_chunked(k=512):
    F = [sb.DeviceArray(torch.rand(128 * 128, dtype=torch.float32, device="cuda"), (128, 128)) for _ in range(3)]
    n = 128 ** 3
    V = sb.DeviceArray(torch.rand(n * k, dtype=torch.float32, device="cuda"), (n, k))
    W = sb.DeviceArray.empty((n, k), np.float32)
    ops = [sb.MatrixOperator(f) for f in F]
    L = sb.cache_operator(sb.TensorProductOperator(*ops), V)
    t0 = ev(lambda: sb.mul_(W, L, V), reps=3, warm=1)
    print(json.dumps({"probe": "c3 plain", "k": k, "ms": round(t0, 3)}), flush=True)
    ref = W.t[: n * 2].clone()
    for chunk in (2, 4, 8, 16):
        Lc = sb.cache_operator(sb.TensorProductOperator(*ops), V.cols(0, chunk))
        def run():
            for j in range(0, k, chunk):
                sb.mul_(W.cols(j, j + chunk), Lc, V.cols(j, j + chunk))
        g = graph_of(run, 1)
        W.t.zero_()
        t = ev(g.replay, reps=3, warm=1)
        same = bool(torch.equal(W.t[: n * 2], ref))
        print(json.dumps({"probe": "c3 chunked graph", "chunk_cols": chunk, "launches": 3 * (k // chunk), "ms": round(t, 3),
                          "bit_identical": same}), flush=True)
        del g


if __name__ == "__main__":
    torch.cuda.set_device(0)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "copy"):
        copy_bw()
    if which in ("all", "mode"):
        mode_l2()
    if which in ("all", "c3"):
        c3_chunked(int(sys.argv[2]) if len(sys.argv) > 2 else 512)
