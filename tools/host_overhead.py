"""Host-side cost of one mul! call (no synchronisation) vs the device time, for small problems (config 1)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scimlop_b200 as sb
rng = np.random.default_rng(0)
for n, k in ((100, 100), (512, 1), (128, 1)):
    A, B = (np.asfortranarray(rng.random((n, n))) for _ in range(2))
    V = sb.DeviceArray(torch.rand(n * n * k, dtype=torch.float64, device="cuda"), (n * n, k))
    W = sb.DeviceArray.zeros((n * n, k), np.float64)
    L = sb.cache_operator(sb.kron(A, B), V)
    for _ in range(20): sb.mul_(W, L, V)
    torch.cuda.synchronize()
    N = 2000
    t0 = time.perf_counter()
    for _ in range(N): sb.mul_(W, L, V)
    t_issue = (time.perf_counter() - t0) / N
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / N
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sb.mul_(W, L, V)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            sb.mul_(W, L, V)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(N): g.replay()
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / N
    print(json.dumps({"n": n, "k": k, "host_issue_us": round(t_issue * 1e6, 1), "loop_total_us": round(t_total * 1e6, 1),
                      "graph_replay_us": round(t_graph * 1e6, 1)}), flush=True)
