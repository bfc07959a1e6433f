"""Strong-scaled config 2 with a replicated result (run under torchrun): 256 columns split over the ranks,
kron mul! per rank, then W all-gathered -- sequentially (kernels, then ncclAllGather) and overlapped
(scimlop_kron_mul_allgather).  Rank 0 prints one JSON line."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import scimlop_b200 as sb
    from scimlop_b200 import distributed as sd
    h = sb.handle(local)
    def bcast(b):
        o = [b]; dist.broadcast_object_list(o, src=0); return o[0]
    sd.init_comm(h, world, rank, bcast)
    kper = int(os.environ.get("KPER", 256 // world))
    n, k = 512 * 512, kper * world
    rng = np.random.default_rng(0)
    A, B = np.asfortranarray(rng.random((512, 512))), np.asfortranarray(rng.random((512, 512)))
    g = torch.Generator(device="cuda").manual_seed(rank)
    V = sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda", generator=g), (n, kper))
    W = sb.DeviceArray.zeros((n, k), np.float64)
    L = sb.cache_operator(sb.kron(A, B), V)
    mine = W.cols(rank * kper, (rank + 1) * kper)
    def timed(fn, reps=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    t_comp = timed(lambda: sb.mul_(mine, L, V))
    t_seq = timed(lambda: (sb.mul_(mine, L, V), sd.allgather_cols(h, W, kper)))
    res = {}
    for chunk in (max(1, kper // 8), max(1, kper // 4)):
        res[f"nccl_overlapped_chunk{chunk}_ms"] = timed(lambda: sd.kron_mul_allgather(h, L, V, W, kper, chunk_cols=chunk))
    def gather_obj(o):
        out = [None] * world
        dist.all_gather_object(out, o)
        return out
    win = sd.PeerWindow(h, W, rank, world, gather_obj)
    res["fused_peer_store_epilogue_ms"] = timed(lambda: sd.kron_mul_peers(h, L, V, win, kper))
    res["fused_peer_store_no_barrier_ms"] = timed(lambda: sd.kron_mul_peers(h, L, V, win, kper, barrier=False))
    # correctness of the replicated result against the local-only computation of this rank's slab
    ref = sb.DeviceArray.zeros((n, kper), np.float64)
    sb.mul_(ref, L, V)
    W.t.zero_(); torch.cuda.synchronize(); dist.barrier()
    sd.kron_mul_peers(h, L, V, win, kper); torch.cuda.synchronize(); dist.barrier()
    res["fused_slab_max_abs_diff"] = float((W.cols(rank * kper, (rank + 1) * kper).t - ref.t).abs().max())
    # every slab against an independent recomputation from that rank's V (seeded by rank)
    worst = 0.0
    for r in range(world):
        gr = torch.Generator(device="cuda").manual_seed(r)
        Vr = sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda", generator=gr), (n, kper))
        sb.mul_(ref, L, Vr)
        worst = max(worst, float((W.cols(r * kper, (r + 1) * kper).t - ref.t).abs().max()))
    res["fused_all_slabs_max_abs_diff"] = worst
    res["fused_all_slabs_nonzero"] = bool((W.t != 0).all())
    win.close()
    # NVSwitch multicast: one multimem.st per element, the switch replicates into every rank's copy
    try:
        mcw = sd.MulticastWindow(h, (n, k), np.float64, rank, world, gather_obj, dist.barrier)
    except Exception as e:
        mcw = None
        res["multicast_unavailable"] = str(e)[:200]
    if mcw is not None:
        res["fused_multicast_ms"] = timed(lambda: sd.kron_mul_mc(h, L, V, mcw, kper))
        res["fused_multicast_no_barrier_ms"] = timed(lambda: sd.kron_mul_mc(h, L, V, mcw, kper, barrier=False))
        mcw.W.t.zero_(); torch.cuda.synchronize(); dist.barrier()
        sd.kron_mul_mc(h, L, V, mcw, kper); torch.cuda.synchronize(); dist.barrier()
        worst = 0.0
        for r in range(world):
            gr = torch.Generator(device="cuda").manual_seed(r)
            Vr = sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda", generator=gr), (n, kper))
            sb.mul_(ref, L, Vr)
            worst = max(worst, float((mcw.W.cols(r * kper, (r + 1) * kper).t - ref.t).abs().max()))
        res["multicast_all_slabs_max_abs_diff"] = worst
        mcw.close()
    if rank == 0:
        print(json.dumps({"workload": "config 2 strong-scaled, W replicated", "n_gpus": world, "cols_per_rank": kper,
                          "compute_only_ms": t_comp, "compute_then_allgather_ms": t_seq, **res,
                          "gathered_bytes_per_rank": (world - 1) * n * kper * 8}))
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
