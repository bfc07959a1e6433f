"""Sweep of the column-chunk pipeline of the fused compute + all-gather entries (scimlop_kron_mul_mc / _peers,
include/scimlop_b200.h: SCIMLOP_OPT_PEER_PIPELINE, _PEER_CHUNKS, _PEER_STAGE2_SMS) on configs[1] strong-scaled over
the ranks of one node: 512 x 512 (x) 512 x 512 Float64, 256 columns in total.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 \
        tools/bench_peer_pipeline.py > profiles/r02_peer_pipeline_sweep.jsonl

One JSON line per (entry, chunks, stage-2 SMs): ms = max over ranks of the mean of 20 calls (device time, barrier on
both sides), every result bit-compared against the unchunked schedule."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import scimlop_b200 as sb
    from scimlop_b200 import distributed as sd
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = sb.handle(local)
    NO = NI = 512
    KCOLS = 256
    n = NO * NI
    rng = np.random.default_rng(0)
    A, B = np.asfortranarray(rng.random((NO, NO))), np.asfortranarray(rng.random((NI, NI)))
    first, kper = sd.shard_columns(KCOLS, world, rank)
    g = torch.Generator(device="cuda").manual_seed(4242 + rank)
    V = sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda", generator=g), (n, kper))
    L = sb.cache_operator(sb.kron(sb.MatrixOperator(A), sb.MatrixOperator(B)), V)

    def bcast(b):
        o = [b]
        dist.broadcast_object_list(o, src=0)
        return o[0]

    def gather_obj(o):
        res = [None] * world
        dist.all_gather_object(res, o)
        return res

    sd.init_comm(h, world, rank, bcast)

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        return float(t.item())

    W = sb.DeviceArray.zeros((n, KCOLS), np.float64)
    win = sd.PeerWindow(h, W, rank, world, gather_obj)
    try:
        mcw = sd.MulticastWindow(h, (n, KCOLS), np.float64, rank, world, gather_obj, dist.barrier)
    except Exception as e:
        mcw = None
        if rank == 0:
            print(json.dumps({"multicast_unavailable": str(e)[:160]}), flush=True)
    entries = [("peer", lambda: sd.kron_mul_peers(h, L, V, win, kper), W)]
    if mcw is not None:
        entries.insert(0, ("multicast", lambda: sd.kron_mul_mc(h, L, V, mcw, kper), mcw.W))
    shapes = [(0, 0)] + [(c, s) for c in (2, 4) for s in (40, 56, 64, 74, 88, 100, 116)]
    for name, fn, Wfull in entries:
        h.set_option(sb.OPT_PEER_PIPELINE, 0)
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        ref = Wfull.t.clone()
        for chunks, sms in shapes:
            h.set_option(sb.OPT_PEER_PIPELINE, 2 if chunks else 0)
            h.set_option(sb.OPT_PEER_CHUNKS, chunks)
            h.set_option(sb.OPT_PEER_STAGE2_SMS, sms)
            ms = timed(fn)
            Wfull.t.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            fn()
            torch.cuda.synchronize()
            dist.barrier()
            same = torch.tensor([1 if torch.equal(Wfull.t, ref) else 0], device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            if rank == 0:
                print(json.dumps({"entry": name, "ranks": world, "cols_per_rank": kper, "chunks": chunks, "stage2_sms": sms,
                                  "ms": round(ms, 4), "bit_identical_to_unchunked": bool(same.item())}), flush=True)
        h.set_option(sb.OPT_PEER_PIPELINE, 1)
        h.set_option(sb.OPT_PEER_CHUNKS, 0)
        h.set_option(sb.OPT_PEER_STAGE2_SMS, 0)
        ms = timed(fn)
        if rank == 0:
            print(json.dumps({"entry": name, "ranks": world, "cols_per_rank": kper, "chunks": "auto", "stage2_sms": "auto",
                              "ms": round(ms, 4)}), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    win.close()
    if mcw is not None:
        mcw.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    t0 = time.time()
    main()
