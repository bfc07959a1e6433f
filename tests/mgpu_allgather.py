"""Multi-GPU check (run under torchrun on >= 2 GPUs; driven by tests/test_multigpu.py on a multi-GPU box):
column-sharded kron mul! with replicated factors, then the OPTIONAL in-place NCCL all-gather of W
(scimlop_allgather_cols) -- every rank must end up with the full, correct W."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import scimlop_b200 as sb
    from scimlop_b200 import distributed as sd
    from oracle import ref_numpy as R
    h = sb.handle(local)
    rng = np.random.default_rng(0)                       # same seed everywhere: replicated factors and global V
    A, B = np.asfortranarray(rng.random((48, 40))), np.asfortranarray(rng.random((32, 24)))
    kper = 6
    k = kper * world
    V = np.asfortranarray(rng.random((40 * 24, k)))
    sh = sb.ColumnShard(k, world, rank)
    assert (sh.first, sh.count) == (rank * kper, kper)
    # NCCL communicator of the library (unique id travels through torch.distributed)
    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    sd.init_comm(h, world, rank, bcast)
    Wfull = sb.DeviceArray.full((48 * 32, k), float("nan"), np.float64)
    v_loc = sb.DeviceArray.from_numpy(sh.take(V))
    L = sb.cache_operator(sb.kron(A, B), v_loc)
    sb.mul_(Wfull.cols(sh.first, sh.first + sh.count), L, v_loc)     # this rank's slab, in place in the full W
    sd.allgather_cols(h, Wfull, kper)
    torch.cuda.synchronize()
    err = R.rel_err(Wfull.numpy(), np.kron(A, B) @ V)
    # the overlapped form: compute chunk i+1 while chunk i is exchanged
    W2 = sb.DeviceArray.full((48 * 32, k), float("nan"), np.float64)
    sd.kron_mul_allgather(h, L, v_loc, W2, kper, alpha=0.5, chunk_cols=2)
    torch.cuda.synchronize()
    err = max(err, R.rel_err(W2.numpy(), 0.5 * (np.kron(A, B) @ V)))
    # the fused form: stage-2 epilogue stores every tile into all ranks' W (peer memory), no collective
    def gather_obj(o):
        out = [None] * world
        dist.all_gather_object(out, o)
        return out
    Ab, Bb = np.asfortranarray(rng.random((128, 96))), np.asfortranarray(rng.random((64, 80)))
    Vb = np.asfortranarray(rng.random((96 * 80, k)))
    W3 = sb.DeviceArray.full((128 * 64, k), float("nan"), np.float64)
    win = sd.PeerWindow(h, W3, rank, world, gather_obj)
    vb_loc = sb.DeviceArray.from_numpy(sh.take(Vb))
    Lb = sb.cache_operator(sb.kron(Ab, Bb), vb_loc)
    want3 = 0.5 * (np.kron(Ab, Bb) @ Vb)
    for rep in range(3):
        W3.t.fill_(float("nan"))
        torch.cuda.synchronize(); dist.barrier()
        sd.kron_mul_peers(h, Lb, vb_loc, win, kper, alpha=0.5)
        torch.cuda.synchronize()
        got3 = W3.numpy()
        e3 = R.rel_err(got3, want3)
        if e3 > 1e-12:
            bad = [r for r in range(world) if R.rel_err(got3[:, r * kper:(r + 1) * kper], want3[:, r * kper:(r + 1) * kper]) > 1e-12]
            print(f"rank {rank} rep {rep}: fused result wrong, rel err {e3:.3e}, bad slabs {bad}", flush=True)
        err = max(err, e3)
        dist.barrier()
    win.close()
    # headline factor shape, 32 columns per rank, bit-exact against the unfused path, repeated: this is the case
    # that exposed a ring-slot release racing ld.shared under NVLink store back-pressure
    Ac, Bc = np.asfortranarray(rng.random((512, 512))), np.asfortranarray(rng.random((512, 512)))
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    kc = 32
    Vc = sb.DeviceArray(torch.rand(512 * 512 * kc, dtype=torch.float64, device="cuda", generator=g), (512 * 512, kc))
    Lc = sb.cache_operator(sb.kron(Ac, Bc), Vc)
    plain = sb.DeviceArray.zeros((512 * 512, kc), np.float64)
    # one summation order, as the fused entries use (512 tiles: the default would run the last partial wave split-K)
    h.set_option(sb.OPT_TAIL_SPLIT, 0)
    sb.mul_(plain, Lc, Vc)
    h.set_option(sb.OPT_TAIL_SPLIT, 1)
    h.set_option(sb.OPT_PEER_PIPELINE, 2)    # column-chunk pipeline forced on: must stay bit-identical
    slabs = [torch.empty_like(plain.t) for _ in range(world)]
    dist.all_gather(slabs, plain.t)
    want4 = torch.cat(slabs)
    W4 = sb.DeviceArray.zeros((512 * 512, kc * world), np.float64)
    win4 = sd.PeerWindow(h, W4, rank, world, gather_obj)
    for rep in range(12):
        W4.t.zero_()
        torch.cuda.synchronize(); dist.barrier()
        sd.kron_mul_peers(h, Lc, Vc, win4, kc)
        torch.cuda.synchronize()
        d4 = float((W4.t - want4).abs().max())
        if d4 != 0.0:
            print(f"rank {rank} rep {rep}: fused 512x512 result differs from the unfused path by {d4:.3e}", flush=True)
            err = max(err, 1.0)
        dist.barrier()
    win4.close()
    # the same through the NVSwitch multicast alias (one multimem.st per element, replicated by the switch)
    try:
        mcw = sd.MulticastWindow(h, (512 * 512, kc * world), np.float64, rank, world, gather_obj, dist.barrier)
    except Exception as e:   # no multicast on this node (no NVSwitch / driver support): reported, not failed
        mcw = None
        if rank == 0:
            print(f"MULTICAST_UNAVAILABLE {e}", flush=True)
    if mcw is not None:
        for rep in range(6):
            mcw.W.t.zero_()
            torch.cuda.synchronize(); dist.barrier()
            sd.kron_mul_mc(h, Lc, Vc, mcw, kc)
            torch.cuda.synchronize(); dist.barrier()
            d5 = float((mcw.W.t - want4).abs().max())
            if d5 != 0.0:
                print(f"rank {rank} rep {rep}: multicast result differs from the unfused path by {d5:.3e}", flush=True)
                err = max(err, 1.0)
            dist.barrier()
        if rank == 0:
            print("MULTICAST_OK" if err < 1e-12 else "MULTICAST_FAILED", flush=True)
        mcw.close()
    ok = torch.tensor([1.0 if err < 1e-12 else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("ALLGATHER_OK" if ok.item() == 1.0 else f"ALLGATHER_FAILED err={err:.3e}")
    dist.destroy_process_group()
    return 0 if ok.item() == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
