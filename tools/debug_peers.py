import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
kper = int(os.environ.get("KPER", 32)); n = 512 * 512; k = kper * world
rng = np.random.default_rng(0)
A, B = np.asfortranarray(rng.random((512, 512))), np.asfortranarray(rng.random((512, 512)))
g = torch.Generator(device="cuda").manual_seed(rank)
V = sb.DeviceArray(torch.rand(n * kper, dtype=torch.float64, device="cuda", generator=g), (n, kper))
L = sb.cache_operator(sb.kron(A, B), V)
r1, r2 = sb.DeviceArray.zeros((n, kper), np.float64), sb.DeviceArray.zeros((n, kper), np.float64)
sb.mul_(r1, L, V); sb.mul_(r2, L, V); torch.cuda.synchronize()
out = {"plain_vs_plain": float((r1.t - r2.t).abs().max())}
W = sb.DeviceArray.zeros((n, k), np.float64)
def gather_obj(o):
    res = [None] * world; dist.all_gather_object(res, o); return res
win = sd.PeerWindow(h, W, rank, world, gather_obj)
full = [torch.empty_like(r1.t) for _ in range(world)]
dist.all_gather(full, r1.t); torch.cuda.synchronize()
REF = torch.cat(full)
for rep in range(int(os.environ.get("REPS", 6))):
    W.t.zero_(); torch.cuda.synchronize(); dist.barrier()
    sd.kron_mul_peers(h, L, V, win, kper); torch.cuda.synchronize(); dist.barrier()
    mine = W.cols(rank * kper, (rank + 1) * kper).t
    d = (mine - r1.t).abs()
    out[f"fused_rep{rep}_local_slab_diff"] = float(d.max())
    if float(d.max()) > 0:
        idx = torch.nonzero(d.reshape(kper, n) > 0)
        out[f"fused_rep{rep}_bad_count"] = int(idx.shape[0])
        out[f"fused_rep{rep}_bad_cols"] = sorted(set(idx[:, 0].tolist()))[:10]
        rows = idx[:, 1]
        out[f"fused_rep{rep}_bad_rows_minmax"] = [int(rows.min()), int(rows.max())]
        out[f"fused_rep{rep}_bad_row_mod512"] = sorted(set((rows % 512).tolist()))[:12]
        out[f"fused_rep{rep}_bad_row_div512"] = sorted(set((rows // 512).tolist()))[:12]
    dall = (W.t - REF).abs().reshape(world, kper, 512, 512)
    for r in range(world):
        if r != rank and float(dall[r].max()) > 0:
            idx = torch.nonzero(dall[r] > 0)
            out[f"fused_rep{rep}_REMOTE_slab{r}"] = {"max": float(dall[r].max()), "count": int(idx.shape[0]),
                "k": sorted(set(idx[:, 0].tolist())), "mo": sorted(set(idx[:, 1].tolist()))[:40], "l": sorted(set(idx[:, 2].tolist()))[:40]}
    if float(d.max()) > 0:
        idx = torch.nonzero(dall[rank] > 0)
        out[f"fused_rep{rep}_LOCAL"] = {"count": int(idx.shape[0]), "k": sorted(set(idx[:, 0].tolist())),
            "mo": sorted(set(idx[:, 1].tolist())), "l": sorted(set(idx[:, 2].tolist()))}
    # plain again after fused
    sb.mul_(r2, L, V); torch.cuda.synchronize()
    out[f"plain_after_rep{rep}"] = float((r1.t - r2.t).abs().max())
print(rank, json.dumps(out), flush=True)
win.close(); dist.destroy_process_group()
