"""Multi-GPU path on the CPU: world_size-2 gloo processes exercise the column-sharding logic end to end
(shard -> per-rank apply -> gather == unsharded apply).  The per-rank `mul!` is the oracle here (no GPU in this
suite); what is under test is that sharding by contiguous column slabs with replicated factors and sharded
batched diagonals needs NO data-path collective and reproduces the global result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, k):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import scimlop_b200 as sb
    from oracle import ref_numpy as R
    rng = np.random.default_rng(0)                       # same seed on every rank: replicated factors
    A, B = rng.random((6, 5)), rng.random((4, 7))
    N = 24
    d, Dk = rng.random(N), rng.random((N, k))
    Vk, Vt = np.asfortranarray(rng.random((35, k))), np.asfortranarray(rng.random((N, k)))
    sh = sb.ColumnShard(k, world, rank)
    # kron path: replicated factors, sharded V
    v_loc = sh.take(Vk)
    w_loc = R.kron_apply([A, B], v_loc) if sh.count else np.zeros((24, 0))
    # tree path with a batched diagonal: the diagonal shards with V (src/batch.jl:6,44-47)
    Lloc = R.AddedOperator(R.ComposedOperator(R.DiagonalOperator(d), R.BatchedDiagonalOperator(sh.shard_batched_diag(Dk))),
                           R.IdentityOperator(N))
    t_loc = R.dense_apply(Lloc, sh.take(Vt)) if sh.count else np.zeros((N, 0))
    # gather only to CHECK (the product path has no collective); pad to the largest slab for all_gather
    cmax = (k + world - 1) // world
    def gather(x, rows):
        buf = torch.zeros(rows, cmax, dtype=torch.float64)
        buf[:, :x.shape[1]] = torch.from_numpy(np.ascontiguousarray(x))
        outs = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(outs, buf)
        cols = [sb.shard_columns(k, world, r)[1] for r in range(world)]
        return np.concatenate([o.numpy()[:, :c] for o, c in zip(outs, cols)], axis=1)
    Wg, Tg = gather(w_loc, 24), gather(t_loc, N)
    if rank == 0:
        assert R.rel_err(Wg, np.kron(A, B) @ Vk) < 1e-13
        Lfull = R.AddedOperator(R.ComposedOperator(R.DiagonalOperator(d), R.BatchedDiagonalOperator(Dk)),
                                R.IdentityOperator(N))
        assert R.rel_err(Tg, R.dense_apply(Lfull, Vt)) < 1e-13
    dist.barrier()
    dist.destroy_process_group()


def test_column_sharding_world2_gloo():
    for k in (8, 7, 1):       # even split, ragged split, fewer columns than ranks
        mp.spawn(_worker, args=(2, _free_port(), k), nprocs=2, join=True)
