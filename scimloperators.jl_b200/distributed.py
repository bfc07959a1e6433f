"""Multi-GPU: column sharding with replicated factors.

Every operator of the path acts column-wise on V (src/SciMLOperators.jl:177-198), and columns are
contiguous in column-major storage, so a shard is a contiguous slab of V and of W.  One process per GPU;
no collective on the compute path.  `allgather_cols` (NCCL over NVLink) exists only for callers that want
the whole W replicated on every rank."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def shard_columns(k, nranks, rank):
    """(first, count): rank's contiguous column slab of a k-column batch (scimlop_shard_columns)."""
    first, count = C.c_int64(0), C.c_int64(0)
    rc = _lib.load().scimlop_shard_columns(int(k), int(nranks), int(rank), C.byref(first), C.byref(count))
    if rc != _lib.OK:
        raise _lib.ScimlopError(rc, "scimlop_shard_columns")
    return first.value, count.value


class ColumnShard:
    """This rank's view of a globally (n x k) batch: columns [first, first+count)."""

    def __init__(self, k, nranks, rank):
        self.k, self.nranks, self.rank = int(k), int(nranks), int(rank)
        self.first, self.count = shard_columns(k, nranks, rank)

    def take(self, host_matrix):
        """Slice this rank's columns out of a host (NumPy) matrix."""
        return np.asfortranarray(host_matrix[:, self.first:self.first + self.count])

    def shard_batched_diag(self, host_diag):
        """A BatchedDiagonalOperator's diagonal is batch-shaped, so it shards with V (src/batch.jl:6,44-47)."""
        return self.take(host_diag)


def init_comm(handle, nranks, rank, broadcast_bytes):
    """Create the NCCL communicator. `broadcast_bytes(b: bytes|None) -> bytes` broadcasts rank 0's 128-byte
    unique id (e.g. via torch.distributed.broadcast_object_list)."""
    lib = handle.lib
    buf = C.create_string_buffer(128)
    if rank == 0:
        handle.check(lib.scimlop_comm_unique_id(buf), "scimlop_comm_unique_id")
    uid = broadcast_bytes(bytes(buf.raw) if rank == 0 else None)
    buf2 = C.create_string_buffer(uid, 128)
    handle.check(lib.scimlop_comm_init(handle.ptr, nranks, rank, buf2), "scimlop_comm_init")


def allgather_cols(handle, W_full, cols_per_rank, stream=None):
    """In-place all-gather of equal column slabs of W_full (rows x cols_per_rank*nranks, contiguous)."""
    import torch
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
    rows = W_full.shape[0]
    handle.check(handle.lib.scimlop_allgather_cols(handle.ptr, _lib.F64 if W_full.dtype == np.float64 else _lib.F32,
                                                   C.c_void_p(W_full.ptr), rows, int(cols_per_rank), st),
                 "scimlop_allgather_cols")
    return W_full


def kron_mul_allgather(handle, L, V_local, W_full, cols_per_rank, alpha=None, chunk_cols=0, stream=None):
    """Column-sharded `mul!` of a cached, natively flattened TensorProductOperator with the all-gather of W
    overlapped chunk by chunk (scimlop_kron_mul_allgather).  W_full ends up replicated on every rank."""
    import torch
    from .operators import _DT, _scalar
    nat = L._native
    if nat is None:
        raise TypeError("kron_mul_allgather needs a TensorProductOperator of native factors (cache_operator it first)")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
    a = alpha
    for s in nat["scalars"]:
        a = s.convert() * (1.0 if a is None else float(a))
    handle.check(handle.lib.scimlop_kron_mul_allgather(
        handle.ptr, _DT[V_local.dtype], L.precision, nat["nf"], nat["ptrs"], nat["dims"], nat["lds"], nat["kinds"],
        C.c_void_p(V_local.ptr), C.c_void_p(W_full.ptr), int(cols_per_rank), _scalar(V_local.dtype, a), int(chunk_cols),
        C.c_void_p(nat["ws"].data_ptr()), nat["ws_bytes"], st), "scimlop_kron_mul_allgather")
    return W_full


class PeerWindow:
    """Every rank's W_full mapped into this process (CUDA IPC): what the fused compute + all-gather epilogue
    stores into.  `all_gather_object(obj) -> [obj of rank 0, ...]` is the caller's rendezvous
    (e.g. torch.distributed.all_gather_object)."""

    def __init__(self, handle, W_full, rank, world, all_gather_object):
        self.h, self.rank, self.world, self.W = handle, int(rank), int(world), W_full
        buf, off = C.create_string_buffer(64), C.c_int64(0)
        handle.check(handle.lib.scimlop_peer_export(handle.ptr, C.c_void_p(W_full.ptr), buf, C.byref(off)), "scimlop_peer_export")
        everyone = all_gather_object((bytes(buf.raw), off.value))
        self.base, self._opened = [None] * world, []
        for r, (hb, o) in enumerate(everyone):
            if r == rank:
                self.base[r] = W_full.ptr
                continue
            p = C.c_void_p()
            handle.check(handle.lib.scimlop_peer_open(handle.ptr, C.create_string_buffer(hb, 64), o, C.byref(p)), "scimlop_peer_open")
            self.base[r] = p.value
            self._opened.append((p.value, o))

    def close(self):
        for p, o in self._opened:
            self.h.lib.scimlop_peer_close(self.h.ptr, C.c_void_p(p), o)
        self._opened = []


def kron_mul_peers(handle, L, V_local, window, cols_per_rank, alpha=None, stream=None, barrier=True, pre_barrier=False):
    """This rank's column slab of W = (kron factors) V, stored by the stage-2 GEMM epilogue into EVERY rank's
    W_full (peer stores over NVLink), then a one-word stream barrier.  Float64, replicated result.
    `pre_barrier`: a caller that READS W between two calls must order "every rank has finished reading the previous
    W" before any rank's next remote stores (include/scimlop_b200.h, ordering contract): a barrier on the stream in
    front of the launch does that."""
    import torch
    from .operators import _DT, _scalar
    nat = L._native
    if nat is None:
        raise TypeError("kron_mul_peers needs a TensorProductOperator of native factors (cache_operator it first)")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
    rows = window.W.shape[0]
    es = 8
    slab = window.rank * int(cols_per_rank) * rows * es
    peers = [window.base[r] + slab for r in range(window.world) if r != window.rank]
    arr = (C.c_void_p * max(1, len(peers)))(*peers)
    a = alpha
    for s in nat["scalars"]:
        a = s.convert() * (1.0 if a is None else float(a))
    if pre_barrier:
        handle.check(handle.lib.scimlop_comm_barrier(handle.ptr, st), "scimlop_comm_barrier")
    handle.check(handle.lib.scimlop_kron_mul_peers(
        handle.ptr, _DT[V_local.dtype], L.precision, nat["nf"], nat["ptrs"], nat["dims"], nat["lds"], nat["kinds"],
        C.c_void_p(V_local.ptr), int(cols_per_rank), C.c_void_p(window.base[window.rank] + slab), arr, len(peers),
        _scalar(V_local.dtype, a), C.c_void_p(nat["ws"].data_ptr()), nat["ws_bytes"], st), "scimlop_kron_mul_peers")
    if barrier:
        handle.check(handle.lib.scimlop_comm_barrier(handle.ptr, st), "scimlop_comm_barrier")
    return window.W


class _RawCuda:
    """A device allocation that torch did not make (cuMemMap'ed), exposed through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class MulticastWindow:
    """A replicated W on every rank of the node plus its NVSwitch multicast alias (scimlop_mc_*): one store through
    the alias lands in every rank's copy.  Collective: construct it on all ranks at the same time.
    `all_gather_object(obj)` and `barrier()` are the caller's rendezvous (torch.distributed's); the multicast
    object's file descriptor travels from rank 0 to the others over a Unix socket (SCM_RIGHTS)."""

    def __init__(self, handle, shape, dtype, rank, world, all_gather_object, barrier):
        import os, socket, tempfile
        import numpy as np
        import torch
        from .array import DeviceArray
        self.h, self.rank, self.world = handle, int(rank), int(world)
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        padded = C.c_size_t(0)
        handle.check(handle.lib.scimlop_mc_padded_bytes(handle.ptr, nbytes, world, C.byref(padded)), "scimlop_mc_padded_bytes")
        fd_out, win = C.c_int(-1), C.c_void_p()
        if rank == 0:
            handle.check(handle.lib.scimlop_mc_open(handle.ptr, padded.value, world, 0, -1, C.byref(fd_out), C.byref(win)),
                         "scimlop_mc_open")
            path = os.path.join(tempfile.mkdtemp(prefix="scimlop_mc_"), "fd.sock")
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path)
            srv.listen(world)
            all_gather_object(path)
            for _ in range(world - 1):
                conn, _ = srv.accept()
                socket.send_fds(conn, [b"fd"], [fd_out.value])
                conn.recv(2)          # the peer has imported the descriptor
                conn.close()
            srv.close()
            os.close(fd_out.value)
            os.unlink(path)
        else:
            path = all_gather_object(None)[0]
            cli = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            cli.connect(path)
            _, fds, _, _ = socket.recv_fds(cli, 16, 1)
            handle.check(handle.lib.scimlop_mc_open(handle.ptr, padded.value, world, rank, fds[0], C.byref(fd_out), C.byref(win)),
                         "scimlop_mc_open")
            cli.send(b"ok")
            cli.close()
            os.close(fds[0])
        self._win = win
        barrier()                       # every device has joined the multicast object
        loc, mc = C.c_void_p(), C.c_void_p()
        handle.check(handle.lib.scimlop_mc_bind(handle.ptr, win, C.byref(loc), C.byref(mc)), "scimlop_mc_bind")
        self.local_ptr, self.mc_ptr = loc.value, mc.value
        raw = torch.as_tensor(_RawCuda(loc.value, nbytes), device=torch.device("cuda", handle.device))
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        self.W = DeviceArray(raw.view(tdt), shape)
        barrier()                       # every copy is bound before anyone stores through the alias

    def close(self):
        if self._win:
            self.h.lib.scimlop_mc_free(self.h.ptr, self._win)
            self._win = None


def kron_mul_mc(handle, L, V_local, window, cols_per_rank, alpha=None, stream=None, barrier=True, pre_barrier=False):
    """This rank's column slab of W = (kron factors) V, stored ONCE by the last GEMM's epilogue through the NVSwitch
    multicast alias of W (multimem.st): the switch writes it into every rank's copy.  Float64.  `pre_barrier` as in
    kron_mul_peers (needed when W is read between calls)."""
    import torch
    from .operators import _DT, _scalar
    nat = L._native
    if nat is None:
        raise TypeError("kron_mul_mc needs a TensorProductOperator of native factors (cache_operator it first)")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
    rows = window.W.shape[0]
    slab = window.rank * int(cols_per_rank) * rows * 8
    a = alpha
    for s in nat["scalars"]:
        a = s.convert() * (1.0 if a is None else float(a))
    if pre_barrier:
        handle.check(handle.lib.scimlop_comm_barrier(handle.ptr, st), "scimlop_comm_barrier")
    handle.check(handle.lib.scimlop_kron_mul_mc(
        handle.ptr, _DT[V_local.dtype], L.precision, nat["nf"], nat["ptrs"], nat["dims"], nat["lds"], nat["kinds"],
        C.c_void_p(V_local.ptr), int(cols_per_rank), C.c_void_p(window.local_ptr + slab), C.c_void_p(window.mc_ptr + slab),
        _scalar(V_local.dtype, a), C.c_void_p(nat["ws"].data_ptr()), nat["ws_bytes"], st), "scimlop_kron_mul_mc")
    if barrier:
        handle.check(handle.lib.scimlop_comm_barrier(handle.ptr, st), "scimlop_comm_barrier")
    return window.W
