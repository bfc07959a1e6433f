"""Device-resident solver loop around the operator.

`cg` is a batched conjugate-gradient iteration for `L X = B` where every column of B is an independent
system -- the way LinearSolve's Krylov methods drive `mul!(w, L, v)` thousands of times with fixed shapes.
Everything between two matvecs (dot products, per-column step lengths, the x / r / p updates) runs in the
library's vector kernels with the per-column scalars kept in device memory, so an iteration never
synchronises with the host and a block of iterations can be captured in one CUDA graph."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .array import DeviceArray
from .operators import _DT, _ncols, _stream, mul_


def col_dot(X, Y, out=None):
    """out[j] = <X[:,j], Y[:,j]> on the device (scimlop_col_dot)."""
    h = _lib.handle(X.device_index)
    n, k = X.shape[0], _ncols(X)
    if out is None:
        out = DeviceArray.empty((k,), X.dtype, X.device_index)
    h.check(h.lib.scimlop_col_dot(h.ptr, _DT[X.dtype], n, k, C.c_void_p(X.ptr), X.ld, C.c_void_p(Y.ptr), Y.ld,
                                  C.c_void_p(out.ptr), _stream()), "scimlop_col_dot")
    return out


def col_axpby(X, W, sa=1.0, an=None, ad=None, sb=1.0, bn=None, bd=None, dot=None):
    """W[:,j] = (sa*an[j]/ad[j]) X[:,j] + (sb*bn[j]/bd[j]) W[:,j]; optional dot[j] = |W_new[:,j]|^2."""
    h = _lib.handle(W.device_index)
    n, k = W.shape[0], _ncols(W)
    p = lambda a: C.c_void_p(a.ptr) if a is not None else None
    h.check(h.lib.scimlop_col_axpby_dot(h.ptr, _DT[W.dtype], n, k, C.c_void_p(X.ptr), X.ld, float(sa), p(an), p(ad),
                                        C.c_void_p(W.ptr), W.ld, float(sb), p(bn), p(bd), p(dot), _stream()),
            "scimlop_col_axpby_dot")
    return W


class CGState:
    """Scratch of a batched CG solve (allocated once, like a cached operator's scratch)."""

    def __init__(self, L, B):
        n, k = B.shape[0], _ncols(B)
        mk = lambda: DeviceArray.zeros(B.shape, B.dtype, B.device_index)
        self.r, self.p, self.Ap = mk(), mk(), mk()
        self.rs = [DeviceArray.zeros((k,), B.dtype, B.device_index) for _ in range(2)]
        self.pAp = DeviceArray.zeros((k,), B.dtype, B.device_index)
        self.L = L if L.iscached() else L.cache_operator(B)
        self.graph = None
        self.iters_per_graph = 0


def _cg_iteration(st, X, parity):
    rs, rs_new = st.rs[parity], st.rs[parity ^ 1]
    mul_(st.Ap, st.L, st.p)                                               # Ap = L p            (the hot path)
    col_dot(st.p, st.Ap, st.pAp)                                          # pAp_j
    col_axpby(st.p, X, sa=1.0, an=rs, ad=st.pAp)                          # x += (rs/pAp) p
    col_axpby(st.Ap, st.r, sa=-1.0, an=rs, ad=st.pAp, dot=rs_new)         # r -= (rs/pAp) Ap ; rs_new = |r|^2
    col_axpby(st.r, st.p, sa=1.0, sb=1.0, bn=rs_new, bd=rs)               # p = r + (rs_new/rs) p


def cg(L, B, X, iters, state=None, use_graph=True, iters_per_graph=8):
    """Runs `iters` CG iterations on L X = B in place in X (all columns at once); returns (X, |r_j|^2 as a
    DeviceArray, state).  With `use_graph`, blocks of `iters_per_graph` iterations replay one CUDA graph."""
    st = state or CGState(L, B)
    # r = B - L X ; p = r ; rs = |r|^2
    mul_(st.r, st.L, X)
    col_axpby(B, st.r, sa=1.0, sb=-1.0)
    st.p.t.copy_(st.r.t)
    col_dot(st.r, st.r, st.rs[0])
    done = 0
    if use_graph and iters >= iters_per_graph and iters_per_graph % 2 == 0:
        if st.graph is None or st.iters_per_graph != iters_per_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for i in range(iters_per_graph):
                        _cg_iteration(st, X, i & 1)
            torch.cuda.current_stream().wait_stream(side)
            st.graph, st.iters_per_graph, st.graph_x = g, iters_per_graph, X.ptr
            # capture does not execute: nothing has been applied yet
        if st.graph_x != X.ptr:
            raise AssertionError("the captured CG graph is bound to the X it was captured with")
        while iters - done >= iters_per_graph:
            st.graph.replay()
            done += iters_per_graph
    parity = 0  # an even number of graph iterations leaves rs[0] current
    for i in range(iters - done):
        _cg_iteration(st, X, parity)
        parity ^= 1
    return X, st.rs[parity], st


def _scalar_signature(L, seen=None):
    """Values of every ScalarOperator reachable from L (ScaledOperator λ's, WOperator γ, ...): what a captured launch
    bakes into its kernel parameters (src/scalar.jl:267-270 updates them on the host)."""
    from .operators import AbstractSciMLOperator, ScalarOperator
    seen = set() if seen is None else seen
    if id(L) in seen:
        return ()
    seen.add(id(L))
    sig = []
    for val in vars(L).values() if hasattr(L, "__dict__") else ():
        items = val if isinstance(val, (tuple, list)) else (val,)
        for it in items:
            if isinstance(it, ScalarOperator):
                sig.append(float(it.convert()))
            elif isinstance(it, AbstractSciMLOperator):
                sig.extend(_scalar_signature(it, seen))
    return tuple(sig)


class CapturedMul:
    """`mul!(w, L, v[, α, β])` with fixed buffers captured once into a CUDA graph: an ODE / Krylov loop replays it.
    `update_coefficients!` between replays is honoured: the graph holds POINTERS to the operator's device arrays, so
    in-place updates of matrices / diagonals are seen directly; host scalars (α, β and every ScalarOperator λ, which
    the kernels take by value) are compared at each replay and the graph is re-captured when one changed.  For small
    operators (config 1, vector inputs) this removes the per-call host work between the launches.
    Construction does not modify `w` (the warm-up outside the capture runs into a scratch copy when β ≠ 0)."""

    def __init__(self, L, w, v, alpha=None, beta=None):
        self.L, self.w, self.v, self.alpha, self.beta = L, w, v, alpha, beta
        self.stream = torch.cuda.Stream(device=w.t.device)
        self.captures = 0
        self._capture(warm=True)

    def _capture(self, warm):
        L, w, v = self.L, self.w, self.v
        self.stream.wait_stream(torch.cuda.current_stream(w.t.device))   # pending producers of v / w
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(self.stream):
            if warm:
                # lazy set-up (module loading, handle creation) must happen outside the capture; a real mul! with
                # beta != 0 would update w once at construction, so it runs into a scratch copy
                from .operators import _beta_is_zero
                target = w if _beta_is_zero(self.beta) else w.copy()
                L.mul_(target, v, self.alpha, self.beta)
                self.stream.synchronize()
            with torch.cuda.graph(self.graph, stream=self.stream):
                L.mul_(w, v, self.alpha, self.beta)
        self.stream.synchronize()
        self._sig = _scalar_signature(L)
        self.captures += 1

    def replay(self, alpha=None, beta=None):
        """Replays the captured mul!.  New α / β may be passed; changed scalars trigger one re-capture."""
        changed = False
        if alpha is not None and alpha != self.alpha:
            self.alpha, changed = alpha, True
        if beta is not None and beta != self.beta:
            self.beta, changed = beta, True
        if changed or _scalar_signature(self.L) != self._sig:
            self._capture(warm=False)
        self.graph.replay()
        return self.w
