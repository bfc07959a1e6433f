"""ctypes loader for the C restatement (oracle/ref_c.c). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_I64 = ctypes.c_int64


def build():
    src = os.path.join(_HERE, "ref_c.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _Lib:
    def __init__(self, lib):
        self.lib = lib
        for s, T in (("d", ctypes.c_double), ("s", ctypes.c_float)):
            P = ctypes.c_void_p
            getattr(lib, f"oracle_gemm_{s}").argtypes = [ctypes.c_int, _I64, _I64, _I64, T, P, _I64, P, _I64, T, P, _I64]
            getattr(lib, f"oracle_permute213_{s}").argtypes = [P, P, _I64, _I64, _I64]
            getattr(lib, f"oracle_lv_outer_mul_{s}").argtypes = [P, P, _I64, P, _I64, _I64, _I64, _I64, ctypes.c_int, T, T]
            f = getattr(lib, f"oracle_tpo2_mul_{s}")
            f.argtypes = [P, P, P, P, _I64, _I64, _I64, _I64, _I64, ctypes.c_int, T, T, ctypes.c_int]
            f.restype = ctypes.c_int
            getattr(lib, f"oracle_diag_mul_{s}").argtypes = [P, P, P, _I64, _I64, ctypes.c_int, ctypes.c_int, T, T]
            getattr(lib, f"oracle_axpby_{s}").argtypes = [P, P, _I64, T, T]
            getattr(lib, f"oracle_lmul_{s}").argtypes = [P, _I64, T]

    @staticmethod
    def _sfx(a):
        return {np.dtype("float64"): "d", np.dtype("float32"): "s"}[a.dtype]

    @staticmethod
    def _p(a):
        assert a.flags.f_contiguous or a.ndim == 1
        return a.ctypes.data

    def tpo2_mul(self, w, A, B, v, alpha=None, beta=None, lv=True):
        """w = [α](A⊗B)v[+βw] following the reference's 2-factor `mul!` (see ref_c.c)."""
        mo, no = A.shape
        mi, ni = B.shape
        k = v.shape[1] if v.ndim == 2 else 1
        use_ab = alpha is not None
        f = getattr(self.lib, "oracle_tpo2_mul_" + self._sfx(w))
        rc = f(self._p(w), self._p(A), self._p(B), self._p(v), mo, no, mi, ni, k, int(use_ab),
               alpha if use_ab else 1.0, beta if use_ab else 0.0, int(lv))
        if rc != 0:
            raise MemoryError("oracle scratch allocation failed")
        return w

    def lv_outer_mul(self, w, A, C, mi, mo, no, k, alpha=None, beta=None):
        use_ab = alpha is not None
        f = getattr(self.lib, "oracle_lv_outer_mul_" + self._sfx(w))
        f(self._p(w), self._p(A), A.shape[0], self._p(C), mi, mo, no, k, int(use_ab),
          alpha if use_ab else 1.0, beta if use_ab else 0.0)
        return w

    def gemm(self, C, A, B, alpha=1.0, beta=0.0, ta=False):
        M, N = C.shape
        K = B.shape[0]
        f = getattr(self.lib, "oracle_gemm_" + self._sfx(C))
        f(int(ta), M, N, K, alpha, self._p(A), A.shape[0], self._p(B), B.shape[0], beta, self._p(C), C.shape[0])
        return C

    def permute213(self, dst, src):
        d1, d2, d3 = src.shape
        getattr(self.lib, "oracle_permute213_" + self._sfx(dst))(self._p(dst), self._p(src), d1, d2, d3)
        return dst

    def diag_mul(self, w, d, v, alpha=None, beta=None):
        n = v.shape[0]
        k = v.shape[1] if v.ndim == 2 else 1
        use_ab = alpha is not None
        f = getattr(self.lib, "oracle_diag_mul_" + self._sfx(w))
        f(self._p(w), self._p(d), self._p(v), n, k, int(d.ndim == 2), int(use_ab),
          alpha if use_ab else 1.0, beta if use_ab else 0.0)
        return w

    def axpby(self, w, v, alpha, beta):
        getattr(self.lib, "oracle_axpby_" + self._sfx(w))(self._p(w), self._p(v), w.size, alpha, beta)
        return w


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = _Lib(ctypes.CDLL(build()))
    return _cached
