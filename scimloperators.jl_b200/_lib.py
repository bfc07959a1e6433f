"""ctypes binding of libscimlop_b200.so (include/scimlop_b200.h) -- the same entry points the Julia package
extension `ccall`s.  There is no CPU fallback: if the shared library is missing, or no
sm_100 GPU is present when a handle is requested, this raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscimlop_b200.so")

# enums of include/scimlop_b200.h
OK = 0
(ERR_NULL_POINTER, ERR_BAD_SHAPE, ERR_WORKSPACE, ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_BAD_HANDLE, ERR_NO_DEVICE,
 ERR_SINGULAR) = range(1, 10)
F64, F32, C64, C128 = 0, 1, 2, 3
PREC_DEFAULT, PREC_TF32X3, PREC_TF32, PREC_FP32_EXACT, PREC_BF16 = 0, 1, 2, 3, 4
DENSE, IDENTITY, DIAG, BDIAG, NULL, BANDED, CSR = 0, 1, 2, 3, 4, 5, 6
TRANS = 0x100
OPT_TAIL_SPLIT = 1
OPT_PEER_PIPELINE = 2
OPT_PEER_CHUNKS = 3
OPT_PEER_STAGE2_SMS = 4
OPT_MODE_PIPELINE = 5


class Factor(C.Structure):  # scimlop_factor_t
    _fields_ = [("kind", C.c_int32), ("trans", C.c_int32), ("data", C.c_void_p), ("ld", C.c_int64),
                ("rows", C.c_int64), ("cols", C.c_int64)]


class Term(C.Structure):  # scimlop_term_t
    _fields_ = [("scale", C.c_double), ("nfactors", C.c_int32), ("factors", C.POINTER(Factor))]


class ScimlopError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: status {status} ({status_string(status)})" + (f": {detail}" if detail else ""))


_P, _I64, _I32, _INT, _SZ = C.c_void_p, C.c_int64, C.c_int32, C.c_int, C.c_size_t
_I64P, _I32P = C.POINTER(C.c_int64), C.POINTER(C.c_int32)

# name -> argtypes; every function returns int except scimlop_status_string
SIGNATURES = {
    "scimlop_version": [],
    "scimlop_status_string": [_INT],
    "scimlop_create": [C.POINTER(_P), _INT],
    "scimlop_destroy": [_P],
    "scimlop_last_error": [_P, C.c_char_p, _SZ],
    "scimlop_launch_count": [_P, _I64P],
    "scimlop_set_option": [_P, _INT, _INT],
    "scimlop_matmul": [_P, _INT, _INT, _INT, _I64, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _P],
    "scimlop_csr_workspace_bytes": [_INT, _INT, _I64, _I64, _I64, C.POINTER(_SZ)],
    "scimlop_csr_mul": [_P, _INT, _INT, _I64, _I64, _I64, _P, _P, _P, _INT, _P, _I64, _P, _I64, _P, _P, _P, _SZ, _P],
    "scimlop_split_bytes": [_I64, _I64, C.POINTER(_SZ)],
    "scimlop_split_attach": [_P, _P, _I64, _I64, _I64, _P, _SZ, _P],
    "scimlop_split_refresh": [_P, _P, _P],
    "scimlop_split_detach": [_P, _P],
    "scimlop_gemm_strided_batched": [_P, _INT, _INT, _INT, _INT, _I64, _I64, _I64, _P, _P, _I64, _I64, _P, _I64,
                                     _I64, _P, _P, _I64, _I64, _I64, _P],
    "scimlop_diag_mul": [_P, _INT, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _P],
    "scimlop_axpby": [_P, _INT, _I64, _I64, _P, _I64, _P, _I64, _P, _P, _P],
    "scimlop_diag_div": [_P, _INT, _I64, _I64, _P, _I64, _P, _I64, _P, _I64, _P],
    "scimlop_invert_workspace_bytes": [_INT, _I64, C.POINTER(_SZ)],
    "scimlop_invert": [_P, _INT, _I64, _P, _I64, _P, _I64, _P, _SZ, _P],
    "scimlop_factorize_bytes": [_INT, _I64, C.POINTER(_SZ), C.POINTER(_SZ)],
    "scimlop_factorize": [_P, _INT, _I64, _P, _I64, _P, _SZ, _P, _I64, _P, _SZ, C.POINTER(C.c_double), _P],
    "scimlop_kron_ldiv": [_P, _INT, _INT, C.POINTER(_P), _I64P, _I32P, _P, _P, _I64, _P, _SZ, _P],
    "scimlop_kron_workspace_bytes": [_INT, _INT, _I64P, _I32P, _I64, C.POINTER(_SZ)],
    "scimlop_kron_mul": [_P, _INT, _INT, _INT, C.POINTER(_P), _I64P, _I64P, _I32P, _P, _I64, _P, _I64, _I64, _P,
                         _P, _P, _SZ, _P],
    "scimlop_outer_mul_fast": [_P, _INT, _INT, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _P, _P, _P],
    "scimlop_kronsum_mul": [_P, _INT, _INT, _P, _I64, _I64, _P, _I64, _I64, _P, _I64, _P, _I64, _I64, _P, _P, _P],
    "scimlop_kronsum_banded_mul": [_P, _INT, _P, _INT, _I64, _P, _INT, _I64, _P, _I64, _P, _I64, _I64, _P, _P, _P],
    "scimlop_tree_workspace_bytes": [_INT, C.POINTER(Term), _INT, _I64, C.POINTER(_SZ)],
    "scimlop_tree_mul": [_P, _INT, _INT, C.POINTER(Term), _INT, _P, _I64, _P, _I64, _I64, _P, _P, _P, _SZ, _P],
    "scimlop_blockdiag_mul": [_P, _INT, _INT, _INT, C.POINTER(Factor), _P, _I64, _P, _I64, _I64, _P, _P, _P],
    "scimlop_kron_mul_host": [_P, _INT, _INT, _INT, C.POINTER(_P), _I64P, _I64P, _I32P, _P, _I64, _P, _I64, _I64,
                              _P, _P, _I64],
    "scimlop_col_dot": [_P, _INT, _I64, _I64, _P, _I64, _P, _I64, _P, _P],
    "scimlop_col_axpby_dot": [_P, _INT, _I64, _I64, _P, _I64, C.c_double, _P, _P, _P, _I64, C.c_double, _P, _P, _P, _P],
    "scimlop_shard_columns": [_I64, _INT, _INT, _I64P, _I64P],
    "scimlop_comm_unique_id": [_P],
    "scimlop_comm_init": [_P, _INT, _INT, _P],
    "scimlop_comm_destroy": [_P],
    "scimlop_allgather_cols": [_P, _INT, _P, _I64, _I64, _P],
    "scimlop_peer_export": [_P, _P, _P, _I64P],
    "scimlop_peer_open": [_P, _P, _I64, C.POINTER(_P)],
    "scimlop_peer_close": [_P, _P, _I64],
    "scimlop_kron_mul_peers": [_P, _INT, _INT, _INT, C.POINTER(_P), _I64P, _I64P, _I32P, _P, _I64, _P, C.POINTER(_P), _INT,
                               _P, _P, _SZ, _P],
    "scimlop_comm_barrier": [_P, _P],
    "scimlop_mc_padded_bytes": [_P, _SZ, _INT, C.POINTER(_SZ)],
    "scimlop_mc_open": [_P, _SZ, _INT, _INT, _INT, C.POINTER(_INT), C.POINTER(_P)],
    "scimlop_mc_bind": [_P, _P, C.POINTER(_P), C.POINTER(_P)],
    "scimlop_mc_free": [_P, _P],
    "scimlop_kron_mul_mc": [_P, _INT, _INT, _INT, C.POINTER(_P), _I64P, _I64P, _I32P, _P, _I64, _P, _P, _P, _P, _SZ, _P],
    "scimlop_kron_mul_allgather": [_P, _INT, _INT, _INT, C.POINTER(_P), _I64P, _I64P, _I32P, _P, _P, _I64, _P, _I64, _P, _SZ, _P],
}

class CsrFactor(C.Structure):
    """scimlop_csr_factor (include/scimlop_b200.h): a sparse Kronecker factor, device pointers in a host struct."""
    _fields_ = [("rowptr", C.c_void_p), ("colidx", C.c_void_p), ("values", C.c_void_p), ("index_base", C.c_int32),
                ("reserved", C.c_int32)]


_lib = None


def load():
    """dlopen the shared library and declare every entry point of the header. Raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library disagree
            fn.argtypes = args
            fn.restype = C.c_char_p if name == "scimlop_status_string" else C.c_int
        _lib = lib
    return _lib


def status_string(status):
    return load().scimlop_status_string(int(status)).decode()


class Handle:
    """Per-device library handle (scimlop_create / scimlop_destroy)."""

    def __init__(self, device=0):
        self.lib = load()
        self._h = _P()
        rc = self.lib.scimlop_create(C.byref(self._h), int(device))
        if rc != OK:
            raise ScimlopError(rc, "scimlop_create")
        self.device = int(device)

    @property
    def ptr(self):
        return self._h

    def check(self, rc, where):
        if rc != OK:
            buf = C.create_string_buffer(512)
            self.lib.scimlop_last_error(self._h, buf, 512)
            raise ScimlopError(rc, where, buf.value.decode(errors="replace"))

    def set_option(self, option, value):
        """scimlop_set_option: OPT_TAIL_SPLIT = 1, OPT_PEER_PIPELINE = 2 (include/scimlop_b200.h)."""
        self.check(self.lib.scimlop_set_option(self._h, int(option), int(value)), "scimlop_set_option")

    def launch_count(self):
        n = C.c_int64(0)
        self.check(self.lib.scimlop_launch_count(self._h, C.byref(n)), "scimlop_launch_count")
        return n.value

    def close(self):
        if self._h:
            self.lib.scimlop_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def handle(device=None):
    """The process-wide handle of a CUDA device (default: torch's current device)."""
    if device is None:
        import torch
        device = torch.cuda.current_device()
    device = int(device)
    if device not in _handles:
        _handles[device] = Handle(device)
    return _handles[device]
