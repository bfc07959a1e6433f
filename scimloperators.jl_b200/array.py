"""Column-major device arrays.

Julia arrays are column-major; the C ABI takes column-major device pointers with explicit leading
dimensions.  `DeviceArray` is the Python stand-in for the `B200Array` type of the Julia extension
(INTEGRATION.md): a flat torch CUDA tensor (torch is only the allocator here) plus a column-major shape.
`reshape` is Julia's `reshape`: a view over the same storage (src/tensor.jl:558 relies on that)."""
from __future__ import annotations

import numpy as np
import torch

_TORCH = {np.dtype("float64"): torch.float64, np.dtype("float32"): torch.float32,
          np.dtype("int32"): torch.int32,   # int32: index arrays of sparse operators only
          np.dtype("complex64"): torch.complex64, np.dtype("complex128"): torch.complex128}   # interleaved, as Julia stores them


class DeviceArray:
    __slots__ = ("t", "shape")

    def __init__(self, flat: torch.Tensor, shape):
        shape = tuple(int(s) for s in shape)
        assert flat.dim() == 1 and flat.is_contiguous(), "DeviceArray wraps flat contiguous storage"
        assert flat.numel() == int(np.prod(shape)) if shape else True
        self.t = flat
        self.shape = shape

    # ---- constructors -------------------------------------------------------------------------------
    @staticmethod
    def from_numpy(a, device=None) -> "DeviceArray":
        a = np.asarray(a)
        if a.dtype not in _TORCH:
            raise TypeError(f"unsupported dtype {a.dtype}: the B200 backend takes Float64 / Float32 / ComplexF64 / ComplexF32")
        flat = np.ascontiguousarray(a.reshape(-1, order="F"))
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        return DeviceArray(torch.from_numpy(flat).to(dev), a.shape)

    @staticmethod
    def empty(shape, dtype, device=None) -> "DeviceArray":
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        shape = (shape,) if np.isscalar(shape) else tuple(shape)
        return DeviceArray(torch.empty(int(np.prod(shape)), dtype=_TORCH[np.dtype(dtype)], device=dev), shape)

    @staticmethod
    def zeros(shape, dtype, device=None) -> "DeviceArray":
        a = DeviceArray.empty(shape, dtype, device)
        a.t.zero_()
        return a

    @staticmethod
    def full(shape, value, dtype, device=None) -> "DeviceArray":
        a = DeviceArray.empty(shape, dtype, device)
        a.t.fill_(value)
        return a

    # ---- views ----------------------------------------------------------------------------------------
    def reshape(self, *shape) -> "DeviceArray":
        if len(shape) == 1 and not np.isscalar(shape[0]):
            shape = tuple(shape[0])
        return DeviceArray(self.t, shape)

    def cols(self, j0, j1) -> "DeviceArray":
        """Contiguous column slab [j0, j1) -- the unit of multi-GPU sharding."""
        assert self.ndim == 2
        n = self.shape[0]
        return DeviceArray(self.t[j0 * n:j1 * n], (n, j1 - j0))

    def copy(self) -> "DeviceArray":
        return DeviceArray(self.t.clone(), self.shape)

    # ---- properties -----------------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return {torch.float64: np.dtype("float64"), torch.float32: np.dtype("float32"), torch.int32: np.dtype("int32"),
                torch.complex64: np.dtype("complex64"), torch.complex128: np.dtype("complex128")}[self.t.dtype]

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @property
    def ld(self) -> int:
        return max(1, self.shape[0]) if self.shape else 1

    @property
    def ncols(self) -> int:
        return self.shape[1] if self.ndim == 2 else 1

    @property
    def device_index(self) -> int:
        return self.t.device.index

    def numpy(self) -> np.ndarray:
        return self.t.cpu().numpy().reshape(self.shape, order="F")

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device=cuda:{self.device_index})"
