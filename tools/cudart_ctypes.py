"""Minimal libcudart binding (ctypes) for the torch-free GPU harnesses in tools/: a fresh gpurun box spends up to a minute
importing torch, these start in a second.  Device memory allocated here is used by libvtb200.so (which links its own
static runtime) through the shared primary context."""
import ctypes as C

import numpy as np

rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]
rt.cudaGetErrorString.restype = C.c_char_p
rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
rt.cudaEventSynchronize.argtypes = [C.c_void_p]
rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
H2D, D2H, D2D = 1, 2, 3


def ck(rc, what="cuda"):
    if rc != 0:
        raise SystemExit(f"FAIL {what}: cuda error {rc} {rt.cudaGetErrorString(rc).decode()}")


def init(device=0):
    ck(rt.cudaSetDevice(device), "cudaSetDevice")
    ck(rt.cudaFree(None), "context")


class Buf:
    """A device allocation with a numpy-ish shape/dtype tag."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = C.c_void_p()
        ck(rt.cudaMalloc(C.byref(self.ptr), max(self.nbytes, 1)), "cudaMalloc")

    @property
    def addr(self):
        return self.ptr.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, self.dtype)
        assert arr.nbytes == self.nbytes, (arr.shape, self.shape)
        ck(rt.cudaMemcpy(self.ptr, arr.ctypes.data, self.nbytes, H2D), "H2D")
        return self

    def download(self):
        out = np.empty(self.shape, self.dtype)
        ck(rt.cudaMemcpy(out.ctypes.data, self.ptr, self.nbytes, D2H), "D2H")
        return out

    def zero(self):
        ck(rt.cudaMemset(self.ptr, 0, self.nbytes), "memset")
        return self

    def fill_from(self, seed_buf):
        """Tile the contents of `seed_buf` (device) over this buffer with device-to-device copies."""
        off = 0
        while off < self.nbytes:
            n = min(seed_buf.nbytes, self.nbytes - off)
            ck(rt.cudaMemcpy(C.c_void_p(self.addr + off), seed_buf.ptr, n, D2D), "D2D")
            off += n
        return self

    def free(self):
        rt.cudaFree(self.ptr)


def to_bf16_bits(x):
    """float32 array -> uint16 bf16 bit patterns (round to nearest even)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def from_bf16_bits(b):
    return (np.asarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32)


class Timer:
    def __init__(self):
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        ck(rt.cudaEventCreate(C.byref(self.e0)))
        ck(rt.cudaEventCreate(C.byref(self.e1)))

    def time(self, fn, n=10, warmup=3):
        """Mean microseconds per call of `fn` on the default stream."""
        for _ in range(warmup):
            fn()
        ck(rt.cudaDeviceSynchronize(), "sync")
        ck(rt.cudaEventRecord(self.e0, None))
        for _ in range(n):
            fn()
        ck(rt.cudaEventRecord(self.e1, None))
        ck(rt.cudaEventSynchronize(self.e1), "sync")
        ms = C.c_float()
        ck(rt.cudaEventElapsedTime(C.byref(ms), self.e0, self.e1))
        return ms.value / n * 1e3
