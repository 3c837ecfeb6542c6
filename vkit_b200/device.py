"""Device plumbing: PyTorch is the allocator / stream provider, nothing more.

Every function raises when CUDA is unavailable -- the distortion path has no CPU fallback.
"""
import ctypes

import numpy as np

from . import _native

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


_cuda_ok = False


def require_cuda():
    global _cuda_ok
    t = torch()
    if not _cuda_ok:  # torch.cuda.is_available() costs ~10 us per call: ask once
        if not t.cuda.is_available():
            raise _native.NativeError(
                'vkit_b200 needs a CUDA device (sm_100a); no CPU fallback exists for this path.')
        _native.lib()
        _cuda_ok = True
    return t


def is_tensor(obj):
    return _torch is not None and isinstance(obj, _torch.Tensor) or (
        type(obj).__module__.startswith('torch') and hasattr(obj, 'data_ptr'))


def device():
    t = require_cuda()
    return t.device('cuda', t.cuda.current_device())


def stream_ptr():
    t = torch()
    try:  # the raw handle without building a Stream object (~60 us -> ~1 us per call)
        return ctypes.c_void_p(t._C._cuda_getCurrentRawStream(t.cuda.current_device()))
    except AttributeError:
        return ctypes.c_void_p(t.cuda.current_stream().cuda_stream)


def to_device(array: np.ndarray):
    """Host array -> device tensor of the same shape/dtype (contiguous)."""
    t = require_cuda()
    array = np.ascontiguousarray(array)
    if not array.flags.writeable:
        # torch.from_numpy warns on read-only arrays; the tensor is only a copy source here.
        array = array.view()
        try:
            array.flags.writeable = True
        except ValueError:
            array = array.copy()
    return t.from_numpy(array).to(device(), non_blocking=False)


def to_host(tensor) -> np.ndarray:
    return tensor.detach().cpu().numpy()


def empty(shape, dtype):
    t = require_cuda()
    return t.empty(shape, dtype=_torch_dtype(dtype), device=device())


def zeros(shape, dtype):
    t = require_cuda()
    return t.zeros(shape, dtype=_torch_dtype(dtype), device=device())


def _torch_dtype(dtype):
    t = torch()
    dtype = np.dtype(dtype)
    table = {
        np.dtype(np.uint8): t.uint8, np.dtype(np.int16): t.int16, np.dtype(np.uint16): t.uint16,
        np.dtype(np.int32): t.int32, np.dtype(np.uint32): t.uint32, np.dtype(np.int64): t.int64,
        np.dtype(np.float32): t.float32, np.dtype(np.float64): t.float64,
        np.dtype(np.bool_): t.bool,
    }
    return table[dtype]


def upload_structs(records: np.ndarray):
    """Structured NumPy array (parameter blocks) -> uint8 device tensor.  Staged through pinned
    memory and copied asynchronously: a pageable source makes cudaMemcpyAsync wait for the stream,
    which would put a host round trip into every batch.  (The caching host allocator keeps the
    pinned block alive until the copy has run.)"""
    t = require_cuda()
    raw = np.frombuffer(records.tobytes(), dtype=np.uint8)
    host = t.empty((raw.size,), dtype=t.uint8, pin_memory=True)
    host.numpy()[:] = raw
    return host.to(device(), non_blocking=True)


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else ctypes.c_void_p(0)
