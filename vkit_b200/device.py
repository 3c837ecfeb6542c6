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


# Pinned staging blocks, reused: cudaHostAlloc costs ~1 ms and torch's caching host allocator hands
# a block back only after the copy that used it has run, which under a deep queue of work means a
# fresh cudaHostAlloc per upload.  Blocks are kept per (device, size class) with the event of their
# last copy; a block is reused when that event has completed.
_PINNED = {}


def _pinned_block(nbytes: int):
    t = torch()
    size = 4096
    while size < nbytes:
        size *= 2
    pool = _PINNED.setdefault((t.cuda.current_device(), size), [])
    for entry in pool:
        if entry[1] is None or entry[1].query():
            return entry
    entry = [t.empty((size,), dtype=t.uint8, pin_memory=True), None]
    pool.append(entry)
    return entry


def upload_structs(records: np.ndarray):
    """Structured NumPy array (parameter blocks) -> uint8 device tensor.  Staged through pinned
    memory and copied asynchronously BY A KERNEL: a pageable source makes cudaMemcpyAsync wait for
    the stream (a host round trip in every batch), and any cudaMemcpyAsync queues on the copy
    engine behind the bulk page copies of the end-to-end pipeline."""
    t = require_cuda()
    raw = np.frombuffer(records.tobytes(), dtype=np.uint8)
    entry = _pinned_block(raw.size)
    entry[0].numpy()[:raw.size] = raw
    padded = (raw.size + 15) // 16 * 16
    out = t.empty((padded,), dtype=t.uint8, device=device())
    # a kernel copies the block (vkb_stage_params): the copy engine may be busy with page copies
    _native.check(_native.lib().vkb_stage_params(ptr(out), ctypes.c_void_p(entry[0].data_ptr()),
                                                 padded, stream_ptr()), 'vkb_stage_params')
    out = out[:raw.size]
    event = t.cuda.Event()
    event.record()
    entry[1] = event
    return out


class MirrorBlock:
    """A pinned host block that kernels mirror small results into (result shapes, layouts).  The
    block never goes back to torch's allocator: an abandoned plan's kernels may still be queued,
    and a block handed to somebody else meanwhile would be overwritten under them (parameter
    blocks corrupted by a late shape mirror: an illegal address in the remap).  `written()` is
    called right after the launch of the last kernel that writes the block; the pool reuses a
    block only when that point of the stream has been passed."""

    def __init__(self, tensor):
        self.tensor = tensor
        self.event = None

    def written(self):
        t = torch()
        self.event = t.cuda.Event()
        self.event.record()

    def data_ptr(self):
        return self.tensor.data_ptr()

    def numpy(self):
        return self.tensor.numpy()

    def release(self):
        t = torch()
        _PINNED.setdefault((t.cuda.current_device(), 'mirror', int(self.tensor.numel())), []).append(self)


def pinned_mirror(nbytes: int) -> MirrorBlock:
    t = require_cuda()
    size = 4096
    while size < nbytes:
        size *= 2
    pool = _PINNED.setdefault((t.cuda.current_device(), 'mirror', size), [])
    for i, block in enumerate(pool):
        if block.event is None or block.event.query():
            return pool.pop(i)
    return MirrorBlock(t.empty((size,), dtype=t.uint8, pin_memory=True))


def release_mirror(block: MirrorBlock):
    block.release()


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr()) if tensor is not None else ctypes.c_void_p(0)
