"""Dual host/device storage behind `Image.mat`, `Mask.mat`, `ScoreMap.mat`.

The reference containers wrap a read-only `np.ndarray` (vkit/element/image.py:254-255).  Here the
pixels may live on the GPU (a torch CUDA tensor), on the host, or both:

  * `.mat`  -> NumPy array for reference-style code; downloaded lazily, read-only by default;
  * `.dev`  -> CUDA tensor for the kernels; uploaded lazily.

A chain of distortions therefore never leaves HBM; a device->host copy happens only when
somebody reads `.mat`.  The attrs field `_mat` (init alias `mat`) holds whichever object is
authoritative, `_alt` caches the counterpart, so `attrs.evolve(element, box=...)` keeps working
for device-resident elements.
"""
from contextlib import ContextDecorator

import numpy as np

from .. import device as dv


class WritableContext(ContextDecorator):
    """`with element.writable_context:` -- host-side mutation window (image.py:158-180)."""

    def __init__(self, element):
        super().__init__()
        self.element = element

    def __enter__(self):
        mat = self.element.mat  # materialise on host
        try:
            mat.flags.writeable = True
        except ValueError:
            mat = np.array(mat)  # copy on write for views that cannot be made writable
        # the host array becomes authoritative, the device copy is stale from now on
        object.__setattr__(self.element, '_mat', mat)
        object.__setattr__(self.element, '_alt', None)

    def __exit__(self, *exc):
        self.element._mat.flags.writeable = False
        self.element._after_host_write()


class DualStorage:
    """Mixin; the concrete attrs class declares `_mat` (alias `mat`) and `_alt`."""

    def _adopt(self, mat):
        """Called from __attrs_post_init__ with whatever was passed as `mat`."""
        if isinstance(mat, np.ndarray):
            mat.flags.writeable = False
        elif dv.is_tensor(mat):
            if not mat.is_cuda:
                raise TypeError('tensor-backed elements must live on a CUDA device')
            if not mat.is_contiguous():
                mat = mat.contiguous()
        else:
            raise TypeError(f'mat must be np.ndarray or a CUDA tensor, got {type(mat)}')
        object.__setattr__(self, '_mat', mat)
        object.__setattr__(self, '_alt', None)

    # -- introspection without forcing a transfer -------------------------------------------
    @property
    def on_device(self):
        return not isinstance(self._mat, np.ndarray)

    @property
    def storage(self):
        """The authoritative object (ndarray or CUDA tensor), e.g. to pass to attrs.evolve."""
        return self._mat

    @property
    def mat_shape(self):
        return tuple(self._mat.shape)

    @property
    def mat_ndim(self):
        return len(self._mat.shape)

    @property
    def mat_dtype(self):
        if isinstance(self._mat, np.ndarray):
            return self._mat.dtype
        return np.dtype(str(self._mat.dtype).replace('torch.', ''))

    # -- the two views ----------------------------------------------------------------------
    @property
    def mat(self) -> np.ndarray:
        if isinstance(self._mat, np.ndarray):
            return self._mat
        if self._alt is None:
            host = dv.to_host(self._mat)
            host.flags.writeable = False
            object.__setattr__(self, '_alt', host)
        return self._alt

    @property
    def dev(self):
        if not isinstance(self._mat, np.ndarray):
            return self._mat
        if self._alt is None:
            object.__setattr__(self, '_alt', dv.to_device(self._mat))
        return self._alt

    def _after_host_write(self):
        pass

    def _after_device_write(self):
        """A kernel wrote into `self.dev`: the device tensor is authoritative now."""
        if isinstance(self._mat, np.ndarray):
            object.__setattr__(self, '_mat', self._alt)
        object.__setattr__(self, '_alt', None)

    @property
    def height(self):
        return self._mat.shape[0]

    @property
    def width(self):
        return self._mat.shape[1]

    @property
    def writable_context(self):
        return WritableContext(self)

    def _clone_storage(self):
        """An independent copy of the pixels, staying where they are."""
        if self.on_device:
            return self._mat.clone()
        return self._mat.copy()

    def _crop_storage(self, up, down, left, right):
        if self.on_device:
            return self._mat[up:down + 1, left:right + 1].contiguous()
        return self._mat[up:down + 1, left:right + 1]
