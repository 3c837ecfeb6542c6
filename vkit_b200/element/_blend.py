"""Launches the device blend that stands in for `fill_np_array`
(vkit/element/opt.py:118-209): masked assign, keep-max / keep-min merge and the float32
alpha blend whose result is TRUNCATED when cast back to uint8.
"""
import ctypes
from typing import Optional, Tuple, Union

import numpy as np

from .. import _native
from .. import device as dv


def _as_device(arr, dtype=None):
    """ndarray / tensor -> contiguous CUDA tensor (optionally cast like ndarray.astype)."""
    t = dv.require_cuda()
    if isinstance(arr, np.ndarray):
        if dtype is not None and arr.dtype != dtype:
            arr = arr.astype(dtype)
        return dv.to_device(arr)
    if not arr.is_cuda:
        raise _native.NativeError('blend operands must be NumPy arrays or CUDA tensors '
                                  '(a CPU tensor was given)')
    if dtype is not None:
        arr = arr.to(dv._torch_dtype(dtype))
    # the kernel addresses operands as dense rows (pitch = width): sliced / strided views are
    # copied.  (Views produced by Box.extract_* are copies already -- unlike the reference's NumPy
    # views, filling an extracted element does not write through to its parent.)
    return arr.contiguous()


def fill_region(
    dst,  # CUDA tensor, HxW or HxWxC, uint8 or float32 (modified in place)
    box: Tuple[int, int, int, int],  # up, down, left, right (inclusive) within dst
    value,
    np_mask=None,
    alpha: Union[float, np.ndarray, object] = 1.0,
    keep_max_value: bool = False,
    keep_min_value: bool = False,
):
    lib = _native.lib()
    up, down, left, right = box
    box_h, box_w = down - up + 1, right - left + 1
    full_h, full_w = int(dst.shape[0]), int(dst.shape[1])
    channels = 1 if dst.dim() == 2 else int(dst.shape[2])
    dst_f32 = str(dst.dtype) == 'torch.float32'
    np_dtype = np.float32 if dst_f32 else np.uint8

    item = _native.BlendItem()
    item.dst = dst.data_ptr()
    item.dst_f32 = int(dst_f32)
    item.channels = channels
    item.dst_w = full_w
    item.box_y, item.box_x, item.box_h, item.box_w = up, left, box_h, box_w
    keep = []

    # ---- value (prep_value, opt.py:96-115; Box.prep_mat_and_value, box.py:277-296) ----------
    if isinstance(value, np.ndarray) or dv.is_tensor(value):
        vshape = tuple(value.shape)
        if vshape[:2] == (box_h, box_w):
            origin = 0
            pitch = box_w
        elif vshape[:2] == (full_h, full_w):
            origin = up * full_w + left
            pitch = full_w
        else:
            raise RuntimeError('value is np.ndarray but shape is not matched.')
        vch = 1 if len(vshape) == 2 else vshape[2]
        if vch != channels:
            raise RuntimeError('value is np.ndarray but shape is not matched.')
        vdev = _as_device(value, np_dtype)
        keep.append(vdev)
        item.value_arr = vdev.data_ptr() + origin * channels * vdev.element_size()
        item.value_pitch = pitch
    else:
        if isinstance(value, tuple):
            if channels > 1 and len(value) != channels:
                raise RuntimeError('value is tuple but len(value) != num_channels.')
            consts = list(value)
        else:
            consts = [value] * max(channels, 1)
        # np.full_like(mat, value): cast to the destination dtype
        consts = np.asarray(consts).astype(np_dtype)
        for i in range(min(4, len(consts))):
            item.value_const[i] = float(consts[i])
        item.value_arr = None

    # ---- alpha ---------------------------------------------------------------------------
    if isinstance(alpha, np.ndarray) or dv.is_tensor(alpha):
        if tuple(alpha.shape) != (box_h, box_w):
            raise RuntimeError('alpha array shape does not match the box.')
        adev = _as_device(alpha, np.float32)
        keep.append(adev)
        item.alpha_arr = adev.data_ptr()
        item.alpha_pitch = box_w
        item.alpha = 1.0
    else:
        alpha = float(alpha)
        if alpha < 0.0 or alpha > 1.0:
            raise RuntimeError(f'alpha={alpha} is invalid.')
        if alpha == 0.0:
            return
        item.alpha_arr = None
        item.alpha = alpha

    # ---- mask ----------------------------------------------------------------------------
    if np_mask is not None:
        if tuple(np_mask.shape) != (box_h, box_w):
            raise RuntimeError('mask shape does not match the box.')
        if isinstance(np_mask, np.ndarray):
            np_mask = np_mask.astype(np.uint8) if np_mask.dtype != np.uint8 else np_mask
        mdev = _as_device(np_mask, np.uint8)
        keep.append(mdev)
        item.mask = mdev.data_ptr()
        item.mask_pitch = box_w
    else:
        item.mask = None

    assert not (keep_max_value and keep_min_value)
    item.keep_mode = 1 if keep_max_value else (2 if keep_min_value else 0)
    if box_h <= 0 or box_w <= 0:
        return
    _native.check(lib.vkb_blend_fill(ctypes.byref(item), dv.stream_ptr()), 'vkb_blend_fill')
    del keep
