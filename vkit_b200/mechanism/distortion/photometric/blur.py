"""Blur ops (vkit/mechanism/distortion/photometric/blur.py).  gaussian_blur runs as a separable
8.8 fixed-point shared-memory stencil that reproduces cv.GaussianBlur on uint8 bit-exactly."""
import ctypes
from typing import Tuple, Any, Mapping, Optional

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import Image

from ..interface import Distortion, DistortionConfig, DistortionNopState
from .opt import to_original_image, to_rgb_image


def _estimate_gaussian_kernel_size(sigma: float):
    kernel_size = max(3, round(3 * sigma) + 1)
    if kernel_size % 2 == 0:
        kernel_size += 1
    return kernel_size


def gaussian_kernel_u8(ksize: int, sigma: float):
    """Integer taps (sum 256) cv2 derives for uint8 images: float Gaussian, normalised, then
    rounded to 8.8 fixed point from the outside in with error diffusion, centre = remainder."""
    r = ksize // 2
    xs = np.arange(ksize, dtype=np.float64) - r
    k = np.exp(-(xs * xs) / (2.0 * sigma * sigma))
    k = k / k.sum()
    taps = [0] * ksize
    err = 0.0
    for i in range(r):
        adj = k[i] * 256.0 + err
        v = float(np.rint(adj))
        err = adj - v
        taps[i] = taps[ksize - 1 - i] = int(v)
    taps[r] = 256 - sum(taps)
    return taps


def gaussian_kernels_u8(sigmas) -> Tuple[np.ndarray, np.ndarray]:
    """gaussian_kernel_u8 for many pages at once (batched chains build one tap row per page):
    returns (ksizes, taps) with taps[i, :ksizes[i]] the integer taps of page i.  Same float64
    operations per element as the scalar form, vectorised over the pages of each kernel size
    (tests/test_host_logic.py compares the two on random sigmas)."""
    sigmas = np.asarray(sigmas, dtype=np.float64)
    ksizes = np.asarray([_estimate_gaussian_kernel_size(float(s)) for s in sigmas], dtype=np.int64)
    taps = np.zeros((sigmas.shape[0], 17), dtype=np.int32)
    for ksize in np.unique(ksizes):
        ksize = int(ksize)
        if ksize > 17:
            continue  # the caller rejects these
        sel = np.flatnonzero(ksizes == ksize)
        r = ksize // 2
        xs = np.arange(ksize, dtype=np.float64) - r
        k = np.exp(-(xs * xs)[None, :] / (2.0 * sigmas[sel] * sigmas[sel])[:, None])
        k = k / k.sum(axis=1)[:, None]
        out = np.zeros((sel.shape[0], ksize), dtype=np.int64)
        err = np.zeros(sel.shape[0], dtype=np.float64)
        for i in range(r):
            adj = k[:, i] * 256.0 + err
            v = np.rint(adj)
            err = adj - v
            out[:, i] = out[:, ksize - 1 - i] = v.astype(np.int64)
        out[:, r] = 256 - out.sum(axis=1)
        taps[sel, :ksize] = out
    return ksizes, taps


def gaussian_blur_device(image: Image, sigma: float) -> Image:
    ksize = _estimate_gaussian_kernel_size(sigma)
    if ksize > 17:
        raise NotImplementedError('gaussian_blur kernels wider than 17 taps are not provided')
    taps = gaussian_kernel_u8(ksize, sigma)
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    arr = (ctypes.c_int32 * ksize)(*taps)
    nv.check(nv.lib().vkb_gaussian_blur_u8(dv.ptr(src), dv.ptr(dst), image.height, image.width,
                                           image.num_channels or 1, arr, ksize, dv.stream_ptr()),
             'vkb_gaussian_blur_u8')
    return attrs.evolve(image, mat=dst)


@attrs.define
class GaussianBlurConfig(DistortionConfig):
    sigma: float


def gaussian_blur_image(config: GaussianBlurConfig, state, image: Image,
                        rng: Optional[RandomGenerator]):
    mode = image.mode
    image = to_rgb_image(image, mode)
    image = gaussian_blur_device(image, config.sigma)
    return to_original_image(image, mode)


gaussian_blur = Distortion(config_cls=GaussianBlurConfig,
                           state_cls=DistortionNopState[GaussianBlurConfig],
                           func_image=gaussian_blur_image)


def _next_row(name):
    def func(config, state, image, rng):
        raise NotImplementedError(
            f'{name} is a "next" row of the scope table (SURVEY.md section 8f) and has no device '
            'kernel yet; disable it via RandomDistortionFactoryConfig.disabled_policy_names.')
    return func


# ---- defocus / motion blur: a small float32 kernel built on the host, cv.filter2D on the device ---
def _get_anti_aliasing_kernel_size_and_padding(anti_aliasing_sigma: float):
    kernel_size = _estimate_gaussian_kernel_size(anti_aliasing_sigma)
    return kernel_size, kernel_size // 2 * 2


def gaussian_taps_f32(ksize: int, sigma: float) -> np.ndarray:
    """cv.getGaussianKernel(ksize, sigma, CV_32F): double taps, normalised, cast to float32."""
    xs = np.arange(ksize, dtype=np.float64) - ksize // 2
    taps = np.exp(-(xs * xs) / (2.0 * sigma * sigma))
    return (taps / taps.sum()).astype(np.float32)


def _reflect101(index: int, size: int) -> int:
    if size == 1:
        return 0
    while index < 0 or index >= size:
        index = -index if index < 0 else 2 * (size - 1) - index
    return index


def gaussian_blur_f32(mat: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv.GaussianBlur on a small float32 array (BORDER_REFLECT_101), separable, float32 sums in
    the symmetric order centre*k0 + (left + right)*k1 + ...  Matches cv2 to 1 ulp of float32 (cv2's
    own summation order depends on its SIMD / IPP backend); the array is a blur kernel whose taps
    feed a uint8-rounded convolution."""
    taps = gaussian_taps_f32(ksize, sigma)
    r = ksize // 2
    f32 = np.float32

    def one_axis(src: np.ndarray, axis: int) -> np.ndarray:
        n = src.shape[axis]
        out = np.zeros_like(src)
        for pos in range(n):
            centre = np.take(src, pos, axis=axis)
            acc = (centre * taps[r]).astype(f32)
            for d in range(1, r + 1):
                pair = (np.take(src, _reflect101(pos - d, n), axis=axis)
                        + np.take(src, _reflect101(pos + d, n), axis=axis)).astype(f32)
                acc = (acc + (pair * taps[r + d]).astype(f32)).astype(f32)
            if axis == 0:
                out[pos, :] = acc
            else:
                out[:, pos] = acc
        return out

    return one_axis(one_axis(np.asarray(mat, dtype=f32), 1), 0)


def rotation_matrix_2d(center, angle: float, scale: float) -> np.ndarray:
    """cv.getRotationMatrix2D (double)."""
    import math
    angle = angle * (np.pi / 180)
    alpha = scale * math.cos(angle)
    beta = scale * math.sin(angle)
    cx, cy = center
    return np.asarray([[alpha, beta, (1 - alpha) * cx - beta * cy],
                       [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def filter2d_device(image: Image, kernel: np.ndarray) -> Image:
    """cv.filter2D(image.mat, -1, kernel) on the device (uint8, BORDER_REFLECT_101)."""
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    kh, kw = kernel.shape
    if kh > 63 or kw > 63:
        raise NotImplementedError('filter kernels larger than 63 x 63 are not provided')
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    taps = dv.to_device(kernel)
    nv.check(nv.lib().vkb_filter2d_u8(dv.ptr(src), dv.ptr(dst), image.height, image.width,
                                      image.num_channels or 1, dv.ptr(taps), kh, kw,
                                      dv.stream_ptr()), 'vkb_filter2d_u8')
    return attrs.evolve(image, mat=dst)


@attrs.define
class DefocusBlurConfig(DistortionConfig):
    radius: int
    anti_aliasing_sigma: float = 0.5


def defocus_blur_kernel(radius: int, anti_aliasing_sigma: float) -> np.ndarray:
    # blur.py:91-112: disc of ones / its sum, then anti-aliased with a Gaussian
    assert 0 < radius
    aa_ksize, aa_padding = _get_anti_aliasing_kernel_size_and_padding(anti_aliasing_sigma)
    kernel_size = 2 * radius + 1 + aa_padding
    begin = -(kernel_size // 2)
    coords = np.arange(begin, begin + kernel_size)
    x, y = np.meshgrid(coords, coords)
    kernel = ((x**2 + y**2) <= radius**2).astype(np.float32)
    kernel /= kernel.sum()
    return gaussian_blur_f32(kernel, aa_ksize, anti_aliasing_sigma)


def defocus_blur_image(config: DefocusBlurConfig, state, image: Image,
                       rng: Optional[RandomGenerator]):
    kernel = defocus_blur_kernel(config.radius, config.anti_aliasing_sigma)
    mode = image.mode
    image = to_rgb_image(image, mode)
    image = filter2d_device(image, kernel)
    return to_original_image(image, mode)


defocus_blur = Distortion(config_cls=DefocusBlurConfig,
                          state_cls=DistortionNopState[DefocusBlurConfig],
                          func_image=defocus_blur_image)


@attrs.define
class MotionBlurConfig(DistortionConfig):
    radius: int
    angle: int
    anti_aliasing_sigma: float = 0.5


def motion_blur_kernel(radius: int, angle: int, anti_aliasing_sigma: float) -> np.ndarray:
    # blur.py:145-177: a horizontal line of ones, rotated with cv.warpAffine (float32, bilinear;
    # here through the device warp that reproduces it bit for bit), / its sum, anti-aliased
    from vkit_b200.element import ScoreMap
    from ..geometric.affine import warp_planes
    kernel_size = 2 * radius + 1
    aa_ksize, aa_padding = _get_anti_aliasing_kernel_size_and_padding(anti_aliasing_sigma)
    half = aa_padding // 2
    center = radius + half
    left = half
    right = left + kernel_size - 1
    kernel_size += aa_padding
    kernel = np.zeros((kernel_size, kernel_size), dtype=np.float32)
    kernel[center, left:right + 1] = 1.0
    trans_mat = rotation_matrix_2d((center, center), 360 - (angle % 360), 1.0)
    _, _, rotated = warp_planes(trans_mat, (kernel_size, kernel_size),
                                score_map=ScoreMap(mat=kernel, is_prob=False))
    kernel = dv.to_host(rotated).astype(np.float32).reshape(kernel_size, kernel_size).copy()
    kernel /= kernel.sum()
    return gaussian_blur_f32(kernel, aa_ksize, anti_aliasing_sigma)


def motion_blur_image(config: MotionBlurConfig, state, image: Image,
                      rng: Optional[RandomGenerator]):
    kernel = motion_blur_kernel(config.radius, config.angle, config.anti_aliasing_sigma)
    mode = image.mode
    image = to_rgb_image(image, mode)
    image = filter2d_device(image, kernel)
    return to_original_image(image, mode)


motion_blur = Distortion(config_cls=MotionBlurConfig,
                         state_cls=DistortionNopState[MotionBlurConfig],
                         func_image=motion_blur_image)


@attrs.define
class GlassBlurConfig(DistortionConfig):
    sigma: float
    delta: int = 1
    loop: int = 5

    _rng_state: Optional[Mapping[str, Any]] = None

    @property
    def supports_rng_state(self) -> bool:
        return True

    @property
    def rng_state(self) -> Optional[Mapping[str, Any]]:
        return self._rng_state

    @rng_state.setter
    def rng_state(self, val: Mapping[str, Any]):
        self._rng_state = val


def glass_swap_maps(shape, delta: int, loop: int, rng: RandomGenerator):
    """The pixel permutation of glass_blur as two int32 index maps (blur.py:232-262): `loop` rounds
    in which every (2*delta+1)-th pixel trades places with a random neighbour OF THE PIXEL THAT
    CURRENTLY SITS THERE.  Drawn on the host from the caller's generator -- it is the random field
    of the op; duplicates among the targets resolve like NumPy's fancy assignment (the last writer
    in C order wins).  The centres form a regular sub-grid, so they are strided views; only the
    ~(H / period) x (W / period) targets go through fancy indexing."""
    height, width = shape
    pos_y = np.broadcast_to(np.arange(height, dtype=np.int32)[:, None], (height, width)).copy()
    pos_x = np.broadcast_to(np.arange(width, dtype=np.int32)[None, :], (height, width)).copy()
    period = 2 * delta + 1
    for _ in range(loop):
        row0 = int(rng.integers(0, period))
        col0 = int(rng.integers(0, period))
        centres = (slice(row0, height - delta, period), slice(col0, width - delta, period))
        grid_shape = pos_y[centres].shape
        shift_y = rng.integers(-delta, delta + 1, grid_shape)
        shift_x = rng.integers(-delta, delta + 1, grid_shape)
        target_y = np.clip(pos_y[centres] + shift_y, 0, height - 1)
        target_x = np.clip(pos_x[centres] + shift_x, 0, width - 1)
        for pos in (pos_y, pos_x):
            at_centre, at_target = pos[centres].copy(), pos[target_y, target_x]
            pos[centres] = at_target
            pos[target_y, target_x] = at_centre
    return pos_y, pos_x


def glass_swap_maps_device(shape, delta: int, loop: int, rng: RandomGenerator):
    """glass_swap_maps on the device, from the caller's generator stream: the host draws the two
    offsets of every round, the device regenerates the round's bounded integers (`vkb_glass_round`)
    and the host moves its PCG64 past them (whole 64-bit outputs by `advance`, the last one drawn
    for the half it may leave pending).  Returns (pos_y, pos_x) as CUDA int32 tensors, or None when
    the generator is not a PCG64, the page is too large for the round keys, or the device saw a
    draw NumPy would have rejected (2^-32 per draw) -- the generator is then back in its initial
    state and the caller draws the maps on the host."""
    bit_generator = rng.bit_generator
    height, width = shape
    period = 2 * delta + 1
    if (type(bit_generator).__name__ != 'PCG64' or delta < 1 or loop > 100
            or -(-height // period) * -(-width // period) >= (1 << 24)):
        return None
    initial = bit_generator.state
    lib = nv.lib()
    n_max = -(-height // period) * -(-width // period)
    pos_y = dv.empty((height, width), np.int32)
    pos_x = dv.empty((height, width), np.int32)
    owner = dv.empty((height, width), np.int32)
    target = dv.empty((max(1, n_max),), np.int32)
    at_centre = dv.empty((max(1, 2 * n_max),), np.int32)
    at_target = dv.empty((max(1, 2 * n_max),), np.int32)
    flag = dv.empty((1,), np.int32)
    nv.check(lib.vkb_glass_init(dv.ptr(pos_y), dv.ptr(pos_x), dv.ptr(owner), height, width,
                                dv.ptr(flag), dv.stream_ptr()), 'vkb_glass_init')
    mask64 = (1 << 64) - 1
    for k in range(loop):
        row0 = int(rng.integers(0, period))
        col0 = int(rng.integers(0, period))
        n = len(range(row0, height - delta, period)) * len(range(col0, width - delta, period))
        state = bit_generator.state
        s128, inc128 = state['state']['state'], state['state']['inc']
        pending = int(state['has_uint32'])
        nv.check(lib.vkb_glass_round(
            dv.ptr(pos_y), dv.ptr(pos_x), dv.ptr(owner), height, width, row0, col0, delta, k,
            s128 >> 64, s128 & mask64, inc128 >> 64, inc128 & mask64, pending,
            int(state['uinteger']) if pending else 0, dv.ptr(target), dv.ptr(at_centre),
            dv.ptr(at_target), dv.ptr(flag), dv.stream_ptr()), 'vkb_glass_round')
        # the generator moves past the round's 2 * n halves
        fresh = 2 * n - pending  # halves taken from new 64-bit outputs
        if fresh > 0:
            outputs = (fresh + 1) // 2
            bit_generator.advance(outputs - 1)  # (resets the pending half)
            last = int(bit_generator.random_raw())
            after = bit_generator.state
            after['has_uint32'] = fresh & 1  # an odd number of halves leaves the high half pending
            after['uinteger'] = last >> 32   # (NumPy keeps the value after it has been used, too)
            bit_generator.state = after
        elif 2 * n > 0:  # the single half of the round was the pending one
            after = bit_generator.state
            after['has_uint32'] = 0
            bit_generator.state = after
    if int(dv.to_host(flag)[0]):
        bit_generator.state = initial
        return None
    return pos_y, pos_x


def glass_blur_image(config: GlassBlurConfig, state, image: Image,
                     rng: Optional[RandomGenerator]):
    mode = image.mode
    image = to_rgb_image(image, mode)
    image = gaussian_blur_device(image, config.sigma)
    assert rng is not None
    maps = glass_swap_maps_device(image.shape, config.delta, config.loop, rng)
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    if maps is not None:
        py, px = maps
    else:  # a generator whose stream cannot be split, or a rejected draw: the maps come from the host
        pos_y, pos_x = glass_swap_maps(image.shape, config.delta, config.loop, rng)
        py = dv.to_device(np.ascontiguousarray(pos_y, dtype=np.int32))
        px = dv.to_device(np.ascontiguousarray(pos_x, dtype=np.int32))
    nv.check(nv.lib().vkb_gather_pixels_u8(dv.ptr(src), dv.ptr(dst), image.height, image.width,
                                           image.num_channels or 1, dv.ptr(py), dv.ptr(px),
                                           dv.stream_ptr()), 'vkb_gather_pixels_u8')
    image = attrs.evolve(image, mat=dst)
    return to_original_image(image, mode)


glass_blur = Distortion(config_cls=GlassBlurConfig,
                        state_cls=DistortionNopState[GlassBlurConfig],
                        func_image=glass_blur_image)


@attrs.define
class ZoomInBlurConfig(DistortionConfig):
    ratio: float = 0.1
    step: float = 0.01
    alpha: float = 0.5


def zoom_in_blur_image(config: ZoomInBlurConfig, state, image: Image,
                       rng: Optional[RandomGenerator]):
    # blur.py:285-323: the page + its enlargements by 1+step .. 1+ratio (cubic, centre crops),
    # averaged and blended; one launch samples every enlargement on the fly
    mode = image.mode
    image = to_rgb_image(image, mode)
    height, width = image.shape
    levels = []
    for ratio in np.arange(1 + config.step, 1 + config.ratio + config.step, config.step):
        resized_height = round(height * ratio)
        resized_width = round(width * ratio)
        levels.append((1.0 / (resized_width / width), 1.0 / (resized_height / height),
                       (resized_height - height) // 2, (resized_width - width) // 2))
    rec = np.zeros(len(levels), dtype=nv.ZOOM_LEVEL_DTYPE)
    for i, (sx, sy, up, left) in enumerate(levels):
        rec[i] = (sx, sy, up, left)
    src = image.dev
    dst = dv.empty(tuple(src.shape), np.uint8)
    rec_dev = dv.upload_structs(rec) if len(levels) else None
    nv.check(nv.lib().vkb_zoom_in_blur_u8(dv.ptr(src), dv.ptr(dst), height, width,
                                          image.num_channels or 1, dv.ptr(rec_dev), len(levels),
                                          float(config.alpha), dv.stream_ptr()),
             'vkb_zoom_in_blur_u8')
    image = attrs.evolve(image, mat=dst)
    return to_original_image(image, mode)


zoom_in_blur = Distortion(config_cls=ZoomInBlurConfig,
                          state_cls=DistortionNopState[ZoomInBlurConfig],
                          func_image=zoom_in_blur_image)
