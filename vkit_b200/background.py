"""Background synthesis on the device -- the step before text-layer compositing
(reference: ImageCombinerEngine, vkit/engine/image/combiner.py).

The reference fills a page-sized canvas with texture images along a rising skyline
(combiner.py:178-333): segments of the canvas width are kept in a heap keyed by their current
height, the lowest one receives the next randomly chosen texture, neighbouring segments that
reach the same height merge, and at the end a band around every pasted rectangle is replaced by
a Gaussian-blurred copy of the canvas to hide the seams.

Split here:

  * the skyline walk stays on the host -- a few dozen heap operations per page, and it fixes the
    order of the rng draws, so it is written to consume the generator exactly like the reference
    (`rng.choice` for the anchor and for every texture, `rng.random()` for the anchor-only and the
    rotate decisions, `rng.integers` for the initial cuts);
  * the pixels never touch the host: textures are uploaded once (plus their 90-degree rotation
    when the walk asks for it, through the `rotate` distortion like the reference), the walk
    yields a table of (texture, rectangle) records, and ONE launch of `vkb_background_compose`
    writes the finished canvas: paste, seam bands and the fixed-point blur fused.
"""
import ctypes
import heapq
import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from . import _native as nv
from . import device as dv
from .element import Image, ImageMode
from .mechanism.distortion import rotate
from .mechanism.distortion.geometric.affine import RotateConfig, RotateState
from .mechanism.distortion.photometric.blur import gaussian_kernel_u8


@attrs.define
class Texture:
    """One background image with the statistics metas.json carries (combiner.py:36-40)."""
    name: str
    image: Image
    grayscale_mean: float
    grayscale_std: float


@attrs.define
class ImageCombinerConfig:
    """ImageCombinerEngineInitConfig (combiner.py:72-81) without the folder."""
    target_image_mode: ImageMode = ImageMode.RGB
    enable_cache: bool = False
    prob_use_only_the_anchor_image: float = 0.7
    prob_rotate_image: float = 0.5
    sigma: float = 3.0
    init_segment_width_min_ratio: float = 0.25
    gaussian_blur_kernel_size: int = 5


@attrs.define
class Placement:
    """One pasted rectangle (inclusive bounds) and where its pixels come from."""
    texture: int
    rotated: bool
    up: int
    down: int
    left: int
    right: int


class _Span:
    """A stretch [left, right] of the canvas filled up to row `y` (exclusive).  Ordered by `y`
    alone, like the reference's PrioritizedSegment (combiner.py:84-88): ties are resolved by the
    heap's own sift order, which is part of the behaviour to reproduce."""
    __slots__ = ('y', 'left', 'right')

    def __init__(self, y: int, left: int, right: int):
        self.y, self.left, self.right = y, left, right

    def __lt__(self, other: '_Span'):
        return self.y < other.y


def load_textures_from_folder(folder: str) -> List[Texture]:
    """`<folder>/metas.json` + `<folder>/image/*` (combiner.py:48-69)."""
    from PIL import Image as PilImage
    folder = os.path.expandvars(os.path.expanduser(folder))
    with open(os.path.join(folder, 'metas.json')) as fin:
        metas = json.load(fin)
    textures = []
    for meta in metas:
        path = os.path.join(folder, 'image', meta['image_file'])
        pil_image = PilImage.open(path)
        pil_image.load()
        if pil_image.mode not in ('L', 'RGB', 'RGBA'):
            pil_image = pil_image.convert('RGB')
        # the mode follows the array like Image.from_pil_image (image.py:326-329)
        mat = np.array(pil_image, dtype=np.uint8)
        textures.append(Texture(name=path, image=Image(mat=mat),
                                grayscale_mean=meta['grayscale_mean'],
                                grayscale_std=meta['grayscale_std']))
    return textures


class ImageCombiner:
    """Device counterpart of ImageCombinerEngine (combiner.py:91-345): `run(height, width, rng)`
    returns the synthesised background as a device-resident Image."""

    def __init__(self, textures: Sequence[Texture], config: Optional[ImageCombinerConfig] = None):
        assert textures
        self.config = config or ImageCombinerConfig()
        # sorted by grayscale mean (stable, like sorted() in combiner.py:111-114)
        self.textures = sorted(textures, key=lambda texture: texture.grayscale_mean)
        self.grayscale_means = [texture.grayscale_mean for texture in self.textures]
        mode = self.config.target_image_mode
        self._plain: List[Image] = [texture.image.to_target_mode_image(mode)
                                    for texture in self.textures]
        self._rotated: Dict[int, Image] = {}
        # enable_cache keeps the FIRST orientation a texture was used in, across runs, and skips
        # the rotate draw for it afterwards (combiner.py:261-283)
        self._cached_orientation: Dict[int, bool] = {}

    @classmethod
    def from_folder(cls, image_meta_folder: str, config: Optional[ImageCombinerConfig] = None):
        return cls(load_textures_from_folder(image_meta_folder), config)

    # -- host side: which textures, where -----------------------------------------------------
    def sample_candidates(self, rng: RandomGenerator) -> List[int]:
        """sample_image_metas_based_on_random_anchor (combiner.py:122-145): indices into
        `self.textures`."""
        anchor = int(rng.choice(len(self.textures)))
        if rng.random() < self.config.prob_use_only_the_anchor_image:
            return [anchor]
        std = self.textures[anchor].grayscale_std
        mean = self.textures[anchor].grayscale_mean
        import bisect
        begin = bisect.bisect_left(self.grayscale_means, round(mean - self.config.sigma * std))
        end = bisect.bisect_right(self.grayscale_means, round(mean + self.config.sigma * std))
        candidates = list(range(begin, end))
        assert candidates
        return candidates

    def _texture_shape(self, index: int, rotated: bool) -> Tuple[int, int]:
        shape = self._plain[index].shape
        if rotated:  # host arithmetic only: the walk needs no pixels
            shape = RotateState(RotateConfig(angle=90), shape, None).result_shape
        return shape

    def _oriented(self, index: int, rotated: bool) -> Image:
        if not rotated:
            return self._plain[index]
        if index not in self._rotated:
            self._rotated[index] = rotate.distort_image({'angle': 90}, image=self._plain[index])
        return self._rotated[index]

    def plan(self, height: int, width: int, candidates: Sequence[int],
             rng: RandomGenerator) -> List[Placement]:
        """The skyline walk of synthesize_image (combiner.py:190-318) without the pixels."""
        config = self.config
        heap: List[_Span] = []
        cut_min = int(np.clip(round(config.init_segment_width_min_ratio * width), 1, width - 1))
        left = 0
        while left + cut_min - 1 < width:
            right = int(rng.integers(left + cut_min - 1, width))
            if right + 1 - left < cut_min or width - right - 1 < cut_min:
                break
            heap.append(_Span(0, left, right))
            left = right + 1
        if left < width:
            heap.append(_Span(0, left, width - 1))

        orientation: Dict[int, bool] = {}
        placements: List[Placement] = []
        while heap:
            span = heapq.heappop(heap)
            level: List[_Span] = []
            while heap and heap[0].y == span.y:
                level.append(heapq.heappop(heap))
            if level:
                level.append(span)
                level.sort(key=lambda item: item.left)
                at = next(i for i, item in enumerate(level)
                          if item.left == span.left and item.right == span.right)
                first = last = at
                while first > 0 and level[first - 1].right + 1 == level[first].left:
                    first -= 1
                while last + 1 < len(level) and level[last].right + 1 == level[last + 1].left:
                    last += 1
                if first < last:
                    span.left, span.right = level[first].left, level[last].right
                for item in level[:first]:
                    heapq.heappush(heap, item)
                for item in level[last + 1:]:
                    heapq.heappush(heap, item)

            index = candidates[int(rng.choice(len(candidates)))]
            if config.enable_cache and index in self._cached_orientation:
                rotated = self._cached_orientation[index]
            else:
                if index not in orientation:
                    orientation[index] = bool(rng.random() < config.prob_rotate_image)
                rotated = orientation[index]
                if config.enable_cache:
                    self._cached_orientation[index] = rotated
            tex_h, tex_w = self._texture_shape(index, rotated)

            up = span.y
            down = min(height - 1, up + tex_h - 1)
            left = span.left
            right = min(span.right, left + tex_w - 1)
            placements.append(Placement(index, rotated, up, down, left, right))

            if right == span.right:
                span.y = down + 1
                if span.y < height:
                    heapq.heappush(heap, span)
            else:
                below = _Span(down + 1, left, right)
                if below.y < height:
                    heapq.heappush(heap, below)
                span.left = right + 1
                heapq.heappush(heap, span)
        return placements

    # -- device side -------------------------------------------------------------------------
    def compose(self, height: int, width: int, placements: Sequence[Placement]) -> Image:
        config = self.config
        channels = self._plain[0].num_channels or 1
        items = np.zeros(len(placements), dtype=nv.PASTE_ITEM_DTYPE)
        keep = []
        for i, placement in enumerate(placements):
            source = self._oriented(placement.texture, placement.rotated)
            tensor = source.dev
            keep.append(tensor)
            items['src'][i] = tensor.data_ptr()
            items['src_pitch'][i] = source.width
            items['up'][i], items['down'][i] = placement.up, placement.down
            items['left'][i], items['right'][i] = placement.left, placement.right
        band = config.gaussian_blur_kernel_size // 2 + 1
        ksize = config.gaussian_blur_kernel_size
        taps = gaussian_kernel_u8(ksize, band / 3)
        shape = (height, width) if channels == 1 else (height, width, channels)
        dst = dv.empty(shape, np.uint8)
        table = dv.upload_structs(items) if len(placements) else dv.empty((16,), np.uint8)
        nv.check(nv.lib().vkb_background_compose(dv.ptr(dst), height, width, channels,
                                                 dv.ptr(table), len(placements), band,
                                                 (ctypes.c_int32 * ksize)(*taps), ksize,
                                                 dv.stream_ptr()), 'vkb_background_compose')
        del keep
        return Image(mat=dst, mode=config.target_image_mode)

    def run(self, height: int, width: int, rng: RandomGenerator) -> Image:
        """ImageCombinerEngine.run (combiner.py:335-345)."""
        assert rng is not None
        candidates = self.sample_candidates(rng)
        return self.compose(height, width, self.plan(height, width, candidates, rng))
