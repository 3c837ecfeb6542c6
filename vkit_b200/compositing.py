"""Text-layer compositing on the device.

The reference composes a page by calling `fill_image` / `fill_np_array` once per glyph, text line,
symbol and seal (vkit/engine/font/freetype.py:314-380, vkit/pipeline/text_detection/
page_assembler.py:152-236, page_distortion.py:146-161) -- a Python loop of small NumPy blends.
The element classes of this package already route each of those calls to `vkb_blend_fill`; this
module adds the batched form: a `DrawList` records the fills of one destination and flushes them
as ONE upload + ONE launch (`vkb_blend_draw_list`), applying overlapping items in list order so
the result is identical to the sequential calls.
"""
import ctypes
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _native as nv
from . import device as dv
from .element import Box, Image, Mask, ScoreMap


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class DrawList:
    """Ordered fills into one Image / Mask / ScoreMap."""

    def __init__(self, target: Union[Image, Mask, ScoreMap]):
        self.target = target
        self.dst_f32 = target.mat_dtype == np.float32
        self.channels = 1 if target.mat_ndim == 2 else target.mat_shape[2]
        self.np_dtype = np.float32 if self.dst_f32 else np.uint8
        self.items: List[np.ndarray] = []
        self.blobs: List[Tuple[int, str, np.ndarray]] = []  # (item index, field, host array)
        self.device_refs = []

    def __len__(self):
        return len(self.items)

    def _relative_box(self, box: Box) -> Box:
        relative_box, _ = box.get_boxes_for_box_attached_opt(self.target.box)
        return relative_box

    def fill(self, box: Box, value, alpha: Union[float, np.ndarray, ScoreMap] = 1.0,
             mask: Optional[Union[Mask, np.ndarray]] = None, keep_max_value: bool = False,
             keep_min_value: bool = False):
        """Same meaning as `box.fill_image(target, value, image_mask=mask, alpha=alpha)`
        (element/box.py:394-416) / `fill_mask` / `fill_score_map` with keep_max / keep_min."""
        rel = self._relative_box(box)
        item = np.zeros((), dtype=nv.BLEND_ITEM_DTYPE)
        item['dst_f32'] = int(self.dst_f32)
        item['channels'] = self.channels
        item['dst_w'] = self.target.width
        item['box_y'], item['box_x'] = rel.up, rel.left
        item['box_h'], item['box_w'] = rel.height, rel.width
        idx = len(self.items)

        if isinstance(value, (Image, Mask, ScoreMap)):
            value = value.mat
        if isinstance(value, np.ndarray):
            if value.shape[:2] == self.target.shape and rel.shape != self.target.shape:
                value = value[rel.up:rel.down + 1, rel.left:rel.right + 1]
            if value.shape[:2] != rel.shape:
                raise RuntimeError('value is np.ndarray but shape is not matched.')
            self.blobs.append((idx, 'value_arr', np.ascontiguousarray(value, dtype=self.np_dtype)))
            item['value_pitch'] = rel.width
        else:
            if isinstance(value, tuple):
                if self.channels > 1 and len(value) != self.channels:
                    raise RuntimeError('value is tuple but len(value) != num_channels.')
                consts = list(value)
            else:
                consts = [value] * self.channels
            consts = np.asarray(consts).astype(self.np_dtype)
            item['value_const'][:len(consts[:4])] = consts[:4]

        if isinstance(alpha, ScoreMap):
            assert alpha.is_prob
            alpha = alpha.mat
        if isinstance(alpha, np.ndarray):
            if alpha.shape != rel.shape:
                raise RuntimeError('alpha array shape does not match the box.')
            self.blobs.append((idx, 'alpha_arr', np.ascontiguousarray(alpha, dtype=np.float32)))
            item['alpha_pitch'] = rel.width
            item['alpha'] = 1.0
        else:
            alpha = float(alpha)
            if alpha < 0.0 or alpha > 1.0:
                raise RuntimeError(f'alpha={alpha} is invalid.')
            if alpha == 0.0:
                self.blobs = [b for b in self.blobs if b[0] != idx]
                return
            item['alpha'] = alpha

        if mask is not None:
            if isinstance(mask, Mask):
                mask = mask.mat
            if mask.shape != rel.shape:
                raise RuntimeError('mask shape does not match the box.')
            self.blobs.append((idx, 'mask', np.ascontiguousarray(mask > 0, dtype=np.uint8)))
            item['mask_pitch'] = rel.width
        assert not (keep_max_value and keep_min_value)
        item['keep_mode'] = 1 if keep_max_value else (2 if keep_min_value else 0)
        if rel.height > 0 and rel.width > 0:
            self.items.append(item)
        else:
            self.blobs = [b for b in self.blobs if b[0] != idx]

    def fill_score_map(self, score_map: ScoreMap, value):
        """`score_map.fill_image(target, value)` (score_map.py:678-687): alpha = the scores,
        active where alpha > 0."""
        self.fill(score_map.equivalent_box, value, alpha=score_map)

    def fill_mask(self, mask: Mask, value, alpha: Union[float, np.ndarray, ScoreMap] = 1.0):
        """`mask.fill_image(target, value, alpha)` (mask.py:601-612)."""
        self.fill(mask.equivalent_box, value, alpha=alpha, mask=mask)

    def flush(self):
        """One upload of every alpha / mask / value array + the item table, one launch."""
        if not self.items:
            return self.target
        # staging buffer: [items][blob 0][blob 1]...
        items = np.stack(self.items)
        head = _align(items.nbytes)
        offsets = []
        total = head
        for _, _, arr in self.blobs:
            offsets.append(total)
            total = _align(total + arr.nbytes)
        staging = np.zeros(total, dtype=np.uint8)
        for off, (_, _, arr) in zip(offsets, self.blobs):
            staging[off:off + arr.nbytes] = arr.reshape(-1).view(np.uint8)
        dst = self.target.dev
        staged = dv.empty((total,), np.uint8)
        base = staged.data_ptr()
        items['dst'] = dst.data_ptr()
        for off, (idx, field, _) in zip(offsets, self.blobs):
            items[field][idx] = base + off
        staging[:items.nbytes] = items.reshape(-1).view(np.uint8)
        staged.copy_(dv.torch().from_numpy(staging))
        nv.check(nv.lib().vkb_blend_draw_list(dv.ptr(staged), len(self.items), self.target.height,
                                              self.target.width, dv.stream_ptr()),
                 'vkb_blend_draw_list')
        self.target._after_device_write()
        self.device_refs.append(staged)
        self.items, self.blobs = [], []
        return self.target


def render_char_glyphs_in_text_line(glyph_color: Tuple[int, int, int], text_line_height: int,
                                    text_line_width: int, glyph_images: Sequence[np.ndarray],
                                    glyph_score_maps: Sequence[np.ndarray],
                                    char_boxes: Sequence[Box]):
    """Default / monochrome branch of render_char_glyphs_in_text_line (freetype.py:314-353):
    white line image, glyph colour where the glyph bitmap is non-zero, mask = 1 there, score map
    merged with keep-max.  Three draw lists, three launches, whatever the number of glyphs."""
    image = Image(mat=np.full((text_line_height, text_line_width, 3), 255, dtype=np.uint8))
    mask = Mask(mat=np.zeros((text_line_height, text_line_width), dtype=np.uint8))
    score_map = ScoreMap.from_shape((text_line_height, text_line_width))
    dl_image, dl_mask, dl_score = DrawList(image), DrawList(mask), DrawList(score_map)
    for glyph_image, glyph_score, box in zip(glyph_images, glyph_score_maps, char_boxes):
        glyph_mask = glyph_image > 0
        dl_image.fill(box, tuple(glyph_color), mask=glyph_mask)
        dl_mask.fill(box, 1, mask=glyph_mask)
        dl_score.fill(box, glyph_score, keep_max_value=True)
    dl_image.flush()
    dl_mask.flush()
    dl_score.flush()
    return image, mask, score_map


def assemble_text_lines(background: Image, text_line_score_maps: Sequence[ScoreMap],
                        glyph_colors: Sequence[Tuple[int, int, int]]) -> Image:
    """The text-line loop of PageAssemblerStep.run (page_assembler.py:174-179):
    `text_line.score_map.fill_image(assembled_image, text_line.glyph_color)` for every line."""
    assembled = background.copy()
    draw_list = DrawList(assembled)
    for score_map, color in zip(text_line_score_maps, glyph_colors):
        draw_list.fill_score_map(score_map, tuple(color))
    return draw_list.flush()


def fill_page_inactive_region(page_image: Image, page_active_mask: Mask,
                              page_bottom_layer_image: Image):
    """PageDistortionStep.fill_page_inactive_region (page_distortion.py:146-161), in place."""
    assert page_image.shape == page_active_mask.shape
    if page_bottom_layer_image.shape != page_image.shape:
        # cv.resize INTER_CUBIC like the reference (the wheel's IPP cubic)
        page_bottom_layer_image = page_bottom_layer_image.to_resized_image(
            resized_height=page_image.height, resized_width=page_image.width)
    page_active_mask.to_inverted_mask().fill_image(page_image, page_bottom_layer_image)


# =============================================================================================
# Label rasterisation after distortion
# =============================================================================================
def fill_polygons(target: Union[Mask, ScoreMap], polygons, value=1, keep_max_value: bool = False,
                  keep_min_value: bool = False):
    """`for polygon, v in zip(polygons, values): polygon.fill_mask(target, v, keep_max_value=...)`
    (or `fill_score_map`) as ONE ordered device pass: every polygon is rasterised with cv.fillPoly
    semantics and later polygons overwrite earlier ones (vkit/element/polygon.py:458-487 as looped
    by pipeline/text_detection/page_distortion.py:163-314 and engine/char_mask/default.py:44-56).
    `value`: one number for all polygons or one per polygon.  The target is updated in place."""
    polygons = list(polygons)
    if not polygons:
        return target
    if target.box is not None:
        raise NotImplementedError('fill_polygons expects a target without an attached box')
    values = list(value) if isinstance(value, (list, tuple, np.ndarray)) else [value] * len(polygons)
    if len(values) != len(polygons):
        raise ValueError('one value per polygon expected')
    mode = 1 if keep_max_value else (2 if keep_min_value else 0)
    items = np.zeros(len(polygons), dtype=nv.POLY_ITEM_DTYPE)
    pts = []
    first = 0
    for i, (polygon, v) in enumerate(zip(polygons, values)):
        xy = np.asarray(polygon.to_np_array(), dtype=np.int32).reshape(-1, 2)
        items['first_pt'][i] = first
        items['n_pts'][i] = xy.shape[0]
        items['y_min'][i] = int(xy[:, 1].min())
        items['y_max'][i] = int(xy[:, 1].max())
        items['value'][i] = v
        first += xy.shape[0]
        pts.append(xy)
    dst_f32 = target.mat_dtype == np.float32
    pts_dev = dv.to_device(np.ascontiguousarray(np.concatenate(pts)))
    items_dev = dv.upload_structs(items)
    height, width = target.shape
    keys = dv.empty((height, width), np.int32)
    dst = target.dev
    lib = nv.lib()
    # at most 65535 polygons per launch; ordering across launches is the list order
    for a in range(0, len(polygons), 65535):
        b = min(a + 65535, len(polygons))
        offset = a * nv.POLY_ITEM_DTYPE.itemsize
        nv.check(lib.vkb_fill_polygons(
            dv.ptr(dst), int(dst_f32), height, width, dv.ptr(pts_dev),
            ctypes.c_void_p(items_dev.data_ptr() + offset),
            ctypes.c_void_p(items.ctypes.data + offset), b - a, mode, dv.ptr(keys),
            dv.stream_ptr()), 'vkb_fill_polygons')
    target._after_device_write()
    return target


# =============================================================================================
# page_resizing
# =============================================================================================
def resize_page_elements(page_image: Image, masks: Sequence[Mask], height_score_maps: Sequence[ScoreMap],
                         resize_ratio: float, cv_resize_interpolation: int):
    """The data path of PageResizingStep.run (pipeline/text_detection/page_resizing.py:112-180):
    the page image, its masks (active / char / seal-impression char / text-line) and its height
    score maps (char / text-line) resized to round(ratio * shape) with ONE sampled cv2
    interpolation (NEAREST_EXACT, LINEAR_EXACT, CUBIC, LANCZOS4, or AREA when shrinking --
    utility/opt.py:125-148), the height scores multiplied by the ratio afterwards
    (page_resizing.py:160-161, 177-180).  Seven device resamples, no host copy.
    Returns (image, [masks], [score maps])."""
    height, width = page_image.shape
    resized_height = round(resize_ratio * height)
    resized_width = round(resize_ratio * width)
    image = page_image.to_resized_image(resized_height=resized_height, resized_width=resized_width,
                                        cv_resize_interpolation=cv_resize_interpolation)
    resized_masks = []
    for mask in masks:
        assert mask.shape == (height, width)
        resized_masks.append(mask.to_resized_mask(resized_height=resized_height,
                                                  resized_width=resized_width,
                                                  cv_resize_interpolation=cv_resize_interpolation))
    resized_score_maps = []
    for score_map in height_score_maps:
        assert score_map.shape == (height, width)
        resized = score_map.to_resized_score_map(resized_height=resized_height,
                                                 resized_width=resized_width,
                                                 cv_resize_interpolation=cv_resize_interpolation)
        # "Scores are resized as well": float32 map times the Python float ratio, as NumPy does
        resized.assign_mat(resized.dev * float(np.float32(resize_ratio)))
        resized_score_maps.append(resized)
    return image, resized_masks, resized_score_maps
