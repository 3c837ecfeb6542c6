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
from typing import Dict, Hashable, List, Optional, Sequence, Tuple, Union

import attrs
import numpy as np

from . import _native as nv
from . import device as dv
from .element import Box, Image, Mask, ScoreMap


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


@attrs.define(eq=False)
class DeviceBlob:
    """A region-shaped array that already lives in HBM (a glyph plane of the atlas, a texture):
    draw-list items point at it instead of uploading a copy per fill."""
    tensor: object  # the CUDA tensor that owns the bytes (kept alive by whoever holds the blob)
    offset: int  # bytes from tensor.data_ptr()
    shape: Tuple[int, ...]  # (h, w) or (h, w, channels)
    pitch: int  # pixels per row
    dtype: np.dtype

    @property
    def ptr(self) -> int:
        return self.tensor.data_ptr() + self.offset

    def to_host(self) -> np.ndarray:
        """Download (tests / debugging)."""
        count = int(np.prod(self.shape))
        nbytes = count * np.dtype(self.dtype).itemsize
        assert self.pitch == self.shape[1]
        raw = dv.to_host(self.tensor.view(dv.torch().uint8).reshape(-1)[self.offset:self.offset + nbytes])
        return raw.view(self.dtype).reshape(self.shape).copy()


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

        if isinstance(value, DeviceBlob):
            if tuple(value.shape[:2]) != rel.shape or np.dtype(value.dtype) != self.np_dtype:
                raise RuntimeError('device value does not match the box / destination dtype.')
            item['value_arr'] = value.ptr
            item['value_pitch'] = value.pitch
            self.device_refs.append(value.tensor)
        elif isinstance(value, (Image, Mask, ScoreMap)):
            value = value.mat
        if isinstance(value, DeviceBlob):
            pass
        elif isinstance(value, np.ndarray):
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
        if isinstance(alpha, DeviceBlob):
            if tuple(alpha.shape) != rel.shape or np.dtype(alpha.dtype) != np.float32:
                raise RuntimeError('device alpha does not match the box.')
            item['alpha_arr'] = alpha.ptr
            item['alpha_pitch'] = alpha.pitch
            item['alpha'] = 1.0
            self.device_refs.append(alpha.tensor)
        elif isinstance(alpha, np.ndarray):
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

        if isinstance(mask, DeviceBlob):
            if tuple(mask.shape) != rel.shape or np.dtype(mask.dtype) != np.uint8:
                raise RuntimeError('device mask does not match the box.')
            item['mask'] = mask.ptr
            item['mask_pitch'] = mask.pitch
            self.device_refs.append(mask.tensor)
        elif mask is not None:
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


# =============================================================================================
# Glyph atlas: FreeType coverage bitmaps uploaded once, their
# blend planes derived on the device, text lines rendered from device pointers
# =============================================================================================
def trim_glyph_bitmap(bitmap: np.ndarray):
    """Rows / columns without coverage removed (trim_char_np_image_vert / _hori,
    freetype.py:100-134): returns (trimmed, pad_up, pad_down, pad_left, pad_right)."""
    covered = bitmap if bitmap.ndim == 2 else bitmap.max(axis=2)
    rows = np.flatnonzero(covered.max(axis=1))
    cols = np.flatnonzero(covered.max(axis=0))
    if rows.size == 0 or cols.size == 0:
        raise RuntimeError('trim_glyph_bitmap: empty bitmap.')
    up, down, left, right = int(rows[0]), int(rows[-1]), int(cols[0]), int(cols[-1])
    height, width = covered.shape
    return (bitmap[up:down + 1, left:right + 1], up, height - 1 - down, left, width - 1 - right)


@attrs.define(eq=False)
class AtlasGlyph:
    """One cached glyph: the metrics build_char_glyph derives (freetype.py:136-221) and, once the
    atlas is committed, the device planes the renderer blends from."""
    key: Hashable
    height: int
    width: int
    channels: int  # 1: default / monochrome, 3: LCD
    gamma: float
    ascent: int
    pad_up: int
    pad_down: int
    pad_left: int
    pad_right: int
    bitmap: Optional[DeviceBlob] = None
    mask: Optional[DeviceBlob] = None
    alpha: Optional[DeviceBlob] = None  # float32 score map (default / monochrome glyphs)
    lcd_image: Optional[DeviceBlob] = None  # uint8 H x W x 3 (LCD glyphs)

    @property
    def shape(self):
        return self.height, self.width


class GlyphAtlas:
    """Device-resident glyph cache.  `add` registers a FreeType bitmap under a key (font, char,
    size ...), `commit` uploads everything added since the last commit in ONE copy and derives
    the mask / alpha / LCD planes with ONE launch (`vkb_glyph_prepare`).  Arena pages are never
    moved, so blobs handed out stay valid for the life of the atlas."""

    def __init__(self, page_bytes: int = 8 << 20):
        self.page_bytes = page_bytes
        self.pages: List[object] = []
        self.cursor = 0
        self.glyphs: Dict[Hashable, AtlasGlyph] = {}
        self._pending: List[Tuple[AtlasGlyph, np.ndarray]] = []
        self._luts: Dict[float, Tuple[DeviceBlob, DeviceBlob]] = {}

    def __contains__(self, key):
        return key in self.glyphs

    def __getitem__(self, key) -> AtlasGlyph:
        return self.glyphs[key]

    @staticmethod
    def gamma_tables(gamma: float):
        """The two functions of a coverage byte the reference evaluates per glyph pixel, tabulated
        with the same NumPy expressions: float32 alpha `np.power(v.astype(float32) / 255.0, gamma)`
        (freetype.py:176-180) and the LCD image `((1 - np.power(v / 255.0, gamma)) * 255)
        .astype(uint8)` (freetype.py:361-366)."""
        v = np.arange(256, dtype=np.uint8)
        alpha = np.power(v.astype(np.float32) / 255.0, gamma)
        lcd = ((1 - np.power(v / 255.0, gamma)) * 255).astype(np.uint8)
        return np.ascontiguousarray(alpha, dtype=np.float32), lcd

    def add(self, key: Hashable, bitmap: np.ndarray, gamma: float = 1.0, bitmap_top: int = 0,
            bitmap_left: int = 0, advance_x: Optional[int] = None) -> AtlasGlyph:
        """Register the bitmap FreeType rendered for one char (H x W, or H x W x 3 in LCD mode).
        `bitmap_top`, `bitmap_left`, `advance_x` (26.6 fixed point) are the glyph-slot fields
        build_char_glyph reads; the paddings and the ascent follow freetype.py:150-173."""
        if key in self.glyphs:
            return self.glyphs[key]
        bitmap = np.asarray(bitmap, dtype=np.uint8)
        assert bitmap.ndim == 2 or (bitmap.ndim == 3 and bitmap.shape[2] == 3)
        full_width = bitmap.shape[1]
        trimmed, pad_up, pad_down, pad_left_inc, pad_right_inc = trim_glyph_bitmap(bitmap)
        # vertical trim first, then the bearings, then the horizontal trim, like the reference
        pad_left = max(0, bitmap_left)
        if advance_x is None:
            pad_right = 0
        else:
            pad_right = max(0, round(advance_x / 64) - pad_left - full_width)
        glyph = AtlasGlyph(key=key, height=trimmed.shape[0], width=trimmed.shape[1],
                           channels=1 if trimmed.ndim == 2 else 3, gamma=float(gamma),
                           ascent=bitmap_top - pad_up, pad_up=pad_up, pad_down=pad_down,
                           pad_left=pad_left + pad_left_inc, pad_right=pad_right + pad_right_inc)
        self.glyphs[key] = glyph
        self._pending.append((glyph, np.ascontiguousarray(trimmed)))
        return glyph

    def _reserve(self, nbytes: int):
        nbytes = _align(nbytes)
        if not self.pages or self.cursor + nbytes > self.pages[-1].numel():
            self.pages.append(dv.empty((max(self.page_bytes, nbytes),), np.uint8))
            self.cursor = 0
        page, offset = self.pages[-1], self.cursor
        self.cursor += nbytes
        return page, offset

    def commit(self):
        if not self._pending:
            return
        new_gammas = sorted({glyph.gamma for glyph, _ in self._pending} - set(self._luts))
        # chunk layout: [tables][bitmaps] uploaded, then [masks][alphas][lcd images] derived
        upload_size = 0
        lut_at = {}
        for gamma in new_gammas:
            lut_at[gamma] = upload_size
            upload_size += 1024 + 256
        bitmap_at = []
        for glyph, bitmap in self._pending:
            bitmap_at.append(upload_size)
            upload_size = _align(upload_size + bitmap.nbytes)
        total = upload_size
        mask_at, alpha_at, lcd_at = [], [], []
        for glyph, bitmap in self._pending:
            pixels = glyph.height * glyph.width
            mask_at.append(total)
            total = _align(total + pixels)
            if glyph.channels == 1:
                alpha_at.append(total)
                lcd_at.append(None)
                total = _align(total + pixels * 4)
            else:
                alpha_at.append(None)
                lcd_at.append(total)
                total = _align(total + pixels * 3)
        page, base = self._reserve(total)
        staging = np.zeros(upload_size, dtype=np.uint8)
        for gamma in new_gammas:
            alpha_lut, lcd_lut = self.gamma_tables(gamma)
            at = lut_at[gamma]
            staging[at:at + 1024] = alpha_lut.view(np.uint8)
            staging[at + 1024:at + 1280] = lcd_lut
            self._luts[gamma] = (DeviceBlob(page, base + at, (1, 256), 256, np.float32),
                                 DeviceBlob(page, base + at + 1024, (1, 256), 256, np.uint8))
        for at, (glyph, bitmap) in zip(bitmap_at, self._pending):
            staging[at:at + bitmap.nbytes] = bitmap.reshape(-1)
        page[base:base + upload_size].copy_(dv.torch().from_numpy(staging))
        items = np.zeros(len(self._pending), dtype=nv.GLYPH_ITEM_DTYPE)
        origin = page.data_ptr() + base
        for i, (glyph, bitmap) in enumerate(self._pending):
            shape2 = (glyph.height, glyph.width)
            shape = shape2 if glyph.channels == 1 else shape2 + (3,)
            glyph.bitmap = DeviceBlob(page, base + bitmap_at[i], shape, glyph.width, np.uint8)
            glyph.mask = DeviceBlob(page, base + mask_at[i], shape2, glyph.width, np.uint8)
            alpha_lut, lcd_lut = self._luts[glyph.gamma]
            items['bitmap'][i] = origin + bitmap_at[i]
            items['mask'][i] = origin + mask_at[i]
            items['n_pixels'][i] = glyph.height * glyph.width
            items['channels'][i] = glyph.channels
            if glyph.channels == 1:
                glyph.alpha = DeviceBlob(page, base + alpha_at[i], shape2, glyph.width, np.float32)
                items['alpha'][i] = origin + alpha_at[i]
                items['alpha_lut'][i] = alpha_lut.ptr
            else:
                glyph.lcd_image = DeviceBlob(page, base + lcd_at[i], shape, glyph.width, np.uint8)
                items['lcd_image'][i] = origin + lcd_at[i]
                items['lcd_lut'][i] = lcd_lut.ptr
        table = dv.upload_structs(items)
        nv.check(nv.lib().vkb_glyph_prepare(dv.ptr(table), items.ctypes.data_as(ctypes.c_void_p),
                                            len(items), dv.stream_ptr()), 'vkb_glyph_prepare')
        self._pending = []


def _launch_glyph_items(target, boxes_yxhw: np.ndarray, *, value_const=None, value_ptrs=None,
                        value_pitches=None, mask_ptrs=None, mask_pitches=None, alpha_ptrs=None,
                        alpha_pitches=None, keep_max: bool = False):
    """One ordered draw-list launch whose items all point at device-resident planes: the item table
    is filled column by column (no per-item Python objects)."""
    n = boxes_yxhw.shape[0]
    items = np.zeros(n, dtype=nv.BLEND_ITEM_DTYPE)
    dst = target.dev
    items['dst'] = dst.data_ptr()
    items['dst_f32'] = int(target.mat_dtype == np.float32)
    items['channels'] = 1 if target.mat_ndim == 2 else target.mat_shape[2]
    items['dst_w'] = target.width
    items['box_y'], items['box_x'] = boxes_yxhw[:, 0], boxes_yxhw[:, 1]
    items['box_h'], items['box_w'] = boxes_yxhw[:, 2], boxes_yxhw[:, 3]
    items['alpha'] = 1.0
    if value_const is not None:  # one tuple for all items or one row per item
        consts = np.asarray(value_const).astype(np.float32 if items['dst_f32'][0] else np.uint8)
        items['value_const'][:, :consts.shape[-1]] = consts.astype(np.float32)
    if value_ptrs is not None:
        items['value_arr'], items['value_pitch'] = value_ptrs, value_pitches
    if mask_ptrs is not None:
        items['mask'], items['mask_pitch'] = mask_ptrs, mask_pitches
    if alpha_ptrs is not None:
        items['alpha_arr'], items['alpha_pitch'] = alpha_ptrs, alpha_pitches
    items['keep_mode'] = 1 if keep_max else 0
    table = dv.upload_structs(items)
    nv.check(nv.lib().vkb_blend_draw_list(dv.ptr(table), n, target.height, target.width,
                                          dv.stream_ptr()), 'vkb_blend_draw_list')
    target._after_device_write()
    return table


def render_atlas_glyphs_in_text_line(glyph_color: Tuple[int, int, int], text_line_height: int,
                                     text_line_width: int, glyphs: Sequence[AtlasGlyph],
                                     char_boxes: Sequence[Box]):
    """render_char_glyphs_in_text_line (freetype.py:314-380) from atlas glyphs: white line image;
    default / monochrome glyphs paint `glyph_color` under the glyph mask and merge their alpha
    into the score map with keep-max; LCD glyphs (H x W x 3 bitmaps) paste their gamma-corrected
    inverted image under the mask, ignore `glyph_color` and yield no score map.  No glyph pixel
    crosses the bus: the draw-list items point into the atlas; three launches per line (two for
    LCD), their item tables built as arrays."""
    assert glyphs and len(glyphs) == len(char_boxes)
    t = dv.require_cuda()
    device = dv.device()
    image = Image(mat=t.full((text_line_height, text_line_width, 3), 255, dtype=t.uint8,
                             device=device))
    mask = Mask(mat=t.zeros((text_line_height, text_line_width), dtype=t.uint8, device=device))
    lcd = glyphs[0].channels == 3
    score_map = None
    if not lcd:
        score_map = ScoreMap(mat=t.zeros((text_line_height, text_line_width), dtype=t.float32,
                                         device=device), skip_prob_check=True)
    n = len(glyphs)
    boxes = np.empty((n, 4), dtype=np.int64)
    mask_ptrs = np.empty(n, dtype=np.uint64)
    plane_ptrs = np.empty(n, dtype=np.uint64)
    pitches = np.empty(n, dtype=np.int64)
    for i, (glyph, box) in enumerate(zip(glyphs, char_boxes)):
        if glyph.mask is None:
            raise RuntimeError('GlyphAtlas.commit() has not run for this glyph.')
        assert (glyph.channels == 3) == lcd and box.shape == glyph.shape
        if box.up < 0 or box.left < 0 or box.down >= text_line_height \
                or box.right >= text_line_width:
            raise RuntimeError('char box outside the text line.')
        boxes[i] = (box.up, box.left, glyph.height, glyph.width)
        mask_ptrs[i] = glyph.mask.ptr
        plane_ptrs[i] = glyph.lcd_image.ptr if lcd else glyph.alpha.ptr
        pitches[i] = glyph.width
    keep = []
    if lcd:
        keep.append(_launch_glyph_items(image, boxes, value_ptrs=plane_ptrs, value_pitches=pitches,
                                        mask_ptrs=mask_ptrs, mask_pitches=pitches))
    else:
        keep.append(_launch_glyph_items(image, boxes, value_const=tuple(glyph_color),
                                        mask_ptrs=mask_ptrs, mask_pitches=pitches))
        keep.append(_launch_glyph_items(score_map, boxes, value_ptrs=plane_ptrs,
                                        value_pitches=pitches, keep_max=True))
    keep.append(_launch_glyph_items(mask, boxes, value_const=(1,), mask_ptrs=mask_ptrs,
                                    mask_pitches=pitches))
    return image, mask, score_map


def render_atlas_text_lines(lines):
    """Many text lines at once: `lines` = [(glyph_color, height, width, glyphs, boxes)] with `boxes`
    a list of Box or an (n, 2) integer array of (up, left) per glyph.  Same result per line as
    render_atlas_glyphs_in_text_line, but the lines of a page (or of a batch of pages) share three
    launches: every draw-list item carries its own destination pointer and pitch, so one ordered
    list can paint into many line images.  Returns [(Image, Mask, ScoreMap or None)]."""
    t = dv.require_cuda()
    device = dv.device()
    if not lines:
        return []
    lcd = lines[0][3][0].channels == 3
    sizes = np.asarray([(int(h), int(w)) for _, h, w, _, _ in lines], dtype=np.int64)
    pixels = sizes[:, 0] * sizes[:, 1]
    offsets = np.concatenate([[0], np.cumsum(pixels)])
    total = int(offsets[-1])
    image_arena = t.full((total * 3,), 255, dtype=t.uint8, device=device)
    mask_arena = t.zeros((total,), dtype=t.uint8, device=device)
    score_arena = None if lcd else t.zeros((total,), dtype=t.float32, device=device)
    n_items = sum(len(glyphs) for _, _, _, glyphs, _ in lines)
    boxes = np.empty((n_items, 4), dtype=np.int64)
    line_of = np.empty(n_items, dtype=np.int64)
    mask_ptrs = np.empty(n_items, dtype=np.uint64)
    plane_ptrs = np.empty(n_items, dtype=np.uint64)
    pitches = np.empty(n_items, dtype=np.int64)
    colors = np.zeros((n_items, 3), dtype=np.int64)
    k = 0
    for li, (color, height, width, glyphs, glyph_boxes) in enumerate(lines):
        for gi, glyph in enumerate(glyphs):
            if glyph.mask is None:
                raise RuntimeError('GlyphAtlas.commit() has not run for this glyph.')
            assert (glyph.channels == 3) == lcd
            box = glyph_boxes[gi]
            up, left = (box.up, box.left) if isinstance(box, Box) else (int(box[0]), int(box[1]))
            if up < 0 or left < 0 or up + glyph.height > height or left + glyph.width > width:
                raise RuntimeError('char box outside the text line.')
            boxes[k] = (up, left, glyph.height, glyph.width)
            line_of[k] = li
            mask_ptrs[k] = glyph.mask.ptr
            plane_ptrs[k] = glyph.lcd_image.ptr if lcd else glyph.alpha.ptr
            pitches[k] = glyph.width
            colors[k] = color
            k += 1
    max_h, max_w = int(sizes[:, 0].max()), int(sizes[:, 1].max())
    line_w = sizes[line_of, 1]

    def launch(arena, bytes_per_pixel, channels, dst_f32, **fields):
        items = np.zeros(n_items, dtype=nv.BLEND_ITEM_DTYPE)
        items['dst'] = np.uint64(arena.data_ptr()) + (offsets[line_of] * bytes_per_pixel).astype(np.uint64)
        items['dst_f32'] = int(dst_f32)
        items['channels'] = channels
        items['dst_w'] = line_w
        items['box_y'], items['box_x'] = boxes[:, 0], boxes[:, 1]
        items['box_h'], items['box_w'] = boxes[:, 2], boxes[:, 3]
        items['alpha'] = 1.0
        for name, value in fields.items():
            if name == 'value_const':
                items['value_const'][:, :value.shape[1]] = value
            else:
                items[name] = value
        table = dv.upload_structs(items)
        nv.check(nv.lib().vkb_blend_draw_list(dv.ptr(table), n_items, max_h, max_w, dv.stream_ptr()),
                 'vkb_blend_draw_list')
        return table

    keep = []
    if lcd:
        keep.append(launch(image_arena, 3, 3, False, value_arr=plane_ptrs, value_pitch=pitches,
                           mask=mask_ptrs, mask_pitch=pitches))
    else:
        keep.append(launch(image_arena, 3, 3, False,
                           value_const=colors.astype(np.uint8).astype(np.float32),
                           mask=mask_ptrs, mask_pitch=pitches))
        keep.append(launch(score_arena, 4, 1, True, value_arr=plane_ptrs, value_pitch=pitches,
                           keep_mode=1))
    keep.append(launch(mask_arena, 1, 1, False, value_const=np.ones((n_items, 1), np.float32),
                       mask=mask_ptrs, mask_pitch=pitches))
    out = []
    for li, (height, width) in enumerate(sizes):
        height, width = int(height), int(width)
        a, b = int(offsets[li]), int(offsets[li + 1])
        image = Image(mat=image_arena[a * 3:b * 3].view(height, width, 3))
        mask = Mask(mat=mask_arena[a:b].view(height, width))
        score_map = None
        if not lcd:
            score_map = ScoreMap(mat=score_arena[a:b].view(height, width), skip_prob_check=True)
        out.append((image, mask, score_map))
    return out


def render_char_glyphs_in_text_line(glyph_color: Tuple[int, int, int], text_line_height: int,
                                    text_line_width: int, glyph_images: Sequence[np.ndarray],
                                    glyph_score_maps: Optional[Sequence[np.ndarray]],
                                    char_boxes: Sequence[Box], glyph_color_gamma: float = 1.0):
    """render_char_glyphs_in_text_line (freetype.py:314-380) from host glyph arrays.
    Default / monochrome branch (2-D glyph images): white line image, glyph colour where the glyph
    bitmap is non-zero, mask = 1 there, score map merged with keep-max -- three draw lists, three
    launches, whatever the number of glyphs.  LCD branch (H x W x 3 glyph images,
    freetype.py:355-370): the glyphs go through a throw-away GlyphAtlas."""
    if glyph_images[0].ndim == 3:
        atlas = GlyphAtlas(page_bytes=1 << 20)
        glyphs = []
        for i, glyph_image in enumerate(glyph_images):
            # no trimming here: the caller's char boxes fit the arrays as they are
            glyph = AtlasGlyph(key=i, height=glyph_image.shape[0], width=glyph_image.shape[1],
                               channels=3, gamma=float(glyph_color_gamma), ascent=0, pad_up=0,
                               pad_down=0, pad_left=0, pad_right=0)
            atlas.glyphs[i] = glyph
            atlas._pending.append((glyph, np.ascontiguousarray(glyph_image, dtype=np.uint8)))
            glyphs.append(glyph)
        atlas.commit()
        return render_atlas_glyphs_in_text_line(glyph_color, text_line_height, text_line_width,
                                                glyphs, char_boxes)
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
        if isinstance(polygon, np.ndarray):
            # smooth (x, y) vertices, e.g. from RandomDistortionBatch: the rounded twins of
            # Point.create (Python round, half to even) like Polygon.to_np_array
            xy = np.rint(polygon).astype(np.int32).reshape(-1, 2)
        else:
            xy = np.asarray(polygon.to_np_array(), dtype=np.int32).reshape(-1, 2)
        items['first_pt'][i] = first
        items['n_pts'][i] = xy.shape[0]
        items['y_min'][i] = int(xy[:, 1].min())
        items['y_max'][i] = int(xy[:, 1].max())
        items['value'][i] = v
        first += xy.shape[0]
        pts.append(xy)
    dst_f32 = target.mat_dtype == np.float32
    # staged like the item tables (pinned block + copy kernel): a pageable cudaMemcpy would wait
    # for everything queued on the stream
    pts_dev = dv.upload_structs(np.ascontiguousarray(np.concatenate(pts)))
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
    (page_resizing.py:160-161, 177-180).  Seven device resamples (the score scaling inside them), no host copy.
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
        # "Scores are resized as well": float32 map times the Python float ratio as NumPy does it
        # (one float32 product with float32(ratio)), inside the resize kernel
        resized_score_maps.append(score_map.to_resized_score_map(
            resized_height=resized_height, resized_width=resized_width,
            cv_resize_interpolation=cv_resize_interpolation,
            post_scale=float(np.float32(resize_ratio))))
    return image, resized_masks, resized_score_maps
