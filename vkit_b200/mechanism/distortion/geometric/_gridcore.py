"""Device plan of the grid-based ops for a BATCH of pages (batch = 1 for `Distortion.distort`).

Phase 1  project lattice -> finalise (round, shift, result shape)     [shapes come back]
Phase 2a per-cell homographies, coverage masks, tile candidate lists
Phase 2b fused remap of Image + Mask + ScoreMap

Replaces DistortionStateImageGridBased.initialize_image_grid_based and
ImageGrid.generate_remap_params + cv.remap (grid_rendering/interface.py:98-114,
type.py:209-261, grid_blender.py:54-81).
"""
import ctypes
from typing import List, Optional, Sequence

import numpy as np

from vkit_b200 import _native as nv
from vkit_b200 import device as dv

from ._hostmath import lattice_axis


_LATTICE_DIMS = {}


def new_grid_page(height: int, width: int, grid_size: int, out=None):
    """A zeroed page record with the source shape and lattice dimensions filled in.  `out`: a
    zeroed element of a GRID_PAGE_DTYPE array to fill in place (batch builders)."""
    rec = np.zeros((), dtype=nv.GRID_PAGE_DTYPE) if out is None else out
    key = (height, width, grid_size)
    dims = _LATTICE_DIMS.get(key)
    if dims is None:
        dims = (len(lattice_axis(height, grid_size)), len(lattice_axis(width, grid_size)))
        if len(_LATTICE_DIMS) < 1024:
            _LATTICE_DIMS[key] = dims
    rec['src_h'] = height
    rec['src_w'] = width
    rec['grid_size'] = grid_size
    rec['rows'], rec['cols'] = dims
    return rec


class GridBatch:
    """`dims_bound` = None: the constructor waits for the result shapes (one stream synchronise)
    and sizes every workspace exactly -- what the per-page `Distortion.distort` needs.

    `dims_bound` = (max_dst_h, max_dst_w): the OPTIMISTIC form for batches.  Workspaces are sized
    from the bound, the output layout is computed on the device (`vkb_grid_layout`) and nothing
    waits for the GPU until `finish()`; a batch that does not fit reports it there (`fits`) and is
    run again with exact sizes by the caller."""

    def __init__(self, pages: np.ndarray, keepalive: Sequence = (),
                 given_lattice: Optional[np.ndarray] = None, dims_bound=None, shared=None):
        """`shared`: a dict owned by the caller (one per batch engine) in which the plan keeps what
        does not change between the plans of that engine -- the device copy of the parameter
        blocks and the workspaces.  Successive plans of one engine then cost launches only; they
        must run on one stream at a time (stream order protects the reuse), and the workspaces
        always belong to the LATEST plan."""
        dv.require_cuda()
        self.lib = nv.lib()
        self.pages = np.ascontiguousarray(pages, dtype=nv.GRID_PAGE_DTYPE).reshape(-1)
        self.n = int(self.pages.shape[0])
        self.keepalive = list(keepalive)
        self.shared = shared if shared is not None else {}
        static = self.shared.get('static')
        if static is None:
            projectors = 0
            for kind in np.unique(self.pages['projector']):
                projectors |= 1 << int(kind)
            static = {
                'p_max': int((self.pages['rows'] * self.pages['cols']).max()),
                'c_max': int(((self.pages['rows'] - 1) * (self.pages['cols'] - 1)).max()),
                'projectors': projectors,
                'pages_dev': dv.upload_structs(self.pages),
            }
            self.shared['static'] = static
        self.p_max, self.c_max = static['p_max'], static['c_max']
        self.pages_dev = static['pages_dev']
        projectors = static['projectors']
        self.lattice_f = self._ws('lattice_f', (self.n, self.p_max, 2), np.float64)
        self.lattice_i = self._ws('lattice_i', (self.n, self.p_max, 2), np.int32)
        self.meta_dev = self._ws('meta_dev', (self.n * nv.GRID_META_DTYPE.itemsize,), np.uint8)
        stream = dv.stream_ptr()
        if given_lattice is not None:
            lat = np.zeros((self.n, self.p_max, 2), dtype=np.float64)
            lat[:, :given_lattice.shape[1]] = given_lattice
            self.lattice_f.copy_(dv.to_device(lat))
        nv.check(self.lib.vkb_grid_project(dv.ptr(self.pages_dev), self.n, self.p_max,
                                           dv.ptr(self.lattice_f), projectors, stream),
                 'vkb_grid_project')
        # result shapes are needed on the host to allocate the outputs: the kernel mirrors them
        # into pinned host memory (device accessible under UVA), so one stream synchronise is
        # enough -- a D2H copy would wait behind bulk copies queued on the copy engine
        t = dv.torch()
        self._meta_bytes = self.n * nv.GRID_META_DTYPE.itemsize
        meta_host = dv.pinned_mirror(self._meta_bytes)
        nv.check(self.lib.vkb_grid_finalize(dv.ptr(self.pages_dev), self.n, self.p_max,
                                            dv.ptr(self.lattice_f), dv.ptr(self.lattice_i),
                                            dv.ptr(self.meta_dev),
                                            ctypes.c_void_p(meta_host.data_ptr()), stream),
                 'vkb_grid_finalize')
        meta_host.written()
        self.meta_host = meta_host
        self.stream = t.cuda.current_stream()
        self.layout_host = None
        self._meta = None
        self.fits = True
        self.deferred = dims_bound is not None
        if self.deferred:
            self.max_dst_h, self.max_dst_w = int(dims_bound[0]), int(dims_bound[1])
        else:
            t.cuda.current_stream().synchronize()
            self._read_meta()
            self.max_dst_h = int(self._meta['dst_h'].max())
            self.max_dst_w = int(self._meta['dst_w'].max())
        self.tiles_x = (self.max_dst_w + nv.TILE - 1) // nv.TILE
        self.tiles_y = (self.max_dst_h + nv.TILE - 1) // nv.TILE
        self.t_max = self.tiles_x * self.tiles_y
        self.hinv = None
        self.hfwd = None

    def __del__(self):
        # a plan nobody asked for its shapes: the mirrors go back to the pool, guarded by the
        # events of their writers
        for name in ('meta_host', 'layout_host'):
            block = getattr(self, name, None)
            if block is not None:
                try:
                    block.release()
                except Exception:
                    pass

    def _ws(self, name, shape, dtype):
        """A workspace tensor: from the engine's shared dict when the shape still fits, else new."""
        key = ('ws', name)
        t = self.shared.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = dv.empty(shape, dtype)
            self.shared[key] = t
        return t

    @property
    def meta(self):
        if self._meta is None:
            self.finish()
        return self._meta

    def finish(self):
        """Optimistic plans: wait for the stream, read the result shapes (and the layout status)
        from the pinned mirrors.  Returns `fits`."""
        if self._meta is None:
            self.stream.synchronize()
            self._read_meta()
            if self.layout_host is not None:
                layout = self.layout_host.numpy()[:(self.n + 2) * 8].view(np.int64)
                self.fits = int(layout[self.n + 1]) == 0
                self.pixel_offsets = np.array(layout[:self.n + 1], dtype=np.int64)
                dv.release_mirror(self.layout_host)
                self.layout_host = None
        return self.fits

    def _read_meta(self):
        """(after a synchronise) result shapes out of the pinned mirror; the block goes back."""
        raw = self.meta_host.numpy()[:self._meta_bytes].tobytes()
        self._meta = np.frombuffer(raw, dtype=nv.GRID_META_DTYPE)
        dv.release_mirror(self.meta_host)
        self.meta_host = None

    def result_shape(self, i: int = 0):
        return int(self.meta['dst_h'][i]), int(self.meta['dst_w'][i])

    def build(self, need_forward: bool = False):
        if self.hinv is not None and (self.hfwd is not None or not need_forward):
            return
        self.hinv = self._ws('hinv', (self.n, self.c_max, 9), np.float64)
        if need_forward:
            self.hfwd = self._ws('hfwd', (self.n, self.c_max, 9), np.float64)
        self.cell_box = self._ws('cell_box', (self.n, self.c_max, 4), np.int32)
        self.cell_masks = self._ws('cell_masks', (self.n, self.c_max, nv.CELL_MASK_WORDS), np.uint32)
        self.tile_count = self._ws('tile_count', (self.n, self.t_max), np.int32)
        self.tile_cells = self._ws('tile_cells', (self.n, self.t_max, nv.TILE_CAP), np.uint16)
        # candidate records of the remap kernel: 16 per tile on average is ample (typical 8-10);
        # a page that needs more falls back to the kernel's slow exact path for the excess tiles
        # records of the page: one per cell, then the sorted candidate lists of the tiles
        self.s_cap = max(16 * self.t_max, self.c_max + 2 * self.t_max)
        self.tile_off = self._ws('tile_off', (self.n, self.t_max), np.int32)
        self.tile_base = self._ws('tile_base', (self.n + 1,), np.int32)
        self.tile_slots = self._ws('tile_slots', (self.n, self.s_cap, nv.TILE_SLOT_BYTES), np.uint8)
        self.tile_headers = self._ws(
            'tile_headers',
            (self.n * self.t_max * nv.TILE_HEADER_BYTES + (self.n * self.t_max + 2) * 4,), np.uint8)
        nv.check(self.lib.vkb_grid_build(
            dv.ptr(self.pages_dev), self.n, self.p_max, self.c_max, self.t_max, self.s_cap,
            dv.ptr(self.lattice_i), dv.ptr(self.meta_dev), dv.ptr(self.hinv), dv.ptr(self.hfwd),
            dv.ptr(self.cell_box), dv.ptr(self.cell_masks), dv.ptr(self.tile_count),
            dv.ptr(self.tile_cells), dv.ptr(self.tile_off), dv.ptr(self.tile_base),
            dv.ptr(self.tile_slots), dv.ptr(self.tile_headers), dv.stream_ptr()), 'vkb_grid_build')

    def layout(self, planes: np.ndarray, cap_pixels: int):
        """Optimistic plans: `planes` carries the source fields and the arena BASES in its dst
        pointers; offsets and result shapes are filled in on the device.  Returns the device copy
        of the records for `remap(planes_dev=...)`."""
        t = dv.torch()
        planes = np.ascontiguousarray(planes, dtype=nv.PLANES_DTYPE).reshape(-1)
        planes_dev = dv.upload_structs(planes)
        self.layout_dev = self._ws('layout_dev', (self.n + 2,), np.int64)
        self.layout_host = dv.pinned_mirror((self.n + 2) * 8)
        nv.check(self.lib.vkb_grid_layout(
            dv.ptr(self.meta_dev), self.n, dv.ptr(planes_dev), int(cap_pixels), self.t_max,
            dv.ptr(self.layout_dev), ctypes.c_void_p(self.layout_host.data_ptr()), dv.stream_ptr()),
            'vkb_grid_layout')
        self.layout_host.written()
        return planes_dev

    def remap(self, planes: np.ndarray, launch_events=None, planes_dev=None):
        """planes: structured array (PLANES_DTYPE), one record per page, device pointers.
        `launch_events`: optional list that receives a (start, end) CUDA event pair recorded
        immediately around the kernel launch.  `planes_dev`: the records already on the device
        (from `layout`); `planes` then only names the containers."""
        self.build()
        planes = np.ascontiguousarray(planes, dtype=nv.PLANES_DTYPE).reshape(-1)
        channels = int(planes['image_channels'][0])
        has_mask = int(planes['src_mask'][0] != 0)
        has_score = int(planes['src_score'][0] != 0)
        if ((planes['image_channels'] != channels).any()
                or ((planes['src_mask'] != 0) != bool(has_mask)).any()
                or ((planes['src_score'] != 0) != bool(has_score)).any()):
            raise ValueError('all pages of one remap call must carry the same containers')
        if int(planes['src_h'].max()) >= 32768 or int(planes['src_w'].max()) >= 32768:
            raise ValueError('source planes larger than 32767 px are not supported (cv.remap limit)')
        if planes_dev is None:
            planes_dev = dv.upload_structs(planes)
            if ((planes['dst_h'] != self.meta['dst_h']).any()
                    or (planes['dst_w'] != self.meta['dst_w']).any()):
                raise ValueError('planes.dst_h / dst_w must be the result shapes of the plan')
        if launch_events is not None:
            t = dv.torch()
            ev0, ev1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            ev0.record()
        nv.check(self.lib.vkb_grid_remap(
            dv.ptr(self.pages_dev), dv.ptr(planes_dev), self.n, self.p_max, self.c_max, self.t_max,
            self.s_cap, dv.ptr(self.lattice_i), dv.ptr(self.hinv), dv.ptr(self.cell_box),
            dv.ptr(self.cell_masks), dv.ptr(self.tile_count), dv.ptr(self.tile_off),
            dv.ptr(self.tile_base), dv.ptr(self.tile_slots), dv.ptr(self.tile_headers), channels,
            has_mask, has_score, dv.stream_ptr()), 'vkb_grid_remap')
        if launch_events is not None:
            ev1.record()
            launch_events.append((ev0, ev1))
        return planes_dev

    def transform_points(self, page: int, xy: np.ndarray, cell_rc: np.ndarray) -> np.ndarray:
        """xy: n x 2 float64 smooth (x, y); cell_rc: n x 2 int32 (polygon_row, polygon_col)."""
        self.build(need_forward=True)
        n = int(xy.shape[0])
        if n == 0:
            return np.zeros((0, 2), dtype=np.float64)
        ccols = int(self.pages['cols'][page]) - 1
        xy_dev = dv.to_device(np.ascontiguousarray(xy, dtype=np.float64))
        rc_dev = dv.to_device(np.ascontiguousarray(cell_rc, dtype=np.int32))
        out = dv.empty((n, 2), np.float64)
        nv.check(self.lib.vkb_grid_points(dv.ptr(self.hfwd[page]), ccols, dv.ptr(xy_dev),
                                          dv.ptr(rc_dev), dv.ptr(out), n, dv.stream_ptr()),
                 'vkb_grid_points')
        return dv.to_host(out)

    def lattice_points(self, page: int = 0) -> np.ndarray:
        """Integer dst lattice (rows, cols, 2) as (x, y), on the host."""
        rows, cols = int(self.pages['rows'][page]), int(self.pages['cols'][page])
        return dv.to_host(self.lattice_i[page, :rows * cols]).reshape(rows, cols, 2)


def planes_record(image=None, mask=None, score_map=None, dst_shape=None):
    """Allocate outputs for one page and describe the fused launch.  Inputs are elements;
    returns (record, out_image_tensor, out_mask_tensor, out_score_tensor)."""
    rec = np.zeros((), dtype=nv.PLANES_DTYPE)
    first = image or mask or score_map
    src_h, src_w = first.shape
    dst_h, dst_w = dst_shape
    rec['src_h'], rec['src_w'], rec['dst_h'], rec['dst_w'] = src_h, src_w, dst_h, dst_w
    out_image = out_mask = out_score = None
    if image is not None:
        channels = image.num_channels or 1
        if image.mat_dtype != np.uint8 or channels not in (1, 3, 4):
            raise NotImplementedError('geometric ops support uint8 images with 1, 3 or 4 channels')
        shape = (dst_h, dst_w) if image.num_channels == 0 else (dst_h, dst_w, channels)
        out_image = dv.empty(shape, np.uint8)
        rec['src_image'] = image.dev.data_ptr()
        rec['dst_image'] = out_image.data_ptr()
        rec['image_channels'] = channels
    if mask is not None:
        out_mask = dv.empty((dst_h, dst_w), np.uint8)
        rec['src_mask'] = mask.dev.data_ptr()
        rec['dst_mask'] = out_mask.data_ptr()
    if score_map is not None:
        out_score = dv.empty((dst_h, dst_w), np.float32)
        rec['src_score'] = score_map.dev.data_ptr()
        rec['dst_score'] = out_score.data_ptr()
    return rec, out_image, out_mask, out_score
