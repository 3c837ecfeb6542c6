"""Batched page engine: many independent pages through one set of launches.

`Distortion.distort` handles one page per call, which is launch- and Python-bound (a 1024^2 RGB
page is ~1 us of HBM traffic).  This module keeps the same ops and configs but runs a whole
batch of pages per launch: per-page parameter blocks in one device array, ragged outputs in one
arena, one small D2H per batch for the result shapes.

    engine = GeometricBatch.from_configs(distortions, configs, shape)
    out = engine.run(images_dev)          # images_dev: (B, H, W, 3) uint8 CUDA tensor
    out.image(i)                          # (H'_i, W'_i, 3) view into the output arena
"""
import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nv
from . import device as dv
from .mechanism.distortion.geometric import camera as _camera
from .mechanism.distortion.geometric import mls as _mls
from .mechanism.distortion.geometric._gridcore import GridBatch
from .utility import dyn_structure

_PAGE_BUILDERS = {
    'camera_plane_only': (_camera.CameraPlaneOnlyConfig, _camera.plane_only_page),
    'camera_cubic_curve': (_camera.CameraCubicCurveConfig, _camera.cubic_curve_page),
    'camera_plane_line_fold': (_camera.CameraPlaneLineFoldConfig, _camera.plane_line_fold_page),
    'camera_plane_line_curve': (_camera.CameraPlaneLineCurveConfig, _camera.plane_line_curve_page),
}


def grid_page_record(op_name: str, config, shape: Tuple[int, int], out=None, handle_sink=None):
    """(page record, keepalive) for one page of a grid-based op.  `out`: zeroed element of a
    GRID_PAGE_DTYPE array to fill in place; `handle_sink`: see similarity_mls_page."""
    height, width = shape
    if op_name == 'similarity_mls':
        config = dyn_structure(config, _mls.SimilarityMlsConfig)
        return _mls.similarity_mls_page(config, shape, out, handle_sink)
    config_cls, builder = _PAGE_BUILDERS[op_name]
    config = dyn_structure(config, config_cls)
    rec = builder(config, shape, out)
    camera_model_config = _camera.complete_camera_model_config(height, width,
                                                               config.camera_model_config)
    _camera.fill_camera_model(rec, camera_model_config)
    return rec, None


class BatchOutput:
    """Ragged outputs of one batch: flat arenas + per-page offsets.  For an optimistic batch the
    shapes and offsets are read from the device when they are first asked for (`resolve`)."""

    def __init__(self, shapes, channels, image_arena, mask_arena, score_arena, pixel_offsets,
                 resolve=None):
        self._shapes = shapes
        self.channels = channels
        self.image_arena = image_arena
        self.mask_arena = mask_arena
        self.score_arena = score_arena
        self._pixel_offsets = pixel_offsets
        self._resolve = resolve

    def _ready(self):
        if self._resolve is not None:
            resolve, self._resolve = self._resolve, None
            resolve(self)

    @property
    def shapes(self):
        self._ready()
        return self._shapes

    @property
    def pixel_offsets(self):
        self._ready()
        return self._pixel_offsets

    @property
    def total_pixels(self):
        return int(self.pixel_offsets[-1])

    def packed(self, arena, per_pixel: int = 1):
        """The used part of an arena (optimistic batches allocate by a bound): what the next
        batched stage takes as its flat input."""
        return arena[:self.total_pixels * per_pixel]

    def _view(self, arena, i, per_pixel):
        h, w = self.shapes[i]
        start = int(self.pixel_offsets[i]) * per_pixel
        flat = arena[start:start + h * w * per_pixel]
        return flat.view(h, w, per_pixel) if per_pixel > 1 else flat.view(h, w)

    def image(self, i):
        return self._view(self.image_arena, i, self.channels)

    def mask(self, i):
        return self._view(self.mask_arena, i, 1)

    def score_map(self, i):
        return self._view(self.score_arena, i, 1)


# Optimistic batches: result canvases are assumed to stay within these factors of the source
# (side lengths / total pixels); a batch that does not is detected on the device and run again
# with exact sizes.  The reference's policies reach ~1.3 x per side at the strongest levels.
DIMS_BOUND_FACTOR = 1.5
PIXELS_BOUND_FACTOR = 1.5


def _planes_from_lists(planes, shapes, images, masks, score_maps, channels):
    """Source side of the plane records when every page is its own tensor (pages that went
    through different ops before): one pointer per page."""
    n = len(shapes)
    keep = []
    if images is not None:
        if channels is None:
            channels = 1 if images[0].dim() == 2 else int(images[0].shape[2])
        for i, (t, (h, w)) in enumerate(zip(images, shapes)):
            t = t if t.is_contiguous() else t.contiguous()
            if int(t.numel()) != h * w * channels:
                raise ValueError(f'image {i} does not match its page shape')
            keep.append(t)
            planes['src_image'][i] = t.data_ptr()
        planes['image_channels'] = channels
    else:
        channels = 0
    for name, tensors in (('src_mask', masks), ('src_score', score_maps)):
        if tensors is None:
            continue
        for i, (t, (h, w)) in enumerate(zip(tensors, shapes)):
            t = t if t.is_contiguous() else t.contiguous()
            if int(t.numel()) != h * w:
                raise ValueError(f'{name[4:]} {i} does not match its page shape')
            keep.append(t)
            planes[name][i] = t.data_ptr()
    return planes, channels, keep


class GeometricBatch:
    """A batch of same-size pages, each with its own grid-op config."""

    def __init__(self, op_names: Sequence[str], configs: Sequence, shape):
        """`shape`: (H, W) shared by all pages, or one (H, W) per page (ragged input, e.g. the
        output of a previous geometric batch)."""
        self.n = len(op_names)
        if len(shape) == 2 and not isinstance(shape[0], (tuple, list)):
            self.shapes = [(int(shape[0]), int(shape[1]))] * self.n
        else:
            self.shapes = [(int(h), int(w)) for h, w in shape]
            if len(self.shapes) != self.n:
                raise ValueError('one shape per page expected')
        self.shape = self.shapes[0]
        keepalive = []
        self.pages = np.zeros(self.n, dtype=nv.GRID_PAGE_DTYPE)
        handle_sink = []
        for i, (op_name, config) in enumerate(zip(op_names, configs)):
            _, keep = grid_page_record(op_name, config, self.shapes[i], out=self.pages[i],
                                       handle_sink=handle_sink)
            if keep is not None:
                keepalive.append(keep)
        handles = _mls.upload_handles(handle_sink) if handle_sink else None
        if handles is not None:
            keepalive.append(handles)
        self.keepalive = keepalive
        self.plan: Optional[GridBatch] = None
        # what successive plans of this engine share (device copy of the parameter blocks,
        # workspaces): see GridBatch.  One stream at a time per engine.
        self._shared = {}
        src_pixels = np.asarray([h * w for h, w in self.shapes], dtype=np.int64)
        self._n_src = int(src_pixels.sum())
        self._src_offsets = np.concatenate([[0], np.cumsum(src_pixels)])[:-1].astype(np.uint64)
        self._planes_template = np.zeros(self.n, dtype=nv.PLANES_DTYPE)
        self._planes_template['src_h'] = [s[0] for s in self.shapes]
        self._planes_template['src_w'] = [s[1] for s in self.shapes]

    def plan_batch(self, optimistic: bool = False):
        """Phase 1 + 2a: lattices, result shapes, cell homographies, masks, bins.  The exact form
        waits for the result shapes (one stream synchronise); the optimistic form sizes the
        workspaces from DIMS_BOUND_FACTOR and never waits (see GridBatch)."""
        bound = None
        if optimistic:
            bound = (int(max(s[0] for s in self.shapes) * DIMS_BOUND_FACTOR) + 1,
                     int(max(s[1] for s in self.shapes) * DIMS_BOUND_FACTOR) + 1)
        self.plan = GridBatch(self.pages, keepalive=self.keepalive, dims_bound=bound,
                              shared=self._shared)
        if not optimistic:
            self.plan.build()
        return self.plan

    def run(self, images=None, masks=None, score_maps=None, replan: bool = True,
            launch_events=None, channels: Optional[int] = None,
            optimistic: bool = False) -> BatchOutput:
        """images: (B, H, W, C) uint8, masks: (B, H, W) uint8, score_maps: (B, H, W) float32 --
        CUDA tensors (any subset); or, for ragged pages, flat arenas that hold the pages back to
        back in page order (`channels` then names the image's channel count).  Returns ragged
        outputs in fresh arenas.
        `launch_events`: optional list; a (start, end) pair of CUDA events recorded immediately
        around the fused remap launch is appended (kernel time without host work).
        `optimistic`: no host round trip inside the call -- arenas sized by PIXELS_BOUND_FACTOR,
        output layout computed on the device, shapes read when the BatchOutput is first asked
        for them (a batch that does not fit the bounds is then run again the exact way)."""
        if optimistic:
            return self._run_optimistic(images, masks, score_maps, launch_events, channels)
        if replan or self.plan is None or self.plan.deferred:
            self.plan_batch()
        plan = self.plan
        shapes = [plan.result_shape(i) for i in range(self.n)]
        pixels = np.asarray([h * w for h, w in shapes], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(pixels)])
        total = int(offsets[-1])
        planes, channels = self._source_planes(images, masks, score_maps, channels)
        planes['dst_h'] = [s[0] for s in shapes]
        planes['dst_w'] = [s[1] for s in shapes]
        image_arena = mask_arena = score_arena = None
        if images is not None:
            image_arena = dv.empty((total * channels,), np.uint8)
            planes['dst_image'] = image_arena.data_ptr() + (offsets[:-1] * channels).astype(
                np.uint64)
        if masks is not None:
            mask_arena = dv.empty((total,), np.uint8)
            planes['dst_mask'] = mask_arena.data_ptr() + offsets[:-1].astype(np.uint64)
        if score_maps is not None:
            score_arena = dv.empty((total,), np.float32)
            planes['dst_score'] = score_arena.data_ptr() + (offsets[:-1] * 4).astype(np.uint64)
        plan.remap(planes, launch_events=launch_events)
        return BatchOutput(shapes, channels, image_arena, mask_arena, score_arena, offsets)

    def _source_planes(self, images, masks, score_maps, channels):
        """Plane records with the source side filled in (+ validation of the inputs)."""
        src_offsets, n_src = self._src_offsets, self._n_src
        planes = self._planes_template.copy()
        if isinstance(images, (list, tuple)) or isinstance(masks, (list, tuple)) \
                or isinstance(score_maps, (list, tuple)):
            planes, channels, self._keep_sources = _planes_from_lists(
                planes, self.shapes, images, masks, score_maps, channels)
            return planes, channels
        if images is not None:
            if channels is None:
                channels = 1 if images.dim() == 3 else (int(images.shape[3]) if images.dim() == 4
                                                        else None)
            if channels is None:
                raise ValueError('flat image arenas need `channels`')
            if int(images.numel()) != n_src * channels:
                raise ValueError('images do not hold the pages of this batch')
            planes['src_image'] = np.uint64(images.data_ptr()) + src_offsets * np.uint64(channels)
            planes['image_channels'] = channels
        else:
            channels = 0
        if masks is not None:
            if int(masks.numel()) != n_src:
                raise ValueError('masks do not hold the pages of this batch')
            planes['src_mask'] = np.uint64(masks.data_ptr()) + src_offsets
        if score_maps is not None:
            if int(score_maps.numel()) != n_src:
                raise ValueError('score maps do not hold the pages of this batch')
            planes['src_score'] = np.uint64(score_maps.data_ptr()) + src_offsets * np.uint64(4)
        return planes, channels

    def _run_optimistic(self, images, masks, score_maps, launch_events, channels):
        plan = self.plan_batch(optimistic=True)
        planes, channels = self._source_planes(images, masks, score_maps, channels)
        cap = int(self._n_src * PIXELS_BOUND_FACTOR) + 1
        image_arena = mask_arena = score_arena = None
        if images is not None:
            image_arena = dv.empty((cap * channels,), np.uint8)
            planes['dst_image'] = image_arena.data_ptr()
        if masks is not None:
            mask_arena = dv.empty((cap,), np.uint8)
            planes['dst_mask'] = mask_arena.data_ptr()
        if score_maps is not None:
            score_arena = dv.empty((cap,), np.float32)
            planes['dst_score'] = score_arena.data_ptr()
        planes_dev = plan.layout(planes, cap)
        plan.build()
        plan.remap(planes, launch_events=launch_events, planes_dev=planes_dev)
        inputs = (images, masks, score_maps)

        def resolve(out: BatchOutput):
            if plan.finish():
                out._shapes = [plan.result_shape(i) for i in range(self.n)]
                out._pixel_offsets = plan.pixel_offsets
                return
            # the bounds were too small for this batch: once more, exactly sized
            exact = self.run(*inputs, channels=channels or None)
            out._shapes, out._pixel_offsets = exact._shapes, exact._pixel_offsets
            out.image_arena, out.mask_arena, out.score_arena = (exact.image_arena, exact.mask_arena,
                                                                exact.score_arena)

        return BatchOutput(None, channels, image_arena, mask_arena, score_arena, None, resolve)

    def algorithmic_bytes(self, channels: int = 3, with_mask: bool = False,
                          with_score: bool = False) -> int:
        """Compulsory traffic of the batch: every source pixel read once, every destination
        pixel written once."""
        per_px = channels + (1 if with_mask else 0) + (4 if with_score else 0)
        src = sum(h * w for h, w in self.shapes)
        dst = sum(h * w for h, w in (self.plan.result_shape(i) for i in range(self.n)))
        return per_px * (src + dst)


_AFFINE_STATES = None


def _affine_states():
    global _AFFINE_STATES
    if _AFFINE_STATES is None:
        from .mechanism.distortion.geometric import affine as _affine
        _AFFINE_STATES = {
            'shear_hori': (_affine.ShearHoriConfig, _affine.ShearHoriState),
            'shear_vert': (_affine.ShearVertConfig, _affine.ShearVertState),
            'rotate': (_affine.RotateConfig, _affine.RotateState),
            'skew_hori': (_affine.SkewHoriConfig, _affine.SkewHoriState),
            'skew_vert': (_affine.SkewVertConfig, _affine.SkewVertState),
        }
    return _AFFINE_STATES


class AffineBatch:
    """rotate / shear_* / skew_* over a (possibly ragged) batch of pages in ONE launch of
    `vkb_warp_fused`: per-page forward matrix and dsize from the reference's state classes
    (geometric/affine.py:92-395), per-page inverse in the parameter block, ragged output arena."""

    def __init__(self, op_names: Sequence[str], configs: Sequence, shape):
        from .mechanism.distortion.geometric._hostmath import invert_affine
        self.n = len(op_names)
        if len(shape) == 2 and not isinstance(shape[0], (tuple, list)):
            self.shapes = [(int(shape[0]), int(shape[1]))] * self.n
        else:
            self.shapes = [(int(h), int(w)) for h, w in shape]
        self.records = np.zeros(self.n, dtype=nv.WARP_PAGE_DTYPE)
        self.result_shapes = []
        self.identity = []
        self.forward = []  # forward matrix per page (None: identity), for the points
        for i, (name, config) in enumerate(zip(op_names, configs)):
            config_cls, state_cls = _affine_states()[name]
            config = dyn_structure(config, config_cls)
            state = state_cls(config, self.shapes[i], None)
            trans_mat, dsize = state.trans_mat, state.dsize
            if trans_mat is None or getattr(config, 'is_nop', False):
                self.identity.append(True)
                self.forward.append(None)
                self.result_shapes.append(self.shapes[i])
                self.records['kind'][i] = nv.WARP_AFFINE
                self.records['inv'][i, :6] = [1, 0, 0, 0, 1, 0]
                continue
            self.identity.append(False)
            self.forward.append(trans_mat)
            self.result_shapes.append((int(dsize[1]), int(dsize[0])))
            if trans_mat.shape[0] == 2:
                self.records['kind'][i] = nv.WARP_AFFINE
                self.records['inv'][i, :6] = invert_affine(trans_mat).reshape(-1)
            else:
                self.records['kind'][i] = nv.WARP_PERSPECTIVE
                self.records['inv'][i] = np.linalg.inv(
                    np.asarray(trans_mat, dtype=np.float64)).reshape(-1)

    def run(self, images=None, masks=None, score_maps=None, channels: Optional[int] = None):
        """Same input conventions as GeometricBatch.run (4-D batch or flat arenas)."""
        shapes = self.result_shapes
        pixels = np.asarray([h * w for h, w in shapes], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(pixels)])
        total = int(offsets[-1])
        src_pixels = np.asarray([h * w for h, w in self.shapes], dtype=np.int64)
        src_offsets = np.concatenate([[0], np.cumsum(src_pixels)])[:-1].astype(np.uint64)
        # a fresh block of container pointers per call: a second run() with other containers must
        # not see the (possibly freed) arenas of the previous one
        planes = np.zeros(self.n, dtype=nv.PLANES_DTYPE)
        planes['src_h'] = [s[0] for s in self.shapes]
        planes['src_w'] = [s[1] for s in self.shapes]
        planes['dst_h'] = [s[0] for s in shapes]
        planes['dst_w'] = [s[1] for s in shapes]
        image_arena = mask_arena = score_arena = None
        if isinstance(images, (list, tuple)) or isinstance(masks, (list, tuple)) \
                or isinstance(score_maps, (list, tuple)):
            planes, channels, self._keep_sources = _planes_from_lists(
                planes, self.shapes, images, masks, score_maps, channels)
            if images is not None:
                image_arena = dv.empty((total * channels,), np.uint8)
                planes['dst_image'] = image_arena.data_ptr() + (offsets[:-1] * channels).astype(
                    np.uint64)
            if masks is not None:
                mask_arena = dv.empty((total,), np.uint8)
                planes['dst_mask'] = mask_arena.data_ptr() + offsets[:-1].astype(np.uint64)
            if score_maps is not None:
                score_arena = dv.empty((total,), np.float32)
                planes['dst_score'] = score_arena.data_ptr() + (offsets[:-1] * 4).astype(np.uint64)
            images = masks = score_maps = None
            from_lists = True
        else:
            from_lists = False
        if images is not None:
            if channels is None:
                channels = 1 if images.dim() == 3 else (int(images.shape[3]) if images.dim() == 4
                                                        else None)
            if channels is None:
                raise ValueError('flat image arenas need `channels`')
            if int(images.numel()) != int(src_pixels.sum()) * channels:
                raise ValueError('images do not hold the pages of this batch')
            image_arena = dv.empty((total * channels,), np.uint8)
            planes['src_image'] = np.uint64(images.data_ptr()) + src_offsets * np.uint64(channels)
            planes['dst_image'] = image_arena.data_ptr() + (offsets[:-1] * channels).astype(
                np.uint64)
            planes['image_channels'] = channels
        elif not from_lists:
            channels = 0
        if masks is not None:
            if int(masks.numel()) != int(src_pixels.sum()):
                raise ValueError('masks do not hold the pages of this batch')
            mask_arena = dv.empty((total,), np.uint8)
            planes['src_mask'] = np.uint64(masks.data_ptr()) + src_offsets
            planes['dst_mask'] = mask_arena.data_ptr() + offsets[:-1].astype(np.uint64)
        if score_maps is not None:
            if int(score_maps.numel()) != int(src_pixels.sum()):
                raise ValueError('score maps do not hold the pages of this batch')
            score_arena = dv.empty((total,), np.float32)
            planes['src_score'] = np.uint64(score_maps.data_ptr()) + src_offsets * np.uint64(4)
            planes['dst_score'] = score_arena.data_ptr() + (offsets[:-1] * 4).astype(np.uint64)
        if images is None and masks is None and score_maps is None and not from_lists:
            raise ValueError('nothing to warp')
        if from_lists and image_arena is None:
            channels = 0
        records = self.records.copy()
        records['planes'] = planes
        pages_dev = dv.upload_structs(records)
        nv.check(nv.lib().vkb_warp_fused(dv.ptr(pages_dev), self.n,
                                         max(s[0] for s in shapes), max(s[1] for s in shapes),
                                         dv.stream_ptr()), 'vkb_warp_fused')
        return BatchOutput(shapes, channels, image_arena, mask_arena, score_arena, offsets)


_PIPELINE_STATE = {}


def _pipeline_state():
    """Long-lived side streams per device (the caching allocator keeps one pool per stream, so
    fresh streams per call would mean fresh cudaMallocs per call): one for the H2D copies, two
    alternating work streams (plan + remap), one for the D2H copies."""
    t = dv.require_cuda()
    key = t.cuda.current_device()
    if key not in _PIPELINE_STATE:
        _PIPELINE_STATE[key] = {'h2d': t.cuda.Stream(), 'work': [t.cuda.Stream(), t.cuda.Stream()],
                                'd2h': t.cuda.Stream(), 'staging': None}
    return _PIPELINE_STATE[key]


def distort_pages_host(op_names: Sequence[str], configs: Sequence, shape: Tuple[int, int],
                       host_images, host_out=None, chunk_pages: int = 32,
                       use_thread: bool = False):
    """Host buffers in, host buffers out -- the end-to-end form of the batch engine.

    `host_images`: pinned (B, H, W, C) uint8 CPU tensor; the distorted pages land back to back in
    the pinned flat uint8 tensor `host_out` (allocated when None).  Four streams keep both copy
    engines and the SMs busy at once:
      * every chunk's H2D copy is queued up front on the copy-in stream (pixels do not depend on
        the configs), one event per chunk;
      * per chunk the host turns the configs into parameter blocks (Rodrigues, translation,
        curve / line constants), plans it -- lattice, result shapes (the only host sync, on a
        work stream that never carries pixel copies), cells, masks, records -- and launches its
        remap behind the chunk's H2D event (`use_thread` moves the parameter blocks to a helper
        thread; measured slower under the GIL);
      * the D2H copy of a finished chunk runs on the copy-out stream.
    Returns (host_out, shapes, byte_offsets)."""
    import queue
    import threading
    t = dv.require_cuda()
    n = len(op_names)
    # a short first chunk gets the D2H engine going early (the pipeline is PCIe-bound: the fill
    # time before the first copy-out is pure loss)
    # ... and the chunks then grow by 1.5 x up to `chunk_pages`: a chunk's result must be ready
    # when the copy-out of the chunk before it ends, and D2H moves a page ~1.3 x slower than H2D
    # brings the next one in (measured with tools/e2e_timeline_probe.py: a full-size second chunk
    # left the copy-out engine idle for 0.9 ms)
    first = min(n, max(1, chunk_pages // 4))
    bounds, a, size = [], 0, first
    while a < n:
        b = min(a + size, n)
        bounds.append((a, b))
        a = b
        size = min(chunk_pages, max(size + 1, (size * 3) // 2))
    # ... and a short last chunk keeps the drain (its kernels + its D2H, nothing to overlap) short
    a, b = bounds[-1]
    if b - a > first:
        bounds[-1:] = [(a, b - first), (b - first, b)]
    ready: 'queue.Queue' = queue.Queue(maxsize=3)

    def producer():
        try:
            for a, b in bounds:
                ready.put(GeometricBatch(op_names[a:b], configs[a:b], shape))
        except BaseException as exc:  # surface errors in the consumer
            ready.put(exc)

    channels = 1 if host_images.dim() == 3 else int(host_images.shape[3])
    if host_out is None:
        bound = int(n * shape[0] * shape[1] * channels * 2.25)
        host_out = t.empty((bound,), dtype=t.uint8).pin_memory()
    if use_thread:
        thread = threading.Thread(target=producer, daemon=True)
        thread.start()
    state = _pipeline_state()
    current = t.cuda.current_stream()
    for s in [state['h2d'], state['d2h']] + state['work']:
        s.wait_stream(current)

    # 1. all H2D copies, back to back
    staging = state['staging']
    if staging is None or staging.shape != host_images.shape:
        with t.cuda.stream(state['h2d']):
            staging = t.empty(host_images.shape, dtype=t.uint8, device=dv.device())
        state['staging'] = staging
    copied = []
    with t.cuda.stream(state['h2d']):
        for a, b in bounds:
            staging[a:b].copy_(host_images[a:b], non_blocking=True)
            ev = t.cuda.Event()
            ev.record()
            copied.append(ev)

    # 2. plan + remap per chunk on alternating work streams, D2H behind each
    shapes, offsets = [], [0]
    keep = []
    for i, (a, b) in enumerate(bounds):
        sub = ready.get() if use_thread else GeometricBatch(op_names[a:b], configs[a:b], shape)
        if isinstance(sub, BaseException):
            raise sub
        work = state['work'][i % 2]
        with t.cuda.stream(work):
            sub.plan_batch()
            work.wait_event(copied[i])
            out = sub.run(staging[a:b], replan=False)
            done = t.cuda.Event()
            done.record()
        n_bytes = out.total_pixels * channels
        start = offsets[-1]
        if start + n_bytes > host_out.numel():
            raise ValueError('host_out is too small for the distorted pages')
        with t.cuda.stream(state['d2h']):
            state['d2h'].wait_event(done)
            host_out[start:start + n_bytes].copy_(out.packed(out.image_arena, channels),
                                                  non_blocking=True)
        for k, (h, w) in enumerate(out.shapes):
            shapes.append((h, w))
            offsets.append(start + int(out.pixel_offsets[k + 1]) * channels)
        keep.append((out, sub))
    state['d2h'].synchronize()
    for s in state['work']:
        s.synchronize()
    current.wait_stream(state['h2d'])
    if use_thread:
        thread.join()
    return host_out, shapes, offsets


# =============================================================================================
# Photometric chain over a ragged batch
# =============================================================================================
from .mechanism.distortion.photometric import blur as _blur  # noqa: E402
from .mechanism.distortion.photometric import color as _color  # noqa: E402
from .mechanism.distortion.photometric.opt import OutOfBoundBehavior, make_op  # noqa: E402

_ALL_BITS = {1: 0b1, 3: 0b111, 4: 0b1111}


def _bits(channels: int, selected):
    if not selected:
        return _ALL_BITS[channels]
    bits = 0
    for c in selected:
        if c < 0 or c >= channels:
            raise IndexError(f'channel {c} out of range')
        bits |= 1 << c
    return bits


def _stage_ops(name: str, config, channels: int, stats):
    """Op list of one photometric stage for one RGB / GRAYSCALE uint8 page (the per-page half of
    photometric/color.py); `stats` = (sums, mins, maxs, n_pixels) of the stage's input when the
    stage needs them.  Returns [] for a NOP."""
    if name == 'mean_shift':
        config = dyn_structure(config, _color.MeanShiftConfig)
        if config.delta == 0:
            return []
        return [make_op(nv.OP_MEAN_SHIFT, i0=config.delta,
                        i1=-1 if config.threshold is None else config.threshold,
                        i2=_bits(channels, config.channels),
                        i3=1 if config.oob_behavior == OutOfBoundBehavior.CYCLE else 0)]
    if name == 'color_shift':
        config = dyn_structure(config, _color.ColorShiftConfig)
        return [make_op(nv.OP_HUE_SHIFT_RGB, i0=config.delta)]
    if name == 'brightness_shift':
        config = dyn_structure(config, _color.BrightnessShiftConfig)
        from .element import ImageMode
        via_hsv = 1 if config.intermediate_image_mode == ImageMode.HSV else 0
        return [make_op(nv.OP_LIGHT_SHIFT_RGB, i0=config.delta, i1=via_hsv)]
    if name == 'complement':
        config = dyn_structure(config, _color.ComplementConfig)
        return [make_op(nv.OP_COMPLEMENT, i1=-1 if config.threshold is None else config.threshold,
                        i2=_bits(channels, config.channels), i3=int(config.enable_threshold_lte))]
    if name == 'posterization':
        config = dyn_structure(config, _color.PosterizationConfig)
        if config.num_bits == 0:
            return []
        return [make_op(nv.OP_POSTERIZE, i0=(0xFF >> config.num_bits) << config.num_bits,
                        i2=_bits(channels, config.channels))]
    if name == 'color_balance':
        config = dyn_structure(config, _color.ColorBalanceConfig)
        if channels == 1:
            return []
        return [make_op(nv.OP_COLOR_BALANCE, f0=np.float32(config.ratio),
                        f1=np.float32(1 - config.ratio))]
    if name == 'std_shift':
        config = dyn_structure(config, _color.StdShiftConfig)
        sums, _, _, count = stats
        means = [np.float32(float(s) / count) for s in sums[:min(channels, 3)]]
        sub = [np.float32(m) * np.float32(config.scale - 1) for m in means] + [0.0, 0.0]
        return [make_op(nv.OP_STD_SHIFT, i2=_bits(channels, config.channels),
                        f0=np.float32(config.scale), f1=sub[0], f2=sub[1] if channels > 1 else 0.0,
                        f3=sub[2] if channels > 2 else 0.0)]
    if name == 'boundary_equalization':
        config = dyn_structure(config, _color.BoundaryEqualizationConfig)
        _, mins, maxs, _ = stats
        bits = _bits(channels, config.channels)
        mn, sc, active = [0.0] * 3, [0.0] * 3, 0
        for c in range(min(channels, 3)):
            if not (bits >> c) & 1:
                continue
            delta = np.float32(maxs[c]) - np.float32(mins[c])
            if delta > 0:
                active |= 1 << c
                mn[c] = float(mins[c])
                sc[c] = np.float32(255.0) / delta
        if not active:
            return []
        return [make_op(nv.OP_BOUNDARY_EQ, i2=active, f0=mn[0], f1=mn[1], f2=mn[2], g0=sc[0],
                        g1=sc[1], g2=sc[2])]
    if name == 'line_streak':
        from .mechanism.distortion.photometric import streak as _streak
        config = dyn_structure(config, _streak.LineStreakConfig)
        alpha = _streak._check_alpha(config.alpha)
        if alpha == 0.0 or not (config.enable_vert or config.enable_hori):
            return []
        color = list(config.color) if channels > 1 else [config.color[0]]
        color = [float(np.uint8(c)) for c in color] + [0.0] * 4
        return [make_op(nv.OP_LINE_STREAK, i0=config.thickness, i1=config.gap,
                        i2=(config.dash_thickness & 0xFFFF) | ((config.dash_gap & 0xFFFF) << 16),
                        i3=int(config.enable_vert) | (int(config.enable_hori) << 1), f0=alpha,
                        f1=color[0], f2=color[1], f3=color[2], g0=color[3])]
    raise NotImplementedError(f'{name} has no batched form')


def _noise_op(name: str, config, seed: int):
    """VKB_OP_NOISE record of one page (photometric/noise.py), Philox mode."""
    from .mechanism.distortion.photometric import noise as _noise
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    lo, hi = seed & 0xFFFFFFFF, seed >> 32
    lo = lo - (1 << 32) if lo >= (1 << 31) else lo
    hi = hi - (1 << 32) if hi >= (1 << 31) else hi
    if name == 'gaussion_noise':
        config = dyn_structure(config, _noise.GaussionNoiseConfig)
        return make_op(nv.OP_NOISE, i0=nv.NOISE_GAUSSIAN, i1=lo, i2=hi, f0=config.std)
    if name == 'poisson_noise':
        return make_op(nv.OP_NOISE, i0=nv.NOISE_POISSON, i1=lo, i2=hi)
    if name == 'impulse_noise':
        config = dyn_structure(config, _noise.ImpulseNoiseConfig)
        return make_op(nv.OP_NOISE, i0=nv.NOISE_IMPULSE, i1=lo, i2=hi, f0=config.prob_salt,
                       f1=config.prob_pepper)
    config = dyn_structure(config, _noise.SpeckleNoiseConfig)
    return make_op(nv.OP_NOISE, i0=nv.NOISE_SPECKLE, i1=lo, i2=hi, f0=config.std)


_STATS_STAGES = ('std_shift', 'boundary_equalization')
_NOISE_STAGES = ('gaussion_noise', 'poisson_noise', 'impulse_noise', 'speckle_noise')


class PhotometricBatch:
    """A chain of photometric stages over a ragged batch of uint8 pages (RGB or GRAYSCALE), every
    page with its own configs: `stages` = [(name, configs)], name one of gaussian_blur,
    mean_shift, color_shift, brightness_shift, std_shift, boundary_equalization, complement,
    posterization, color_balance, line_streak, and the four noises as (name, configs, seeds) with
    one Philox seed per page (what the per-page op draws from its rng).

    Consecutive stages are folded into passes of the fused kernel (`vkb_photo_chain_batched`):
    a pass is an optional Gaussian blur followed by up to 8 per-pixel ops; a new pass starts at
    every blur and at every stage that needs statistics of its input."""

    def __init__(self, shapes: Sequence[Tuple[int, int]], channels: int, stages):
        self.shapes = [(int(h), int(w)) for h, w in shapes]
        self.n = len(self.shapes)
        self.channels = channels
        self.stages = []
        for stage in stages:
            name, configs = stage[0], list(stage[1])
            seeds = list(stage[2]) if len(stage) > 2 else None
            if len(configs) != self.n or (seeds is not None and len(seeds) != self.n):
                raise ValueError(f'stage {name}: one config (and seed) per page expected')
            if name in _NOISE_STAGES and seeds is None:
                raise ValueError(f'stage {name}: (name, configs, seeds) expected')
            self.stages.append((name, configs, seeds))
        sizes = np.asarray([h * w for h, w in self.shapes], dtype=np.int64)
        self.pixel_offsets = np.concatenate([[0], np.cumsum(sizes)])
        self.launches = 0
        self.launch_events = None  # set to a list to collect (start, end) CUDA events per pass

    def _new_pass(self):
        rec = np.zeros(self.n, dtype=nv.PHOTO_PAGE_DTYPE)
        rec['h'] = [s[0] for s in self.shapes]
        rec['w'] = [s[1] for s in self.shapes]
        return rec

    def _launch(self, rec, src, dst):
        base = self.pixel_offsets[:-1].astype(np.uint64) * np.uint64(self.channels)
        rec['src'] = np.uint64(src.data_ptr()) + base
        rec['dst'] = np.uint64(dst.data_ptr()) + base
        rec_dev = dv.upload_structs(rec)
        if self.launch_events is not None:
            t = dv.torch()
            ev0, ev1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            ev0.record()
        nv.check(nv.lib().vkb_photo_chain_batched(dv.ptr(rec_dev), rec.ctypes.data_as(ctypes.c_void_p),
                                                  self.n, self.channels, dv.stream_ptr()),
                 'vkb_photo_chain_batched')
        if self.launch_events is not None:
            ev1.record()
            self.launch_events.append((ev0, ev1))
        self.launches += 1
        return rec_dev

    def _stats(self, arena):
        rec = self._new_pass()
        base = self.pixel_offsets[:-1].astype(np.uint64) * np.uint64(self.channels)
        rec['src'] = np.uint64(arena.data_ptr()) + base
        rec['dst'] = rec['src']
        rec_dev = dv.upload_structs(rec)
        out = dv.empty((self.n, 48), np.uint8)
        nv.check(nv.lib().vkb_channel_stats_batched(dv.ptr(rec_dev), self.n, min(self.channels, 3),
                                                    dv.ptr(out), dv.stream_ptr()),
                 'vkb_channel_stats_batched')
        self.launches += 2
        raw = dv.to_host(out)
        sums = raw[:, :24].copy().view(np.uint64)
        mins = raw[:, 24:36].copy().view(np.uint32)
        maxs = raw[:, 36:48].copy().view(np.uint32)
        return sums, mins, maxs

    def run(self, arena, scratch=None):
        """arena: flat uint8 CUDA tensor holding the pages back to back (pixel_offsets order).
        Returns the tensor that holds the result (arena itself or `scratch`)."""
        t = dv.require_cuda()
        cur = arena
        other = scratch
        rec = self._new_pass()
        dirty = False

        def flush():
            nonlocal cur, other, rec, dirty
            if not dirty:
                return
            if (rec['blur_radius'] > 0).any():
                if other is None:
                    other = t.empty_like(cur)
                self._launch(rec, cur, other)
                cur, other = other, cur
            else:
                self._launch(rec, cur, cur)
            rec = self._new_pass()
            dirty = False

        for name, configs, seeds in self.stages:
            if name in _NOISE_STAGES:
                flush()
                for i, (config, seed) in enumerate(zip(configs, seeds)):
                    op = _noise_op(name, config, seed)
                    rec['ops'][i, 0] = np.frombuffer(bytes(op), dtype=nv.COLOR_OP_DTYPE)[0]
                    rec['n_ops'][i] = 1
                base = self.pixel_offsets[:-1].astype(np.uint64) * np.uint64(self.channels)
                rec['src'] = np.uint64(cur.data_ptr()) + base
                rec['dst'] = rec['src']
                rec_dev = dv.upload_structs(rec)
                nv.check(nv.lib().vkb_noise_philox_batched(dv.ptr(rec_dev), self.n, self.channels,
                                                           64, dv.stream_ptr()),
                         'vkb_noise_philox_batched')
                self.launches += 1
                rec = self._new_pass()
                continue
            if name == 'gaussian_blur':
                flush()
                sigmas = [dyn_structure(config, _blur.GaussianBlurConfig).sigma for config in configs]
                ksizes, taps = _blur.gaussian_kernels_u8(sigmas)  # all pages at once
                if (ksizes > 17).any():
                    raise NotImplementedError('gaussian_blur kernels wider than 17 taps')
                rec['blur_radius'] = ksizes // 2
                rec['blur_taps'] = taps
                dirty = True
                continue
            stats = None
            if name in _STATS_STAGES:
                flush()
                stats = self._stats(cur)
            elif int(rec['n_ops'].max()) >= nv.MAX_COLOR_OPS:
                flush()
            for i, config in enumerate(configs):
                page_stats = None
                if stats is not None:
                    h, w = self.shapes[i]
                    page_stats = (stats[0][i], stats[1][i], stats[2][i], h * w)
                for op in _stage_ops(name, config, self.channels, page_stats):
                    k = int(rec['n_ops'][i])
                    rec['ops'][i, k] = np.frombuffer(bytes(op), dtype=nv.COLOR_OP_DTYPE)[0]
                    rec['n_ops'][i] = k + 1
                    dirty = True
        flush()
        return cur


def distort_chain(op_names: Sequence[str], configs: Sequence, shape: Tuple[int, int], images,
                  photometric_stages):
    """Geometric grid op per page followed by a photometric chain, all batched:
    e.g. similarity_mls -> gaussian_blur -> color_shift (BASELINE config 3).
    images: (B, H, W, C) uint8 CUDA tensor.  Returns (BatchOutput, PhotometricBatch)."""
    engine = GeometricBatch(op_names, configs, shape)
    out = engine.run(images)
    photo = PhotometricBatch(out.shapes, out.channels, photometric_stages)
    out.image_arena = photo.run(out.image_arena)
    return out, photo
