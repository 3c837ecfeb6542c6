"""Batched page engine: many independent pages through one set of launches.

`Distortion.distort` handles one page per call, which is launch- and Python-bound (a 1024^2 RGB
page is ~1 us of HBM traffic).  This module keeps the same ops and configs but runs a whole
batch of pages per launch: per-page parameter blocks in one device array, ragged outputs in one
arena, one small D2H per batch for the result shapes.

    engine = GeometricBatch.from_configs(distortions, configs, shape)
    out = engine.run(images_dev)          # images_dev: (B, H, W, 3) uint8 CUDA tensor
    out.image(i)                          # (H'_i, W'_i, 3) view into the output arena
"""
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


def grid_page_record(op_name: str, config, shape: Tuple[int, int], out=None):
    """(page record, keepalive) for one page of a grid-based op.  `out`: zeroed element of a
    GRID_PAGE_DTYPE array to fill in place."""
    height, width = shape
    if op_name == 'similarity_mls':
        config = dyn_structure(config, _mls.SimilarityMlsConfig)
        return _mls.similarity_mls_page(config, shape, out)
    config_cls, builder = _PAGE_BUILDERS[op_name]
    config = dyn_structure(config, config_cls)
    rec = builder(config, shape, out)
    camera_model_config = _camera.complete_camera_model_config(height, width,
                                                               config.camera_model_config)
    _camera.fill_camera_model(rec, camera_model_config)
    return rec, None


class BatchOutput:
    """Ragged outputs of one batch: flat arenas + per-page offsets."""

    def __init__(self, shapes, channels, image_arena, mask_arena, score_arena, pixel_offsets):
        self.shapes = shapes
        self.channels = channels
        self.image_arena = image_arena
        self.mask_arena = mask_arena
        self.score_arena = score_arena
        self.pixel_offsets = pixel_offsets

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


class GeometricBatch:
    """A batch of same-size pages, each with its own grid-op config."""

    def __init__(self, op_names: Sequence[str], configs: Sequence, shape: Tuple[int, int]):
        self.shape = tuple(shape)
        keepalive = []
        self.n = len(op_names)
        self.pages = np.zeros(self.n, dtype=nv.GRID_PAGE_DTYPE)
        for i, (op_name, config) in enumerate(zip(op_names, configs)):
            _, keep = grid_page_record(op_name, config, self.shape, out=self.pages[i])
            if keep is not None:
                keepalive.append(keep)
        self.keepalive = keepalive
        self.plan: Optional[GridBatch] = None

    def plan_batch(self):
        """Phase 1 + 2a: lattices, result shapes (one D2H), cell homographies, masks, bins."""
        self.plan = GridBatch(self.pages, keepalive=self.keepalive)
        self.plan.build()
        return self.plan

    def run(self, images=None, masks=None, score_maps=None, replan: bool = True,
            launch_events=None) -> BatchOutput:
        """images: (B, H, W, C) uint8, masks: (B, H, W) uint8, score_maps: (B, H, W) float32 --
        CUDA tensors (any subset).  Returns ragged outputs in fresh arenas.
        `launch_events`: optional list; a (start, end) pair of CUDA events recorded immediately
        around the fused remap launch is appended (kernel time without host work)."""
        if replan or self.plan is None:
            self.plan_batch()
        plan = self.plan
        height, width = self.shape
        shapes = [plan.result_shape(i) for i in range(self.n)]
        pixels = np.asarray([h * w for h, w in shapes], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(pixels)])
        total = int(offsets[-1])
        planes = np.zeros(self.n, dtype=nv.PLANES_DTYPE)
        planes['src_h'], planes['src_w'] = height, width
        planes['dst_h'] = [s[0] for s in shapes]
        planes['dst_w'] = [s[1] for s in shapes]
        image_arena = mask_arena = score_arena = None
        channels = 0
        if images is not None:
            channels = 1 if images.dim() == 3 else int(images.shape[3])
            image_arena = dv.empty((total * channels,), np.uint8)
            base = images.data_ptr()
            planes['src_image'] = base + np.arange(self.n, dtype=np.uint64) * np.uint64(
                height * width * channels)
            planes['dst_image'] = image_arena.data_ptr() + (offsets[:-1] * channels).astype(
                np.uint64)
            planes['image_channels'] = channels
        if masks is not None:
            mask_arena = dv.empty((total,), np.uint8)
            planes['src_mask'] = masks.data_ptr() + np.arange(self.n, dtype=np.uint64) * np.uint64(
                height * width)
            planes['dst_mask'] = mask_arena.data_ptr() + offsets[:-1].astype(np.uint64)
        if score_maps is not None:
            score_arena = dv.empty((total,), np.float32)
            planes['src_score'] = score_maps.data_ptr() + np.arange(
                self.n, dtype=np.uint64) * np.uint64(height * width * 4)
            planes['dst_score'] = score_arena.data_ptr() + (offsets[:-1] * 4).astype(np.uint64)
        plan.remap(planes, launch_events=launch_events)
        return BatchOutput(shapes, channels, image_arena, mask_arena, score_arena, offsets)

    def algorithmic_bytes(self, channels: int = 3, with_mask: bool = False,
                          with_score: bool = False) -> int:
        """Compulsory traffic of the batch: every source pixel read once, every destination
        pixel written once."""
        per_px = channels + (1 if with_mask else 0) + (4 if with_score else 0)
        height, width = self.shape
        dst = sum(h * w for h, w in (self.plan.result_shape(i) for i in range(self.n)))
        return per_px * (self.n * height * width + dst)


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
    first = min(n, max(1, chunk_pages // 4))
    bounds = [(0, first)] + [(a, min(a + chunk_pages, n)) for a in range(first, n, chunk_pages)]
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
        n_bytes = int(out.image_arena.numel())
        start = offsets[-1]
        if start + n_bytes > host_out.numel():
            raise ValueError('host_out is too small for the distorted pages')
        with t.cuda.stream(state['d2h']):
            state['d2h'].wait_event(done)
            host_out[start:start + n_bytes].copy_(out.image_arena, non_blocking=True)
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
