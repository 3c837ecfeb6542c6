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


def grid_page_record(op_name: str, config, shape: Tuple[int, int]):
    """(page record, keepalive) for one page of a grid-based op."""
    height, width = shape
    if op_name == 'similarity_mls':
        config = dyn_structure(config, _mls.SimilarityMlsConfig)
        return _mls.similarity_mls_page(config, shape)
    config_cls, builder = _PAGE_BUILDERS[op_name]
    config = dyn_structure(config, config_cls)
    rec = builder(config, shape)
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
        records = []
        keepalive = []
        for op_name, config in zip(op_names, configs):
            rec, keep = grid_page_record(op_name, config, self.shape)
            records.append(rec)
            if keep is not None:
                keepalive.append(keep)
        self.pages = np.stack(records).astype(nv.GRID_PAGE_DTYPE)
        self.keepalive = keepalive
        self.n = len(records)
        self.plan: Optional[GridBatch] = None

    def plan_batch(self):
        """Phase 1 + 2a: lattices, result shapes (one D2H), cell homographies, masks, bins."""
        self.plan = GridBatch(self.pages, keepalive=self.keepalive)
        self.plan.build()
        return self.plan

    def run(self, images=None, masks=None, score_maps=None, replan: bool = True) -> BatchOutput:
        """images: (B, H, W, C) uint8, masks: (B, H, W) uint8, score_maps: (B, H, W) float32 --
        CUDA tensors (any subset).  Returns ragged outputs in fresh arenas."""
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
        plan.remap(planes)
        return BatchOutput(shapes, channels, image_arena, mask_arena, score_arena, offsets)

    def algorithmic_bytes(self, channels: int = 3, with_mask: bool = False,
                          with_score: bool = False) -> int:
        """Compulsory traffic of the batch: every source pixel read once, every destination
        pixel written once."""
        per_px = channels + (1 if with_mask else 0) + (4 if with_score else 0)
        height, width = self.shape
        dst = sum(h * w for h, w in (self.plan.result_shape(i) for i in range(self.n)))
        return per_px * (self.n * height * width + dst)


_PIPELINE_STREAMS = {}


def _pipeline_streams():
    """Two long-lived side streams per device (the caching allocator keeps one pool per stream,
    so fresh streams per call would mean fresh cudaMallocs per call)."""
    t = dv.require_cuda()
    key = t.cuda.current_device()
    if key not in _PIPELINE_STREAMS:
        _PIPELINE_STREAMS[key] = [t.cuda.Stream(), t.cuda.Stream()]
    return _PIPELINE_STREAMS[key]


def distort_pages_host(op_names: Sequence[str], configs: Sequence, shape: Tuple[int, int],
                       host_images, host_out=None, chunk_pages: int = 64,
                       use_thread: bool = True):
    """Host buffers in, host buffers out -- the end-to-end form of the batch engine.

    `host_images`: pinned (B, H, W, C) uint8 CPU tensor; the distorted pages land back to back in
    the pinned flat uint8 tensor `host_out` (allocated when None).  The batch is cut into chunks
    that alternate between two CUDA streams, so the H2D copy of chunk i+1, the kernels of chunk i
    and the D2H copy of chunk i-1 overlap, while a helper thread turns the configs of the next
    chunk into parameter blocks (Rodrigues, translation, curve / line constants).
    Returns (host_out, shapes, byte_offsets)."""
    import queue
    import threading
    t = dv.require_cuda()
    n = len(op_names)
    bounds = [(a, min(a + chunk_pages, n)) for a in range(0, n, chunk_pages)]
    ready: 'queue.Queue' = queue.Queue(maxsize=2)

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
    streams = _pipeline_streams()
    for s in streams:
        s.wait_stream(t.cuda.current_stream())
    shapes, offsets = [], [0]
    keep = []
    for i, (a, b) in enumerate(bounds):
        sub = ready.get() if use_thread else GeometricBatch(op_names[a:b], configs[a:b], shape)
        if isinstance(sub, BaseException):
            raise sub
        with t.cuda.stream(streams[i % 2]):
            dev_in = host_images[a:b].to(dv.device(), non_blocking=True)
            out = sub.run(dev_in)
            n_bytes = int(out.image_arena.numel())
            start = offsets[-1]
            if start + n_bytes > host_out.numel():
                raise ValueError('host_out is too small for the distorted pages')
            host_out[start:start + n_bytes].copy_(out.image_arena, non_blocking=True)
            for k, (h, w) in enumerate(out.shapes):
                shapes.append((h, w))
                offsets.append(start + int(out.pixel_offsets[k + 1]) * channels)
            keep.append((dev_in, out, sub))
        if len(keep) > 2:
            # chunk i-2 ran on this stream's sibling two iterations ago: its buffers may be
            # recycled once that stream has drained
            streams[(i + 1) % 2].synchronize()
            keep.pop(0)
    for s in streams:
        s.synchronize()
    if use_thread:
        thread.join()
    return host_out, shapes, offsets
