"""RandomDistortion over a BATCH of pages (vkit/mechanism/distortion_policy/random_distortion.py:350-392
as called by pipeline/text_detection/page_distortion.py:316-484, one page per call there).

A page's chain is drawn on the host first, with the page's own generator and in exactly the
reference's order -- the draws depend on the page shape only (the photometric ops run before the
geometric one, and the post-rotate generator ignores the shape), never on pixels.  The stages then
run bucketed by op:

  photometric   per page, device resident (0 - 2 ops per page, each a few launches);
  geometric     all grid ops of the batch through ONE GeometricBatch (mixed camera_* / MLS pages),
                all affine ops through ONE AffineBatch, image + mask fused;
  post rotate   ONE AffineBatch over the (ragged) results;
  points        corner points, points and polygon vertices of all pages through one launch per
                stage (vkb_grid_points_batched / vkb_affine_points_batched), kept as arrays;
  trim          crop to the bounding box of everything tracked (random_distortion.py:266-348).

Results are `BatchPageResult`s: device tensors for image / mask, NumPy arrays for the coordinates
(`to_distortion_result()` builds the reference's containers when somebody wants objects).
"""
import ctypes
from typing import List, Optional, Sequence

import attrs
import numpy as np

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.batch import AffineBatch, GeometricBatch, _affine_states
from vkit_b200.element import Image, Mask, Point, PointTuple, Polygon
from vkit_b200.mechanism.distortion.interface import DistortionResult

from .random_distortion import RandomDistortion
from .opt import LEVEL_MAX, LEVEL_MIN

_GRID_OPS = ('camera_plane_only', 'camera_cubic_curve', 'camera_plane_line_fold',
             'camera_plane_line_curve', 'similarity_mls')


@attrs.define
class PageChain:
    photometric: list = attrs.field(factory=list)   # (policy, level, config)
    geometric: Optional[tuple] = None                # (policy, level, config)
    post_rotate: Optional[tuple] = None
    inject_corner_points: bool = False
    names: list = attrs.field(factory=list)
    levels: list = attrs.field(factory=list)
    configs: list = attrs.field(factory=list)


@attrs.define
class BatchPageResult:
    shape: tuple
    image: object = None        # CUDA tensor (H, W, C) uint8
    mask: object = None         # CUDA tensor (H, W) uint8
    points: Optional[np.ndarray] = None          # (N, 2) float64 smooth (x, y)
    polygons: Optional[List[np.ndarray]] = None  # per polygon (K, 2) float64 smooth (x, y)
    chain: Optional[PageChain] = None

    def to_distortion_result(self) -> DistortionResult:
        result = DistortionResult(shape=self.shape)
        if self.image is not None:
            result.image = Image(mat=self.image)
        if self.mask is not None:
            result.mask = Mask(mat=self.mask)
        if self.points is not None:
            result.points = PointTuple(Point.create(y=float(y), x=float(x)) for x, y in self.points)
        if self.polygons is not None:
            result.polygons = [Polygon.create(points=[Point.create(y=float(y), x=float(x))
                                                      for x, y in poly]) for poly in self.polygons]
        return result


def _py_round(values: np.ndarray) -> np.ndarray:
    return np.rint(values).astype(np.int64)  # Python round(): half to even


class RandomDistortionBatch:

    def __init__(self, random_distortion: RandomDistortion):
        self.rd = random_distortion

    # ---- host: the chains ----------------------------------------------------------------
    def sample_chain(self, rng, shape) -> PageChain:
        """The draws of RandomDistortion.distort for one page, nothing executed."""
        chain = PageChain()
        for stage in self.rd.stages:
            cfg = stage.config
            if rng.random() > cfg.prob_enable:
                continue
            if cfg.inject_corner_points:
                chain.inject_corner_points = True
            level_min, level_max = self.rd.level_min, self.rd.level_max
            if cfg.force_sample_level_in_full_range:
                level_min, level_max = LEVEL_MIN, LEVEL_MAX
            for policy in stage.sample_distortion_policies(rng):
                level = int(rng.integers(level_min, level_max + 1))
                generator = policy.config_generator_cls(policy.config_for_config_generator, level)
                # config + rng-state capture exactly like Distortion.prepare_config_and_rng
                config, _ = policy.distortion.prepare_config_and_rng(generator, shape, rng)
                item = (policy, level, config)
                if not policy.distortion.is_geometric:
                    chain.photometric.append(item)
                elif cfg.force_sample_level_in_full_range:
                    chain.post_rotate = item
                else:
                    chain.geometric = item
                chain.names.append(policy.name)
                chain.levels.append(level)
                chain.configs.append(config)
        return chain

    # ---- device: coordinates -------------------------------------------------------------
    @staticmethod
    def _corner_points(shape):
        height, width = shape
        step = min(height // 4, width // 4)
        assert step > 0
        ys = list(range(0, height, step))
        if ys[-1] < height - 1:
            ys.append(height - 1)
        xs = list(range(0, width, step))
        if xs[0] == 0:
            xs.pop(0)
        if xs[-1] == width - 1:
            xs.pop()
        pts = [(x, y) for x in (0, width - 1) for y in ys] + [(x, y) for y in (0, height - 1) for x in xs]
        return np.asarray(pts, dtype=np.float64)

    @staticmethod
    def _clip(xy: np.ndarray, shape):
        """Point.to_clipped_point: points whose ROUNDED twin leaves the page are clipped."""
        if xy.shape[0] == 0:
            return xy
        height, width = shape
        r = _py_round(xy)
        outside = (r[:, 0] < 0) | (r[:, 0] >= width) | (r[:, 1] < 0) | (r[:, 1] >= height)
        if outside.any():
            xy = xy.copy()
            xy[outside, 0] = np.clip(xy[outside, 0], 0, width - 1)
            xy[outside, 1] = np.clip(xy[outside, 1], 0, height - 1)
        return xy

    def _move_points(self, kind, pages, sets, engine, shapes_out):
        """One launch for the coordinates of all `pages` of one geometric bucket.
        sets[i]: {'corner': xy, 'points': xy, 'poly': xy} (float64 smooth).  In place."""
        lib = nv.lib()
        chunks, owners = [], []
        for slot, i in enumerate(pages):
            for key, xy in sets[i].items():
                if xy is not None and xy.shape[0]:
                    chunks.append((slot, i, key, xy))
        if not chunks:
            return
        total = sum(c[3].shape[0] for c in chunks)
        xy_in = np.empty((total, 2), dtype=np.float64)
        begin = 0
        if kind == 'grid':
            page_cell = np.empty((total, 3), dtype=np.int32)
            plan = engine.plan
            for slot, i, key, xy in chunks:
                end = begin + xy.shape[0]
                g = int(plan.pages['grid_size'][slot])
                r = _py_round(xy)
                page_cell[begin:end, 0] = slot
                page_cell[begin:end, 1] = r[:, 1] // g
                page_cell[begin:end, 2] = r[:, 0] // g
                rows, cols = int(plan.pages['rows'][slot]), int(plan.pages['cols'][slot])
                if (page_cell[begin:end, 1:] < 0).any() or (page_cell[begin:end, 1] >= rows - 1).any() \
                        or (page_cell[begin:end, 2] >= cols - 1).any():
                    raise IndexError('point outside the source lattice')  # the reference raises too
                xy_in[begin:end] = xy
                begin = end
            xy_dev, pc_dev = dv.to_device(xy_in), dv.to_device(page_cell)
            out = dv.empty((total, 2), np.float64)
            nv.check(lib.vkb_grid_points_batched(dv.ptr(plan.pages_dev), dv.ptr(plan.hfwd), plan.c_max,
                                                 dv.ptr(xy_dev), dv.ptr(pc_dev), dv.ptr(out), total,
                                                 dv.stream_ptr()), 'vkb_grid_points_batched')
        else:
            page_of = np.empty((total,), dtype=np.int32)
            for slot, i, key, xy in chunks:
                end = begin + xy.shape[0]
                page_of[begin:end] = slot
                # PointTuple.to_smooth_np_array hands out the ROUNDED coordinates (point.py:251-252);
                # polygons travel as PointLists (smooth); both as float32
                src = _py_round(xy).astype(np.float64) if key != 'poly' else xy
                xy_in[begin:end] = src.astype(np.float32).astype(np.float64)
                begin = end
            mats = np.zeros((len(pages), 9), dtype=np.float64)
            rows_f32 = np.zeros((len(pages),), dtype=np.int32)
            for slot in range(len(pages)):
                mat = engine.forward[slot]
                if mat is None:  # identity (nop config)
                    mats[slot, [0, 4]] = 1.0
                    rows_f32[slot] = 2
                else:
                    rows = int(mat.shape[0])
                    mats[slot, :rows * 3] = np.asarray(mat, dtype=np.float64).reshape(-1)
                    if rows == 3:
                        rows_f32[slot] = 3
                    else:
                        rows_f32[slot] = 2 | (4 if mat.dtype == np.float32 else 0)
            xy_dev, po_dev = dv.to_device(xy_in), dv.to_device(page_of)
            mats_dev, rf_dev = dv.to_device(mats), dv.to_device(rows_f32)
            out = dv.empty((total, 2), np.float64)
            nv.check(lib.vkb_affine_points_batched(dv.ptr(mats_dev), dv.ptr(rf_dev), dv.ptr(po_dev),
                                                   dv.ptr(xy_dev), dv.ptr(out), total,
                                                   dv.stream_ptr()), 'vkb_affine_points_batched')
        moved = dv.to_host(out)
        begin = 0
        for slot, i, key, xy in chunks:
            end = begin + xy.shape[0]
            res = moved[begin:end]
            if kind != 'grid' and engine.forward[slot] is None:
                res = xy  # nop: the reference returns the points unchanged
            sets[i][key] = self._clip(np.array(res, dtype=np.float64), shapes_out[slot])
            begin = end

    # ---- the batch ---------------------------------------------------------------------------
    def distort(self, rngs: Sequence, images, masks=None, points: Optional[Sequence] = None,
                polygons: Optional[Sequence] = None) -> List[BatchPageResult]:
        """images: (B, H, W, C) uint8 CUDA tensor; masks: (B, H, W) uint8 CUDA tensor or None;
        points[i]: (N, 2) float (x, y) or None; polygons[i]: list of (K, 2) arrays or None."""
        t = dv.require_cuda()
        n = int(images.shape[0])
        shape = (int(images.shape[1]), int(images.shape[2]))
        chains = [self.sample_chain(rng, shape) for rng in rngs]

        cur_img = [images[i] for i in range(n)]
        cur_mask = [masks[i] for i in range(n)] if masks is not None else [None] * n
        cur_shape = [shape] * n
        sets = []
        for i in range(n):
            poly = None
            if polygons is not None and polygons[i] is not None and len(polygons[i]):
                poly = np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, 2)
                                       for p in polygons[i]])
            sets.append({
                'corner': self._corner_points(shape) if chains[i].inject_corner_points else None,
                'points': (np.asarray(points[i], dtype=np.float64).reshape(-1, 2)
                           if points is not None and points[i] is not None else None),
                'poly': poly,
            })

        # ---- photometric stage: per page, device resident ------------------------------------
        for i, chain in enumerate(chains):
            if not chain.photometric:
                continue
            image = Image(mat=cur_img[i])
            for policy, _, config in chain.photometric:
                image = policy.distortion.distort_image(config, image)
            cur_img[i] = image.dev

        # ---- geometric stage, bucketed -----------------------------------------------------------
        def run_bucket(kind, pages, items):
            if not pages:
                return
            names = [it[0].name for it in items]
            configs = [it[2] for it in items]
            shapes_in = [cur_shape[i] for i in pages]
            imgs = [cur_img[i] for i in pages]
            msks = [cur_mask[i] for i in pages] if masks is not None else None
            if kind == 'grid':
                engine = GeometricBatch(names, configs, shapes_in)
                engine.plan_batch()
                engine.plan.build(need_forward=True)
                out = engine.run(imgs, msks, replan=False, channels=int(images.shape[3]))
            else:
                engine = AffineBatch(names, configs, shapes_in)
                out = engine.run(imgs, msks, channels=int(images.shape[3]))
            shapes_out = list(out.shapes)
            self._move_points(kind, pages, sets, engine, shapes_out)
            for slot, i in enumerate(pages):
                cur_img[i] = out.image(slot)
                if masks is not None:
                    cur_mask[i] = out.mask(slot)
                cur_shape[i] = tuple(shapes_out[slot])

        grid_pages = [i for i, c in enumerate(chains) if c.geometric and c.geometric[0].name in _GRID_OPS]
        affine_pages = [i for i, c in enumerate(chains)
                        if c.geometric and c.geometric[0].name not in _GRID_OPS]
        run_bucket('grid', grid_pages, [chains[i].geometric for i in grid_pages])
        run_bucket('affine', affine_pages, [chains[i].geometric for i in affine_pages])
        rot_pages = [i for i, c in enumerate(chains) if c.post_rotate]
        run_bucket('affine', rot_pages, [chains[i].post_rotate for i in rot_pages])

        # ---- trim + results ----------------------------------------------------------------------
        results = []
        for i in range(n):
            img, msk, shp = cur_img[i], cur_mask[i], cur_shape[i]
            s = sets[i]
            if s['corner'] is not None:
                tracked = [a for a in (s['corner'], s['points'], s['poly']) if a is not None and a.shape[0]]
                r = _py_round(np.concatenate(tracked))
                left, right = int(r[:, 0].min()), int(r[:, 0].max())
                up, down = int(r[:, 1].min()), int(r[:, 1].max())
                height, width = shp
                pad_up, pad_down, pad_left, pad_right = up, height - 1 - down, left, width - 1 - right
                assert min(pad_up, pad_down, pad_left, pad_right) >= -1  # rounding slack
                if max(pad_up, pad_down, pad_left, pad_right) > 0:
                    up, down = max(0, up), min(height - 1, down)
                    left, right = max(0, left), min(width - 1, right)
                    pad_up, pad_left = max(0, pad_up), max(0, pad_left)
                    img = img[up:down + 1, left:right + 1].contiguous()
                    if msk is not None:
                        msk = msk[up:down + 1, left:right + 1].contiguous()
                    shp = (down - up + 1, right - left + 1)
                    for key in ('points', 'poly'):
                        if s[key] is not None:
                            s[key] = s[key] - np.asarray([pad_left, pad_up], dtype=np.float64)
            polys = None
            if s['poly'] is not None:
                polys, begin = [], 0
                for p in polygons[i]:
                    k = int(np.asarray(p).reshape(-1, 2).shape[0])
                    polys.append(s['poly'][begin:begin + k])
                    begin += k
            results.append(BatchPageResult(shape=shp, image=img, mask=msk, points=s['points'],
                                           polygons=polys, chain=chains[i]))
        return results
