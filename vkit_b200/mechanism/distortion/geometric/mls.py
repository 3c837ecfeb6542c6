"""similarity_mls: Moving Least Squares with similarity transforms
(vkit/mechanism/distortion/geometric/mls.py; Schaefer et al. 2006).

The reference evaluates the projector once per lattice point in Python (4 900 calls per
1024^2 page); `vkb_grid_project` gives one warp to each lattice point, lanes over the handles
and shuffle reductions for the weight normalisation, centroids and mu.
"""
from typing import Optional, Tuple

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.element import PointTuple

from ..interface import DistortionConfig
from ._gridcore import new_grid_page
from .grid_rendering import DistortionImageGridBased, DistortionStateImageGridBased


@attrs.define
class SimilarityMlsConfig(DistortionConfig):
    src_handle_points: PointTuple
    dst_handle_points: PointTuple
    grid_size: int
    resize_as_src: bool = False


def similarity_mls_page(config: SimilarityMlsConfig, shape: Tuple[int, int], out=None,
                        handle_sink=None):
    """Page record + the device tensors it points to (kept alive by the caller).
    `handle_sink`: a list; when given the handle arrays are appended as (record, src, dst) and
    NOT uploaded -- the batch engine uploads all pages' handles in one copy and patches the
    pointers (`upload_handles`)."""
    height, width = shape
    rec = new_grid_page(height, width, config.grid_size, out)
    rec['projector'] = nv.PROJ_MLS
    rec['resize_as_src'] = int(config.resize_as_src)
    # PointTuple.to_smooth_np_array: rounded coordinates as float32 (mls.py:49-50)
    src = np.ascontiguousarray(PointTuple(config.src_handle_points).to_smooth_np_array())
    dst = np.ascontiguousarray(PointTuple(config.dst_handle_points).to_smooth_np_array())
    assert src.shape == dst.shape and src.ndim == 2
    rec['n_handles'] = src.shape[0]
    if handle_sink is not None:
        handle_sink.append((rec, src, dst))
        return rec, None
    handles = dv.to_device(np.stack([src, dst]))
    rec['handles_src'] = handles[0].data_ptr()
    rec['handles_dst'] = handles[1].data_ptr()
    return rec, handles


def upload_handles(handle_sink):
    """One H2D copy for the handles of every page in `handle_sink`; returns the device tensor the
    page records now point into."""
    if not handle_sink:
        return None
    flat = np.concatenate([np.concatenate([src.reshape(-1), dst.reshape(-1)])
                           for _, src, dst in handle_sink]).astype(np.float32)
    dev = dv.to_device(flat)
    base = dev.data_ptr()
    offset = 0
    for rec, src, dst in handle_sink:
        rec['handles_src'] = base + offset * 4
        rec['handles_dst'] = base + (offset + src.size) * 4
        offset += src.size + dst.size
    return dev


class SimilarityMlsState(DistortionStateImageGridBased):

    def __init__(self, config: SimilarityMlsConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        rec, handles = similarity_mls_page(config, shape)
        self.initialize_grid_plan(rec, keepalive=[handles])
        # for debug only, like the reference (mls.py:156-157)
        self.dst_handle_points = list(map(self.shift_and_resize_point, config.dst_handle_points))


similarity_mls = DistortionImageGridBased(config_cls=SimilarityMlsConfig,
                                          state_cls=SimilarityMlsState)
