"""camera_plane_only / camera_cubic_curve / camera_plane_line_fold / camera_plane_line_curve
(vkit/mechanism/distortion/geometric/camera.py).

Host side: the per-page camera constants (Rodrigues, translation, focal length, curve / line
parameters) with the reference's dtype sequence.  Device side: lattice lifting + projection
(`vkb_grid_project`) and everything downstream.
"""
import math
from typing import Optional, Sequence, Tuple

import attrs
import numpy as np
from numpy.random import Generator as RandomGenerator

from vkit_b200 import _native as nv

from ..interface import DistortionConfig
from ._gridcore import new_grid_page
from ._hostmath import rodrigues
from .grid_rendering import DistortionImageGridBased, DistortionStateImageGridBased


@attrs.define
class CameraModelConfig:
    rotation_unit_vec: Sequence[float]
    rotation_theta: float
    focal_length: Optional[float] = None
    principal_point: Optional[Sequence[float]] = None
    camera_distance: Optional[float] = None


def complete_camera_model_config(height: int, width: int, config: CameraModelConfig):
    # camera.py:220-241 (principal point default is [height // 2, width // 2], consumed as x, y)
    if config.principal_point and config.focal_length and config.camera_distance:
        return config
    principal_point = config.principal_point or [height // 2, width // 2]
    focal_length, camera_distance = config.focal_length, config.camera_distance
    if not focal_length or not camera_distance:
        focal_length = max(height, width)
        camera_distance = focal_length
    return CameraModelConfig(rotation_unit_vec=config.rotation_unit_vec,
                             rotation_theta=config.rotation_theta, focal_length=focal_length,
                             principal_point=principal_point, camera_distance=camera_distance)


_F32 = np.float32


def fill_camera_model(rec: np.ndarray, config: CameraModelConfig):
    """CameraModel.__init__ (camera.py:157-175) -> R (double), t, focal in the page record.

    Same NumPy float32 operations as the reference line by line (the lattice these constants
    feed is rounded to integers, so the last bit matters); only the call overhead is trimmed:
    `norm` is its own definition sqrt(x.dot(x)), and R^T . (0, 0, d) is the third row of R times
    d (the other products are exact zeros), which leaves one real float32 matmul."""
    assert config.focal_length and config.camera_distance and config.principal_point
    unit_vec = np.array(config.rotation_unit_vec, dtype=_F32)
    length = np.sqrt(unit_vec.dot(unit_vec))  # np.linalg.norm of a 1-D float32 vector
    if length != 1.0:
        unit_vec /= length
    theta = float(min(max(config.rotation_theta, -89), 89) / 180 * np.pi)
    rotation_vec = unit_vec * theta  # float32

    pp = config.principal_point
    principal_point = np.array((pp[0], pp[1], pp[2] if len(pp) > 2 else 0), dtype=_F32)

    rotation_mat64 = rodrigues(rotation_vec)  # double, as cv.projectPoints recomputes it below
    rotation_mat = rotation_mat64.astype(_F32)  # cv.Rodrigues(float32) -> float32
    wc_shifted_original_vec = rotation_mat[2] * _F32(config.camera_distance)
    wc_shifted_principal_point_vec = wc_shifted_original_vec - principal_point
    translation_vec = np.matmul(rotation_mat, wc_shifted_principal_point_vec.reshape(3, 1))

    # cv.projectPoints converts rvec / tvec / K to double and recomputes Rodrigues in double
    rec['R'] = rotation_mat64.reshape(-1)
    rec['t'] = translation_vec.reshape(-1)
    rec['focal'] = float(_F32(config.focal_length))
    rec['projector'] = nv.PROJ_CAMERA


class DistortionStateCameraOperation(DistortionStateImageGridBased):

    complete_camera_model_config = staticmethod(complete_camera_model_config)

    def initialize_camera_operation(self, height: int, width: int, grid_size: int,
                                    camera_model_config: CameraModelConfig, rec: np.ndarray):
        camera_model_config = complete_camera_model_config(height, width, camera_model_config)
        fill_camera_model(rec, camera_model_config)
        self.initialize_grid_plan(rec)


# ---- plane only --------------------------------------------------------------------------
@attrs.define
class CameraPlaneOnlyConfig(DistortionConfig):
    camera_model_config: CameraModelConfig
    grid_size: int


def plane_only_page(config: CameraPlaneOnlyConfig, shape: Tuple[int, int], out=None):
    height, width = shape
    rec = new_grid_page(height, width, config.grid_size, out)
    rec['strategy'] = nv.CAM_PLANE
    return rec


class CameraPlaneOnlyState(DistortionStateCameraOperation):

    def __init__(self, config: CameraPlaneOnlyConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        self.initialize_camera_operation(height, width, config.grid_size,
                                         config.camera_model_config,
                                         plane_only_page(config, shape))


camera_plane_only = DistortionImageGridBased(config_cls=CameraPlaneOnlyConfig,
                                             state_cls=CameraPlaneOnlyState)


# ---- cubic curve -------------------------------------------------------------------------
@attrs.define
class CameraCubicCurveConfig(DistortionConfig):
    curve_alpha: float
    curve_beta: float
    # clockwise, [0, 180]
    curve_direction: float
    curve_scale: float
    camera_model_config: CameraModelConfig
    grid_size: int


def cubic_curve_page(config: CameraCubicCurveConfig, shape: Tuple[int, int], out=None):
    # CameraCubicCurvePoint2dTo3dStrategy.__init__ (camera.py:325-373)
    height, width = shape
    rec = new_grid_page(height, width, config.grid_size, out)
    rec['strategy'] = nv.CAM_CUBIC
    alpha = math.tan(min(max(config.curve_alpha, -80), 80) / 180 * np.pi)
    beta = math.tan(min(max(config.curve_beta, -80), 80) / 180 * np.pi)
    direction = (config.curve_direction % 180) / 180 * np.pi
    rotation_mat = np.asarray(
        [[math.cos(direction), math.sin(direction)], [-math.sin(direction), math.cos(direction)]],
        dtype=np.float32)
    corners = np.asarray([[0, 0], [width - 1, 0], [width - 1, height - 1], [0, height - 1]],
                         dtype=np.float32)
    rotated_corners = np.matmul(rotation_mat, corners.transpose())
    proj_min = rotated_corners[0].min()
    proj_range = rotated_corners[0].max() - proj_min
    rec['rot2'] = rotation_mat.reshape(-1)
    rec['proj_min'] = proj_min
    rec['proj_range'] = proj_range
    rec['poly'] = [alpha + beta, -2 * alpha - beta, alpha, 0]
    rec['curve_scale'] = config.curve_scale
    return rec


class CameraCubicCurveState(DistortionStateCameraOperation):

    def __init__(self, config: CameraCubicCurveConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        self.initialize_camera_operation(height, width, config.grid_size,
                                         config.camera_model_config,
                                         cubic_curve_page(config, shape))


camera_cubic_curve = DistortionImageGridBased(config_cls=CameraCubicCurveConfig,
                                              state_cls=CameraCubicCurveState)


# ---- plane line fold / curve -------------------------------------------------------------
def plane_line_page(shape: Tuple[int, int], grid_size: int, point, direction: float, perturb_vec,
                    alpha: float, strategy: int, out=None):
    # CameraPlaneLinePoint2dTo3dStrategy.__init__ (camera.py:434-462)
    height, width = shape
    rec = new_grid_page(height, width, grid_size, out)
    rec['strategy'] = strategy
    point = np.asarray(point, dtype=np.float32)
    direction = (direction % 180) / 180 * np.pi
    cos_theta = np.cos(direction)
    sin_theta = np.sin(direction)
    rec['line_ab'] = np.asarray([sin_theta, -cos_theta], dtype=np.float32)
    rec['line_c'] = -point[0] * sin_theta + point[1] * cos_theta
    rec['dist_max'] = np.sqrt(height**2 + width**2)
    rec['line_alpha'] = alpha
    rec['perturb'] = np.asarray(perturb_vec, dtype=np.float32)
    return rec


@attrs.define
class CameraPlaneLineFoldConfig(DistortionConfig):
    fold_point: Tuple[float, float]
    # clockwise, [0, 180]
    fold_direction: float
    fold_perturb_vec: Tuple[float, float, float]
    fold_alpha: float
    camera_model_config: CameraModelConfig
    grid_size: int


def plane_line_fold_page(config: CameraPlaneLineFoldConfig, shape: Tuple[int, int], out=None):
    return plane_line_page(shape, config.grid_size, config.fold_point, config.fold_direction,
                           config.fold_perturb_vec, config.fold_alpha, nv.CAM_LINE_FOLD, out)


class CameraPlaneLineFoldState(DistortionStateCameraOperation):

    def __init__(self, config: CameraPlaneLineFoldConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        self.initialize_camera_operation(height, width, config.grid_size,
                                         config.camera_model_config,
                                         plane_line_fold_page(config, shape))


camera_plane_line_fold = DistortionImageGridBased(config_cls=CameraPlaneLineFoldConfig,
                                                  state_cls=CameraPlaneLineFoldState)


@attrs.define
class CameraPlaneLineCurveConfig(DistortionConfig):
    curve_point: Tuple[float, float]
    # clockwise, [0, 180]
    curve_direction: float
    curve_perturb_vec: Tuple[float, float, float]
    curve_alpha: float
    camera_model_config: CameraModelConfig
    grid_size: int


def plane_line_curve_page(config: CameraPlaneLineCurveConfig, shape: Tuple[int, int], out=None):
    return plane_line_page(shape, config.grid_size, config.curve_point, config.curve_direction,
                           config.curve_perturb_vec, config.curve_alpha, nv.CAM_LINE_CURVE, out)


class CameraPlaneLineCurveState(DistortionStateCameraOperation):

    def __init__(self, config: CameraPlaneLineCurveConfig, shape: Tuple[int, int],
                 rng: Optional[RandomGenerator]):
        height, width = shape
        self.initialize_camera_operation(height, width, config.grid_size,
                                         config.camera_model_config,
                                         plane_line_curve_page(config, shape))


camera_plane_line_curve = DistortionImageGridBased(config_cls=CameraPlaneLineCurveConfig,
                                                   state_cls=CameraPlaneLineCurveState)
