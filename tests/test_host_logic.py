"""Host-side logic that needs no GPU: the C ABI exports, config structuring, policy sampling,
containers on host arrays, batch bookkeeping and the 2-rank sharding helpers (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from vkit_b200 import _native
    from vkit_b200 import build as vk_build
    if vk_build.needs_build():  # stale or missing in-tree library: nvcc cross-compiles without a GPU
        vk_build.build()
    header = open(os.path.join(ROOT, 'include', 'vkit_b200.h')).read()
    declared = set(re.findall(r'^(?:int|const char\*)\s+(vkb_\w+)\s*\(', header, flags=re.M))
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in include/vkit_b200.h but not exported'
    assert set(_native.EXPORTS) <= declared
    assert _native.lib().vkb_version() >= 1


def test_struct_layouts_match_numpy_dtypes():
    from vkit_b200 import _native as nv
    for struct, dtype in ((nv.Planes, nv.PLANES_DTYPE), (nv.WarpPage, nv.WARP_PAGE_DTYPE),
                          (nv.GridPage, nv.GRID_PAGE_DTYPE), (nv.GridMeta, nv.GRID_META_DTYPE),
                          (nv.BlendItem, nv.BLEND_ITEM_DTYPE), (nv.PhotoPage, nv.PHOTO_PAGE_DTYPE),
                          (nv.PolyItem, nv.POLY_ITEM_DTYPE), (nv.ColorOp, nv.COLOR_OP_DTYPE)):
        assert ctypes.sizeof(struct) == dtype.itemsize
        for name, _ in struct._fields_:
            assert getattr(struct, name).offset == dtype.fields[name][1]


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from vkit_b200 import _native
    from vkit_b200.element import Image
    from vkit_b200.mechanism import distortion
    image = Image(mat=np.zeros((32, 32, 3), np.uint8))
    with pytest.raises(_native.NativeError):
        distortion.rotate.distort({'angle': 30}, image=image)
    with pytest.raises(_native.NativeError):
        distortion.mean_shift.distort({'delta': 3}, image=image)
    with pytest.raises(_native.NativeError):
        distortion.camera_plane_only.distort(
            {'camera_model_config': {'rotation_unit_vec': [1.0, 0.0, 0.0], 'rotation_theta': 10},
             'grid_size': 15}, image=image)


def test_dyn_structure_and_names():
    from vkit_b200.mechanism import distortion
    from vkit_b200.utility import dyn_structure
    cfg = dyn_structure({'camera_model_config': {'rotation_unit_vec': [1, 0, 0],
                                                 'rotation_theta': 5}, 'grid_size': 20},
                        distortion.CameraPlaneOnlyConfig)
    assert isinstance(cfg.camera_model_config, distortion.CameraModelConfig)
    assert cfg.camera_model_config.rotation_theta == 5.0
    with pytest.raises(TypeError):
        dyn_structure({'angle': 3, 'bogus': 1}, distortion.RotateConfig)
    assert dyn_structure({'delta': 3, 'oob_behavior': 'cycle'},
                         distortion.MeanShiftConfig).oob_behavior == distortion.OutOfBoundBehavior.CYCLE
    assert distortion.GaussionNoiseConfig.get_name() == 'gaussion_noise'
    assert distortion.CameraPlaneLineFoldConfig.get_name() == 'camera_plane_line_fold'
    same = distortion.RotateConfig(angle=3)
    assert dyn_structure(same, distortion.RotateConfig) is same


def test_affine_states_match_oracle():
    from oracle import vkit_port as port
    from vkit_b200.mechanism import distortion
    port.use_cv2(False)
    for name, cfg in [('rotate', {'angle': a}) for a in (0, 1, 30, 90, 91, 180, 200, 270, 359, 360,
                                                         -45)] + \
            [('shear_hori', {'angle': a}) for a in (-30, 0, 12)] + \
            [('shear_vert', {'angle': a}) for a in (-7, 0, 29)]:
        for shape in ((64, 97), (512, 512), (1024, 777)):
            op = getattr(distortion, name)
            state = op.state_cls(op.config_cls(**cfg), shape, None)
            trans_mat, dsize = port.affine_state(name, cfg, shape)
            if trans_mat is None:
                assert state.trans_mat is None or cfg['angle'] % 360 == 0
                continue
            assert np.array_equal(state.trans_mat, trans_mat) and tuple(state.dsize) == tuple(dsize)
    for name in ('skew_hori', 'skew_vert'):
        for ratio in (-0.3, 0.12, 0.35):
            op = getattr(distortion, name)
            state = op.state_cls(op.config_cls(ratio=ratio), (300, 431), None)
            trans_mat, dsize = port.affine_state(name, {'ratio': ratio}, (300, 431))
            np.testing.assert_allclose(state.trans_mat, trans_mat, rtol=0, atol=1e-9)


def test_policy_configs_are_rng_identical_to_golden_reference_configs():
    """The golden cases store configs produced by the REFERENCE generators from a seed; the
    product generators must return the same numbers and leave the generator in the same state."""
    from common import golden_cases
    from vkit_b200.mechanism.distortion_policy import random_distortion as rd
    factories = {}
    for group, _ in (rd._PHOTOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS
                     + rd._GEOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS):
        for factory in group:
            factories[factory.name] = factory
    assert len(factories) == 35
    grid_ops = ['camera_plane_only', 'camera_cubic_curve', 'camera_plane_line_fold',
                'camera_plane_line_curve', 'similarity_mls']
    checked = 0
    for i, name in enumerate(grid_ops):
        for level, seed, shape in ((3, 21 + i, (136, 176)), (9, 41 + i, (136, 176)),
                                   (7, 61 + i, (1024, 1024))):
            policy = factories[name].create()
            config = policy.config_generator_cls(policy.config_for_config_generator, level)(
                shape, np.random.default_rng(seed))
            case = [c for c in golden_cases('geometric', op=name)
                    if c['seed'] == seed and tuple(c['shape']) == shape][0]
            ref = case['config']
            assert config.grid_size == ref['grid_size']
            if name == 'similarity_mls':
                mine = [[p.smooth_x, p.smooth_y] for p in config.dst_handle_points]
                assert mine == ref['dst_handle_points']
            else:
                cam = config.camera_model_config
                assert list(map(float, cam.rotation_unit_vec)) == ref['camera_model_config'][
                    'rotation_unit_vec']
                assert cam.rotation_theta == ref['camera_model_config']['rotation_theta']
            checked += 1
    assert checked == 15


def test_random_distortion_sampling_order():
    """Stage gate, policy choice and level draws consume the generator in the documented order
    (random_distortion.py:151-188): replaying with recording stubs gives a stable sequence."""
    from vkit_b200.mechanism.distortion_policy import random_distortion as rd

    calls = []

    class StubPolicy:
        def __init__(self, name):
            self.name = name

        def distort(self, level, rng=None, **kwargs):
            calls.append((self.name, int(level)))
            rng.random()
            from vkit_b200.mechanism.distortion.interface import DistortionResult
            return DistortionResult(shape=kwargs['shapable_or_shape'])

    stage_cfg = rd.RandomDistortionStageConfig(
        distortion_policies=[StubPolicy(n) for n in ('a_blur', 'b_noise', 'c_blur', 'd', 'e')],
        distortion_policy_weights=[1, 2, 3, 4, 5], prob_enable=0.9, num_distortions_min=0,
        num_distortions_max=3, conflict_control_keyword_groups=[['blur'], ['noise']])
    chain = rd.RandomDistortion([stage_cfg], 1, 10)
    first = []
    for seed in range(30):
        calls.clear()
        chain.distort(np.random.default_rng(seed), shapable_or_shape=(50, 60))
        names = [c[0] for c in calls]
        assert len([n for n in names if 'blur' in n]) <= 1  # conflict control
        assert len(set(names)) == len(names)  # without replacement
        first.append(tuple(calls))
    calls.clear()
    chain.distort(np.random.default_rng(7), shapable_or_shape=(50, 60))
    assert tuple(calls) == first[7]  # deterministic given the seed
    assert any(len(c) == 0 for c in first) and any(len(c) >= 2 for c in first)


def test_host_backed_containers():
    from vkit_b200.element import (Box, Image, ImageMode, Mask, Point, PointList, PointTuple,
                                   Polygon, ScoreMap)
    p = Point.create(y=2.5, x=3.5)
    assert (p.y, p.x) == (2, 4)  # Python round: half to even
    assert Point.create(y=2.5, x=0).to_shifted_point(offset_y=-1).y == 2  # round(1.5), not 2 - 1
    pts = PointList([Point.create(y=1.2, x=7.7), Point.create(y=5, x=6)])
    assert pts.to_np_array().tolist() == [[8, 1], [6, 5]]
    assert pts.to_smooth_np_array().dtype == np.float32
    assert PointTuple(pts).to_smooth_np_array().tolist() == [[8.0, 1.0], [6.0, 5.0]]  # rounded
    assert PointList.from_np_array(np.array([[0, 0], [4, 0], [4, 4], [0, 0]])).to_xy_pairs() == [
        (0, 0), (4, 0), (4, 4)]
    box = Box(up=2, down=5, left=1, right=8)
    assert box.shape == (4, 8) and box.to_shifted_box(offset_y=1).up == 3
    poly = box.to_polygon()
    assert poly.bounding_box == box and poly.num_points == 4
    image = Image(mat=np.zeros((6, 9, 3), np.uint8))
    assert image.mode == ImageMode.RGB and image.shape == (6, 9) and not image.on_device
    assert not image.mat.flags.writeable
    with image.writable_context:
        image.mat[0, 0] = 7
    assert image.mat[0, 0, 0] == 7 and not image.mat.flags.writeable
    assert Image(mat=np.zeros((4, 4), np.uint8)).mode == ImageMode.GRAYSCALE
    with pytest.raises(NotImplementedError):
        Image(mat=np.zeros((4, 4, 2), np.uint8))
    with pytest.raises(RuntimeError):
        Mask(mat=np.zeros((4, 4), np.float32))
    with pytest.raises(RuntimeError):
        ScoreMap(mat=np.full((4, 4), 1.5, np.float32))
    ScoreMap(mat=np.full((4, 4), 1.5, np.float32), is_prob=False)
    mask = Mask(mat=np.eye(4, dtype=np.uint8))
    assert mask.to_inverted_mask().mat.sum() == 12 and mask.np_mask.sum() == 4
    crop = image.to_cropped_image(up=1, down=3, left=2, right=5)
    assert crop.shape == (3, 4)
    assert mask.to_external_box() == Box(up=0, down=3, left=0, right=3)


def test_batch_bookkeeping_and_seed_lists():
    import bench
    rngs_a = bench.page_rngs(4, 3, 16)
    rngs_b = bench.page_rngs(0, 16, 16)[4:7]
    assert [r.integers(0, 1 << 30) for r in rngs_a] == [r.integers(0, 1 << 30) for r in rngs_b]
    names, configs = bench.sample_page_configs(0, 8, 8)
    assert names[:4] == list(bench.CAMERA_OPS) and len(configs) == 8
    names2, configs2 = bench.sample_page_configs(4, 4, 8)
    assert names2 == names[4:] and configs2[0] == configs[4]  # a rank's slice == the global list


def test_two_rank_sharding_with_gloo(tmp_path):
    """World size 2 on CPU: the seed broadcast + counter all_reduce used by bench.py."""
    script = tmp_path / 'shard.py'
    script.write_text('''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import bench
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
batch = 4
seeds = torch.arange(batch * world, dtype=torch.int64) + bench.BASE_SEED if rank == 0 else torch.zeros(batch * world, dtype=torch.int64)
dist.broadcast(seeds, src=0)
mine = seeds[rank * batch:(rank + 1) * batch]
names, configs = bench.sample_page_configs(rank * batch, batch, batch * world)
counters = torch.tensor([batch, int(mine.sum())], dtype=torch.int64)
dist.all_reduce(counters, op=dist.ReduceOp.SUM)
stats = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(stats, op=dist.ReduceOp.MAX)
assert int(counters[0]) == batch * world
assert int(counters[1]) == int(seeds.sum())
assert float(stats[0]) == float(world)
assert names[0] == bench.CAMERA_OPS[(rank * batch) %% 4]
open(os.path.join(%r, 'rank%%d.ok' %% rank), 'w').write('ok')
''' % (ROOT, str(tmp_path)))
    import socket
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    out = None
    for _ in range(3):  # a freshly released port can still be refused; pick another one
        with socket.socket() as sock:
            sock.bind(('127.0.0.1', 0))
            port = sock.getsockname()[1]
        out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                              '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port',
                              str(port), str(script)], capture_output=True, text=True, env=env,
                             timeout=300)
        if out.returncode == 0:
            break
    assert out.returncode == 0, out.stdout + out.stderr
    # one marker file per rank (the ranks' stdout interleaves)
    assert (tmp_path / 'rank0.ok').exists() and (tmp_path / 'rank1.ok').exists(), out.stdout


def test_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['e2e']['h2d_bytes_per_step'] == 0


def test_vectorised_gaussian_taps_equal_scalar():
    """PhotometricBatch builds the 8.8 fixed-point taps of all pages at once: same integers as the
    per-page form for every kernel size."""
    from vkit_b200.mechanism.distortion.photometric import blur
    rng = np.random.default_rng(3)
    sigmas = np.concatenate([rng.uniform(0.3, 5.6, 4000), [0.5, 1.0, 1.5, 2.5, 5.0]])
    ksizes, taps = blur.gaussian_kernels_u8(sigmas)
    for i, sigma in enumerate(sigmas):
        ksize = blur._estimate_gaussian_kernel_size(float(sigma))
        assert ksizes[i] == ksize
        if ksize <= 17:
            assert list(taps[i, :ksize]) == blur.gaussian_kernel_u8(ksize, float(sigma))
            assert not taps[i, ksize:].any()


def test_glass_swap_maps_equal_plain_restatement():
    """The strided-view form of the glass_blur permutation against a literal restatement of the
    reference's loop (vkit/mechanism/distortion/photometric/blur.py:232-262): same maps, same
    generator state afterwards."""
    from vkit_b200.mechanism.distortion.photometric.blur import glass_swap_maps

    def plain(shape, delta, loop, rng):
        height, width = shape
        pos_x, pos_y = np.meshgrid(np.arange(width), np.arange(height))
        period = 2 * delta + 1
        for _ in range(loop):
            rows = np.arange(rng.integers(0, period), height - delta, period).reshape(-1, 1)
            cols = np.arange(rng.integers(0, period), width - delta, period).reshape(1, -1)
            grid_shape = (rows.shape[0], cols.shape[1])
            shift_y = rng.integers(-delta, delta + 1, grid_shape)
            shift_x = rng.integers(-delta, delta + 1, grid_shape)
            target_y = np.clip(pos_y[rows, cols] + shift_y, 0, height - 1)
            target_x = np.clip(pos_x[rows, cols] + shift_x, 0, width - 1)
            for pos in (pos_y, pos_x):
                at_centre, at_target = pos[rows, cols], pos[target_y, target_x]
                pos[rows, cols] = at_target
                pos[target_y, target_x] = at_centre
        return pos_y, pos_x

    for shape, delta, loop in (((257, 300), 1, 5), ((64, 64), 3, 4), ((5, 7), 1, 2),
                               ((33, 500), 2, 6), ((1, 9), 1, 3)):
        a_rng, b_rng = np.random.default_rng(5), np.random.default_rng(5)
        ref = plain(shape, delta, loop, a_rng)
        got = glass_swap_maps(shape, delta, loop, b_rng)
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]), shape
        assert a_rng.integers(0, 1 << 30) == b_rng.integers(0, 1 << 30)


def test_diamond_square_mask_equals_plain_restatement():
    """The slice-based fog field against the literal np.roll restatement of the reference
    (vkit/mechanism/distortion/photometric/effect.py:89-146): same float32 field, same generator
    state afterwards."""
    from vkit_b200.mechanism.distortion.photometric.effect import generate_diamond_square_mask

    def _midpoint_level(sums, weight, rng):
        """One displacement level: mean of the four neighbours damped by (1 - weight) plus
        weight * U(0, 1).  The dtype sequence is the reference's (float32 sums, float64 draws)."""
        return (1 - weight) * sums / 4 + weight * rng.uniform(0, 1, sums.shape)


    def plain_mask(shape, roughness, rng):
        """Diamond-square plasma field cropped to `shape` (effect.py:89-146); consumes `rng` exactly
        like the reference: 4 corner draws, then per level the diamond draw, the two square draws,
        and finally the crop offsets."""
        assert 0.0 <= roughness <= 1.0
        height, width = shape
        size = int(2**np.ceil(np.log2(max(height, width))) + 1)
        field = np.zeros((size, size), dtype=np.float32)
        for corner in ((0, 0), (0, -1), (-1, -1), (-1, 0)):
            field[corner] = rng.uniform(0.0, 1.0)

        step, level = size - 1, 0
        while step >= 2:
            weight = roughness**level
            half = step // 2
            corners = field[0:size:step, 0:size:step]
            down_pairs = corners + np.roll(corners, shift=-1, axis=0)
            right_pairs = corners + np.roll(corners, shift=-1, axis=1)

            # centres of the squares
            centres = _midpoint_level((down_pairs + right_pairs)[:-1, :-1], weight, rng)
            field[half:size:step, half:size:step] = centres

            # edge midpoints on the corner rows: left/right corners + centres above/below (wrapping)
            above_below = centres + np.roll(centres, shift=1, axis=0)
            above_below = np.vstack([above_below, above_below[0]])
            field[0:size:step, half:size:step] = _midpoint_level(right_pairs[:, :-1] + above_below,
                                                                 weight, rng)

            # edge midpoints on the corner columns
            left_right = centres + np.roll(centres, shift=1, axis=1)
            left_right = np.hstack([left_right, left_right[0].reshape(-1, 1)])
            field[half:size:step, 0:size:step] = _midpoint_level(down_pairs[:-1] + left_right, weight,
                                                                 rng)
            level += 1
            step = half

        up = rng.integers(0, size - height + 1)
        left = rng.integers(0, size - width + 1)
        return field[up:up + height, left:left + width]

    for shape, roughness in (((257, 300), 0.6), ((64, 96), 0.9), ((100, 133), 0.3), ((5, 3), 0.7),
                             ((2, 2), 0.5), ((1, 1), 0.5), ((513, 40), 0.0), ((40, 513), 1.0)):
        a_rng, b_rng = np.random.default_rng(11), np.random.default_rng(11)
        ref = plain_mask(shape, roughness, a_rng)
        got = generate_diamond_square_mask(shape, roughness, b_rng)
        assert ref.dtype == got.dtype and np.array_equal(ref, got), shape
        assert a_rng.integers(0, 1 << 30) == b_rng.integers(0, 1 << 30)
