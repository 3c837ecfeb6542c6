"""Tie flips of a grid op against the NumPy oracle as a function of the page length (the float32 map
values get coarser with the coordinate: 7 / 4 / 1 / 0 differing pixels at 16 100 / 15 900 / 8 000 /
4 000 px on the B200; the first page is beyond the fast path's fixed point and runs on the float64
path only).

    python tools/longpage_probe.py
"""
import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import attrs
from oracle import vkit_port as port
from vkit_b200 import element
from vkit_b200.mechanism import distortion
from vkit_b200.mechanism.distortion_policy.geometric import camera as cam_policy
from test_gpu_parity import make_inputs
def plain(obj):
    if attrs.has(type(obj)): return {f.name: plain(getattr(obj, f.name)) for f in attrs.fields(type(obj))}
    if isinstance(obj, (list, tuple)): return [plain(v) for v in obj]
    return obj
for shape in ((48, 16100), (48, 15900), (48, 8000), (48, 4000)):
    policy = cam_policy.camera_plane_only_policy_factory.create()
    gen = policy.config_generator_cls(policy.config_for_config_generator, 6)
    config = attrs.evolve(gen(shape, np.random.default_rng(424242)), grid_size=45)
    image, mask, score_map = make_inputs(424242 % 100000, shape)
    r = distortion.camera_plane_only.distort(config, image=element.Image(mat=image), mask=element.Mask(mat=mask))
    port.use_cv2(False)
    ref = port.grid_distort('camera_plane_only', plain(config), shape, image=image, mask=mask)
    d = (r.image.mat != ref['image']).any(axis=-1)
    print(shape, r.shape, 'differing px', int(d.sum()), 'of', d.size)
