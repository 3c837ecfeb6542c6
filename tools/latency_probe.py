"""Single-page latency of the drop-in API (what vkit.pipeline sees: one page per call).
Host NumPy containers in / `.mat` read back out, and device-resident containers, per op."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
import bench  # noqa: E402
from vkit_b200.element import Image, Mask, ScoreMap  # noqa: E402
from vkit_b200.mechanism import distortion  # noqa: E402
from vkit_b200.mechanism.distortion_policy import random_distortion as rd  # noqa: E402

shape = (1024, 1024)
rng = np.random.default_rng(0)
image = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
mask = (rng.random(shape) > 0.5).astype(np.uint8)
score = rng.random(shape).astype(np.float32)


def policy_config(name, level=6, seed=1):
    for group in (rd._PHOTOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS
                  + rd._GEOMETRIC_POLICY_FACTORIES_AND_DEFAULT_WEIGHTS_SUM_PAIRS):
        for fac in group[0]:
            if fac.name == name:
                pol = fac.create()
                return pol.config_generator_cls(pol.config_for_config_generator, level)(
                    shape, np.random.default_rng(seed))
    raise KeyError(name)


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


rows = []
for name, labels in (('rotate', True), ('camera_cubic_curve', True), ('similarity_mls', True),
                     ('gaussian_blur', False), ('color_shift', False), ('mean_shift', False)):
    cfg = policy_config(name)
    op = getattr(distortion, name)

    def host_call():
        kw = dict(image=Image(mat=image))
        if labels:
            kw.update(mask=Mask(mat=mask), score_map=ScoreMap(mat=score))
        r = op.distort(cfg, **kw)
        _ = r.image.mat
        if labels:
            _ = r.mask.mat
            _ = r.score_map.mat

    dimg, dmask, dscore = (torch.from_numpy(a).cuda() for a in (image, mask, score))

    def dev_call():
        kw = dict(image=Image(mat=dimg))
        if labels:
            kw.update(mask=Mask(mat=dmask), score_map=ScoreMap(mat=dscore, is_prob=False))
        op.distort(cfg, **kw)

    rows.append((name, 'image+mask+score_map' if labels else 'image', timeit(host_call), timeit(dev_call)))
print(f'{"op":22s} {"containers":22s} {"host in/out ms":>15s} {"device resident ms":>19s}')
for name, what, a, b in rows:
    print(f'{name:22s} {what:22s} {a:15.2f} {b:19.2f}')
