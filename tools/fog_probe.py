"""fog_image per call (host time and host + device time) with the field computed on the device
(0.25 ms per 1024 x 1024 page on the B200; the host-drawn field took 17 ms) + a cProfile of the call.

    python tools/fog_probe.py
"""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from vkit_b200 import element
from vkit_b200.mechanism import distortion
from vkit_b200.mechanism.distortion.photometric import effect
img = element.Image(mat=torch.randint(0, 256, (1024, 1024, 3), dtype=torch.uint8, device='cuda'))
cfg = effect.FogConfig(roughness=0.5)
for gen_name in ('PCG64', 'MT19937'):
    rng = np.random.default_rng(3) if gen_name == 'PCG64' else np.random.Generator(np.random.MT19937(3))
    for _ in range(3):
        distortion.fog.distort(cfg, image=img, rng=rng)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        distortion.fog.distort(cfg, image=img, rng=rng)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(gen_name, 'host ms per call %.3f, incl. device %.3f' % (1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
import cProfile, pstats
rng = np.random.default_rng(3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    distortion.fog.distort(cfg, image=img, rng=rng)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
