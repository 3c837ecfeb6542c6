import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np, torch
from vkit_b200.mechanism.distortion_policy import random_distortion_factory
from vkit_b200.mechanism.distortion_policy import random_distortion_batch as rdb
from vkit_b200.element import Image
n, shape = 128, (1024, 1024)
rd = random_distortion_factory.create({'disabled_policy_names': ['defocus_blur', 'zoom_in_blur'], 'force_post_rotate': True})
batch = rdb.RandomDistortionBatch(rd)
images = torch.randint(0, 256, (n,) + shape + (3,), dtype=torch.uint8, device='cuda')
masks = (torch.rand((n,) + shape, device='cuda') > 0.5).to(torch.uint8)
gen = np.random.default_rng(7)
points = [gen.uniform(0, 1023, (256, 2)) for _ in range(n)]
polygons = []
for _ in range(n):
    polys = []
    for _ in range(64):
        x0, y0 = gen.uniform(0, 900, 2); w, h = gen.uniform(10, 120, 2)
        polys.append(np.asarray([(x0, y0), (x0 + w, y0), (x0 + w, y0 + h), (x0, y0 + h)]))
    polygons.append(polys)
seqs = np.random.SeedSequence(133700).spawn(n * 4)
# per-op timing of the photometric stage
per_op = {}
for k in range(1, 4):
    rngs = [np.random.default_rng(s) for s in seqs[k * n:(k + 1) * n]]
    t0 = time.perf_counter()
    chains = [batch.sample_chain(r, shape) for r in rngs]
    t_chain = time.perf_counter() - t0
    torch.cuda.synchronize()
    for i, ch in enumerate(chains):
        if not ch.photometric: continue
        image = Image(mat=images[i])
        for policy, _, config in ch.photometric:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            image = policy.distortion.distort_image(config, image)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            e = per_op.setdefault(policy.name, [0, 0.0, 0.0]); e[0] += 1; e[1] += t1 - t0; e[2] += t2 - t0
    print(f'step {k}: chains {t_chain*1e3/n:.3f} ms/page')
tot_h = sum(v[1] for v in per_op.values()); tot = sum(v[2] for v in per_op.values())
for name, (c, h, t) in sorted(per_op.items(), key=lambda kv: -kv[1][2]):
    print(f'{name:26s} n={c:3d} host {h/c*1e3:7.3f} ms  host+gpu {t/c*1e3:7.3f} ms  share {t/tot:5.1%}')
print(f'photometric total per page: host {tot_h/(3*n)*1e3:.3f} ms, host+gpu {tot/(3*n)*1e3:.3f} ms')
# whole distort, and geometric only (no photometric): patch chains
for label in ('full',):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rngs = [np.random.default_rng(s) for s in seqs[0:n]]
    out = batch.distort(rngs, images, masks, points, polygons)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'{label}: host {(t1-t0)/n*1e3:.3f} ms/page, wall {(t2-t0)/n*1e3:.3f} ms/page')
