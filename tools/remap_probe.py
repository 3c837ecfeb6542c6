"""Times the fused remap launch alone (inputs and parameter blocks resident) for debug variants
of the kernel: only the vkb_grid_remap call sits between the CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from vkit_b200 import _native as nv
from vkit_b200 import device as dv
from vkit_b200.batch import GeometricBatch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
names, configs = bench.sample_page_configs(0, n, 256)
pages = torch.randint(0, 256, (n, 1024, 1024, 3), dtype=torch.uint8, device='cuda')
eng = GeometricBatch(names, configs, (1024, 1024))
plan = eng.plan_batch()
shapes = [plan.result_shape(i) for i in range(n)]
offsets = np.concatenate([[0], np.cumsum([h * w for h, w in shapes])])
arena = dv.empty((int(offsets[-1]) * 3,), np.uint8)
planes = np.zeros(n, dtype=nv.PLANES_DTYPE)
planes['src_h'], planes['src_w'] = 1024, 1024
planes['dst_h'] = [s[0] for s in shapes]
planes['dst_w'] = [s[1] for s in shapes]
planes['src_image'] = pages.data_ptr() + np.arange(n, dtype=np.uint64) * np.uint64(1024 * 1024 * 3)
planes['dst_image'] = arena.data_ptr() + (offsets[:-1] * 3).astype(np.uint64)
planes['image_channels'] = 3
planes_dev = dv.upload_structs(planes)
lib = nv.lib()


def launch():
    nv.check(lib.vkb_grid_remap(
        dv.ptr(plan.pages_dev), dv.ptr(planes_dev), plan.n, plan.p_max, plan.c_max, plan.t_max,
        plan.s_cap, dv.ptr(plan.lattice_i), dv.ptr(plan.hinv), dv.ptr(plan.cell_box),
        dv.ptr(plan.cell_masks), dv.ptr(plan.tile_count), dv.ptr(plan.tile_off),
        dv.ptr(plan.tile_base), dv.ptr(plan.tile_slots), dv.ptr(plan.tile_headers), 3, 0, 0,
        dv.stream_ptr()), 'remap')


px = float(offsets[-1])
for _ in range(3):
    launch()
torch.cuda.synchronize()
ts = []
for _ in range(9):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts)
print(f'{os.environ.get("VKB_LIB", "in-tree")}: remap {t*1e3/n:7.2f} us/page, '
      f'{3 * (n * 1024 * 1024 + px) / t / 1e6:7.1f} GB/s algorithmic  (median {sorted(ts)[4]*1e3/n:.2f} us/page)')
# the plan itself
for _ in range(2):
    eng.plan_batch()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); eng.plan_batch(); e1.record(); torch.cuda.synchronize()
print(f'plan_batch (project, finalize, cells, masks, records; incl. host + one D2H): {e0.elapsed_time(e1)*1e3/n:.2f} us/page')
