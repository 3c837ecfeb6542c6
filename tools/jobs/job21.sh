#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/t.py <<'PY'
import ctypes, numpy as np, torch, sys
sys.path.insert(0, '/root/repo')
from vkit_b200 import _native as nv, device as dv
from vkit_b200.mechanism.distortion.photometric.blur import gaussian_kernel_u8
side = int(sys.argv[1])
src = dv.to_device(np.random.default_rng(0).integers(0, 256, (side, side, 3), dtype=np.uint8))
dst = torch.empty_like(src)
taps = gaussian_kernel_u8(5, 1.0)
arr = (ctypes.c_int32 * 5)(*taps)
rc = nv.lib().vkb_gaussian_blur_u8(dv.ptr(src), dv.ptr(dst), side, side, 3, arr, 5, dv.stream_ptr())
torch.cuda.synchronize()
print('ok', side, rc, int(dst.sum()))
PY
for s in 64 128 1024; do timeout 120 python /tmp/t.py $s 2>&1 | tail -2; done
timeout 300 compute-sanitizer --tool memcheck python /tmp/t.py 128 2>&1 | grep -v '^=========     Host Frame\|^=========         ' | head -40
