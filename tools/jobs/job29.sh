python - <<'PY' 2>&1 | tail -40
import cProfile, pstats, sys
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from vkit_b200 import device as dv, _native as nv
rec=np.zeros(24,dtype=nv.BLEND_ITEM_DTYPE)
for _ in range(50): dv.upload_structs(rec)
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
for _ in range(2000): dv.upload_structs(rec)
torch.cuda.synchronize()
pr.disable()
st=pstats.Stats(pr); st.sort_stats('tottime').print_stats(12); st.print_callers('is_available')
PY
