import sys, time, torch, numpy as np
sys.path.insert(0,'/root/repo')
import bench
from vkit_b200.batch import GeometricBatch, distort_pages_host
names,configs=bench.sample_page_configs(0,256,256)
host=torch.randint(0,256,(256,1024,1024,3),dtype=torch.uint8).pin_memory()
out=torch.empty((256*1024*1024*3*2,),dtype=torch.uint8).pin_memory()
def seq():
    eng=GeometricBatch(names,configs,(1024,1024)); d=host.cuda(non_blocking=True); r=eng.run(d); n=int(r.image_arena.numel()); out[:n].copy_(r.image_arena,non_blocking=True); torch.cuda.synchronize()
def t(f,n=4):
    f(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print('sequential %.1f ms'%t(seq))
for chunk in (32,64,128):
    for th in (True,False):
        print('pipeline chunk',chunk,'thread',th,'%.1f ms'%t(lambda: distort_pages_host(names,configs,(1024,1024),host,out,chunk_pages=chunk,use_thread=th)))
# pieces
t0=time.perf_counter(); eng=GeometricBatch(names,configs,(1024,1024)); print('records %.1f ms'%((time.perf_counter()-t0)*1e3))
torch.cuda.synchronize(); t0=time.perf_counter(); d=host.cuda(non_blocking=True); torch.cuda.synchronize(); print('H2D %.1f ms'%((time.perf_counter()-t0)*1e3))
t0=time.perf_counter(); r=eng.run(d); torch.cuda.synchronize(); print('run %.1f ms'%((time.perf_counter()-t0)*1e3))
n=int(r.image_arena.numel()); t0=time.perf_counter(); out[:n].copy_(r.image_arena,non_blocking=True); torch.cuda.synchronize(); print('D2H %.1f ms'%((time.perf_counter()-t0)*1e3))
