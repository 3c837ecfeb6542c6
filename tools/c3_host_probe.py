import sys, cProfile, pstats, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
seeds = list(range(1024))
w = bench.Workload3(1024, 0, 1024, np.arange(1024))
for _ in range(3): w.step(False)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(5): w.step(False)
t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print(f'host {(t1-t0)/5*1e3:.2f} ms/step, wall {(t2-t0)/5*1e3:.2f} ms/step')
pr=cProfile.Profile(); pr.enable()
for _ in range(5): w.step(False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
