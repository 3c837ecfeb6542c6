"""Raw pinned-memory PCIe bandwidth on this box: H2D alone, D2H alone, both at once."""
import time
import torch

n = 768 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device='cuda')
d_b = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for name, a, b in (('H2D', True, False), ('D2H', False, True), ('both', True, True)):
    run(a, b, 1)
    t = run(a, b)
    print(f'{name}: {t * 1e3:.1f} ms for {n / 1e6:.0f} MB each -> {n / t / 1e9:.1f} GB/s per direction')
