#!/usr/bin/env python
"""bench.py -- synthesized 1024x1024 pages/s through the camera-model geometric distortion.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): batch = 256 pages of 1024x1024 RGB per GPU, one camera_*
op per page cycling plane_only / cubic_curve / plane_line_fold / plane_line_curve, configs drawn
from the reference's policy generators with the per-page generator
default_rng(SeedSequence(133700).spawn(N)[i]).

One step = one pass of the hot path over the batch: lattice projection -> finalise -> per-cell
homographies + coverage masks + tile bins + per-tile candidate records -> fused remap.  `value`
times it with the inputs already in HBM; `e2e` times the same thing through the public batch API with HOST buffers
(config -> parameter blocks, H2D of the pages, kernels, D2H of the distorted pages).
`--impl reference` times the reference's own CPU algorithm (oracle port, cv2-backed when cv2
is importable) on the host cores.

N > 1: launched by torchrun, one rank per GPU; pages are independent, so ranks only share the
page seed list (broadcast) and counters (all_reduce) over NCCL -- weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BASE_SEED = 133700  # the reference pool's default rng seed (vkit/utility/pool.py:53)
PAGE_SHAPE = (1024, 1024)
# BASELINE.json: "synthesized 1024x1024 OCR pages/sec at 1/2/4/8 B200; remap HBM GB/s vs peak"
METRIC = 'synthesized_1024x1024_ocr_pages_per_s'
BATCH = 256
CAMERA_OPS = ('camera_plane_only', 'camera_cubic_curve', 'camera_plane_line_fold',
              'camera_plane_line_curve')
# project_camera (no page uses the MLS projector, so that kernel is not launched), finalize, layout,
# cells, masks, tile_base, tile_offsets, tile_records, remap (small-tile launch + large-tile launch)
KERNELS_PER_STEP = 10
# dram__bytes_read.sum + dram__bytes_write.sum of grid_remap_kernel, one 32-page launch under
# `ncu --set full` (profiles/r01_ncu_summary.md): both launches, 176.5 MB read + 85.8 MB written
# per 32 pages
TRAFFIC_PER_PAGE = 262.3e6 / 32
CPU_PAGES_PER_WORKER = 6  # bounded sample of the CPU arm: ~20 s of CPU work in total


def page_rngs(first: int, count: int, total: int):
    seqs = np.random.SeedSequence(BASE_SEED).spawn(total)
    return [np.random.default_rng(seqs[i]) for i in range(first, first + count)]


def sample_page_configs(first: int, count: int, total: int):
    """(op names, configs) for pages first..first+count-1 of a job of `total` pages."""
    from vkit_b200.mechanism.distortion_policy.geometric import camera as cam_policy
    factories = {
        'camera_plane_only': cam_policy.camera_plane_only_policy_factory,
        'camera_cubic_curve': cam_policy.camera_cubic_curve_policy_factory,
        'camera_plane_line_fold': cam_policy.camera_plane_line_fold_policy_factory,
        'camera_plane_line_curve': cam_policy.camera_plane_line_curve_policy_factory,
    }
    policies = {name: fac.create() for name, fac in factories.items()}
    names, configs = [], []
    for idx, rng in zip(range(first, first + count), page_rngs(first, count, total)):
        name = CAMERA_OPS[idx % len(CAMERA_OPS)]
        policy = policies[name]
        level = int(rng.integers(1, 11))
        generator = policy.config_generator_cls(policy.config_for_config_generator, level)
        names.append(name)
        configs.append(generator(PAGE_SHAPE, rng))
    return names, configs


def config_to_plain(config):
    import attrs
    out = {}
    for field in attrs.fields(type(config)):
        value = getattr(config, field.name)
        out[field.name] = config_to_plain(value) if attrs.has(type(value)) else value
    return out


# ---------------------------------------------------------------------------------------------
# NUMA placement of the rank (e2e leg: pinned host buffers next to the GPU's PCIe root)
# ---------------------------------------------------------------------------------------------
_ALL_CPUS = None


def bind_to_gpu_cpus(gpu_index: int):
    """Restrict this process to the CPUs NVML names as local to the GPU, so the pinned host
    buffers of the e2e leg (first touch) and the copy-issuing thread sit on the GPU's NUMA node.
    Returns the number of CPUs in the mask, or None when NVML cannot tell."""
    global _ALL_CPUS
    if os.environ.get('VKB_BENCH_NO_AFFINITY'):
        return None
    try:
        import pynvml as nvml
        _ALL_CPUS = os.sched_getaffinity(0)
        nvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        index = gpu_index
        if visible:
            ids = [v.strip() for v in visible.split(',') if v.strip()]
            if gpu_index < len(ids) and ids[gpu_index].isdigit():
                index = int(ids[gpu_index])
        nvml.nvmlDeviceSetCpuAffinity(nvml.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def unbind_cpus():
    """Back to the full CPU set (the CPU baseline leg uses every core)."""
    if _ALL_CPUS:
        try:
            os.sched_setaffinity(0, _ALL_CPUS)
        except OSError:
            pass


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, every ~5 ms; falls back
    to `nvidia-smi -lms 100` when pynvml is unavailable)."""
    QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.sm = []
        self.max_mhz = None
        self.reasons = set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.proc = None
        self.smi_samples = []

    def _visible_index(self):
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        if visible:
            ids = [v.strip() for v in visible.split(',') if v.strip()]
            if self.gpu_index < len(ids) and ids[self.gpu_index].isdigit():
                return int(ids[self.gpu_index])
        return self.gpu_index

    def _poll_nvml(self, nvml, handle):
        bits = {
            'hw_slowdown': nvml.nvmlClocksThrottleReasonHwSlowdown,
            'hw_thermal_slowdown': nvml.nvmlClocksThrottleReasonHwThermalSlowdown,
            'sw_thermal_slowdown': nvml.nvmlClocksThrottleReasonSwThermalSlowdown,
            'sw_power_cap': nvml.nvmlClocksThrottleReasonSwPowerCap,
        }
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nvml.nvmlDeviceGetClockInfo(handle, nvml.NVML_CLOCK_SM)))
                mask = nvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nvml
            nvml.nvmlInit()
            handle = nvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(nvml.nvmlDeviceGetMaxClockInfo(handle, nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll_nvml, args=(nvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self._visible_index()}', f'--query-gpu={self.QUERY}',
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump_smi, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump_smi(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 6:
                self.smi_samples.append(parts)

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            for s in self.smi_samples:
                if s[0].replace('.', '', 1).isdigit():
                    self.sm.append(float(s[0]))
                if s[1].replace('.', '', 1).isdigit():
                    self.max_mhz = max(self.max_mhz or 0.0, float(s[1]))
                for i in range(4):
                    if s[2 + i].lower().startswith('active'):
                        self.reasons.add(names[i])
        elif self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
        else:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling unavailable']}
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.sm)}


# ---------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference algorithm on the host cores)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(task):
    name, plain_config, seed = task
    from oracle import vkit_port as port
    port.use_cv2(True)
    try:
        import cv2
        cv2.setNumThreads(1)
    except ImportError:
        pass
    rng = np.random.default_rng(seed)
    image = rng.integers(0, 256, PAGE_SHAPE + (3,), dtype=np.uint8)
    t0 = time.perf_counter()
    out = port.grid_distort(name, plain_config, PAGE_SHAPE, image=image)
    return time.perf_counter() - t0, out['shape']


def cpu_reference_throughput(n_pages: int, cores: int, first: int = 0):
    """pages/s of the oracle port over `n_pages` pages on `cores` worker processes."""
    import multiprocessing as mp
    names, configs = sample_page_configs(first, n_pages, max(n_pages + first, BATCH))
    tasks = [(n, config_to_plain(c), BASE_SEED + first + i)
             for i, (n, c) in enumerate(zip(names, configs))]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, tasks[:cores])  # warm the workers (imports, cv2 init)
        t0 = time.perf_counter()
        per_page = pool.map(_cpu_worker, tasks)
        wall = time.perf_counter() - t0
    return n_pages / wall, wall, float(np.mean([p[0] for p in per_page]))


def cpu_backend_name():
    try:
        import cv2
        return f'oracle port of vkit grid path, cv2 {cv2.__version__} for remap/fillPoly/homography'
    except ImportError:
        return 'oracle port of vkit grid path, NumPy models (cv2 not importable)'


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cores = max(1, min(cores, 64))
    pages_per_step = cores * 2  # two pages per worker per step: a bounded sample of the workload
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_throughput(pages_per_step, cores)
    walls = []
    for step in range(args.steps):
        _, wall, _ = cpu_reference_throughput(pages_per_step, cores, first=step * pages_per_step)
        walls.append(wall)
    total_pages = pages_per_step * args.steps
    value = total_pages / sum(walls)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'pages/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * sum(walls) / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': workload_config(world),
        'cpu_baseline': {'value': value, 'unit': 'pages/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{pages_per_step} pages per step x {args.steps} steps, '
                                   f'{cores} worker processes, cv2 threads = 1; '
                                   + cpu_backend_name()},
        'e2e': {'value': value, 'unit': 'pages/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(world: int):
    return {
        'workload': 'camera_model geometric distort (camera_plane_only / cubic_curve / '
                    'plane_line_fold / plane_line_curve cycling), 1024x1024 RGB uint8, '
                    f'batch {BATCH} pages per GPU, image only',
        'batch_per_gpu': BATCH, 'page_shape': list(PAGE_SHAPE),
        'parallelism': f'page-sharded x{world} (no data-path collective)',
        'l2': 'inputs (805 MB per batch) larger than L2 (126 MB); no flush needed',
        'seed': BASE_SEED,
    }


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=20)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    parser.add_argument('--batch', type=int, default=BATCH)
    parser.add_argument('--skip-cpu-baseline', action='store_true')
    parser.add_argument('--kernel-only', action='store_true',
                        help='kernel experiments: no e2e leg, no CPU baseline (not a bench line)')
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    from vkit_b200 import _native
    from vkit_b200.batch import GeometricBatch

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    _native.lib()
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_cpus(local_rank)  # pinned buffers land on the GPU's NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    warmup = max(args.warmup, 3)
    batch = args.batch
    total_pages = batch * world
    first = rank * batch

    # page seed list: rank 0 owns it, everybody receives it (the only data-path-adjacent traffic)
    seeds = torch.arange(total_pages, dtype=torch.int64, device='cuda') + BASE_SEED
    if dist is not None:
        dist.broadcast(seeds, src=0)
    my_seeds = seeds[first:first + batch].cpu().numpy()

    names, configs = sample_page_configs(first, batch, total_pages)

    # synthetic pages: seeded random bytes, generated on the host once, resident in HBM
    height, width = PAGE_SHAPE
    host_pages = torch.empty((batch, height, width, 3), dtype=torch.uint8).pin_memory()
    host_np = host_pages.numpy()
    for i, seed in enumerate(my_seeds):
        rng = np.random.default_rng(int(seed))
        host_np[i] = rng.integers(0, 256, (height, width, 3), dtype=np.uint8)
    pages_dev = host_pages.cuda(non_blocking=True)
    torch.cuda.synchronize()

    engine = GeometricBatch(names, configs, PAGE_SHAPE)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only: inputs resident -----------------------------------------------------
    remap_events = []

    def step(timed: bool):
        # optimistic batch: projection, output layout (on the device), cells, masks, records and
        # the remap are queued without a host round trip; shapes are read after the timed loop
        return engine.run(pages_dev, optimistic=True,
                          launch_events=remap_events if timed else None)

    for _ in range(warmup):
        out = step(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    host_t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step(True)
    host_ms = 1000.0 * (time.perf_counter() - host_t0) / args.steps  # launch-side cost per step
    stop.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = start.elapsed_time(stop)
    # CUDA events recorded immediately around the remap launch, on the launching stream
    remap_ms = [a.elapsed_time(b) for a, b in remap_events]
    algorithmic_bytes = engine.algorithmic_bytes(channels=3)
    out_bytes = out.total_pixels * 3

    # ---- end to end: host buffers, public batch API -----------------------------------------
    host_out = torch.empty((out_bytes,), dtype=torch.uint8).pin_memory()

    from vkit_b200.batch import distort_pages_host

    def e2e_step():
        # configs -> parameter blocks (host), H2D of this step's pages, kernels, D2H of the
        # distorted pages; chunked over two streams so the three overlap
        _, _, offsets = distort_pages_host(names, configs, PAGE_SHAPE, host_pages, host_out,
                                           chunk_pages=32)
        return int(offsets[-1])

    e2e_steps = max(2, min(args.steps, 5))
    if args.kernel_only:
        d2h, e2e_ms = 0, float('nan')
    else:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d2h = e2e_step()
        barrier()
        e2e_ms = 1000.0 * (time.perf_counter() - t0) / e2e_steps

    # ---- reduce over ranks (max time, summed counters) --------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_ms, float(np.mean(remap_ms))], dtype=torch.float64,
                         device='cuda')
    counters = torch.tensor([batch * args.steps, algorithmic_bytes], dtype=torch.int64,
                            device='cuda')
    if dist is not None:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, remap_mean_ms = [float(x) for x in stats.cpu()]
    pages_done = int(counters[0])

    if rank == 0:
        peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(peaks_path):
            with open(peaks_path) as fin:
                peak = float(json.load(fin)['hbm_gbs'])
            peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)'
        else:
            peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
        achieved = algorithmic_bytes / (remap_mean_ms * 1e-3) / 1e9
        value = pages_done / (elapsed_ms * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'pages/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': elapsed_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
            'data': 'synthetic', 'config': workload_config(world),
            'clocks': clocks,
            'e2e': {'value': batch * world / (e2e_ms * 1e-3), 'unit': 'pages/s',
                    'h2d_bytes_per_step': int(host_pages.numel()),
                    'd2h_bytes_per_step': int(d2h),
                    'note': 'vkit_b200.batch.distort_pages_host(configs, pinned host pages) incl. '
                            'parameter-block build, H2D, kernels, D2H (32-page chunks; copy-in, '
                            'two work and copy-out streams); wall clock, max over ranks'},
            'gpu_launches': KERNELS_PER_STEP * args.steps,
            'host_ms_per_step': host_ms,
            'roofline': {
                'bound': 'hbm', 'kernel': 'grid_remap_kernel', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': TRAFFIC_PER_PAGE * batch, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': algorithmic_bytes,
                'launch_ms': remap_mean_ms,
                'note': '3 B x (src pixels + dst pixels) of the batch / mean duration of the '
                        'grid_remap_kernel launch (CUDA events recorded around the launch on its '
                        'stream); traffic = ncu dram bytes of one 32-page launch scaled to the '
                        'batch, see profiles/',
            },
        }
        line['config']['numa_local_cpus'] = numa_cpus
        if not (args.skip_cpu_baseline or args.kernel_only):
            unbind_cpus()
            cores = max(1, min(os.cpu_count() or 1, 64))
            n_pages = cores * CPU_PAGES_PER_WORKER
            cpu_value, cpu_wall, per_page = cpu_reference_throughput(n_pages, cores)
            line['cpu_baseline'] = {
                'value': cpu_value, 'unit': 'pages/s', 'cores': cores, 'kind': 'port',
                'sample': f'{n_pages} pages of the same workload, {CPU_PAGES_PER_WORKER} per worker process '
                          f'({per_page:.2f} s/page/core, wall {cpu_wall:.1f} s); '
                          + cpu_backend_name(),
            }
        print(json.dumps(line))

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
