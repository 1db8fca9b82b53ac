#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json:

    "achieved HBM GB/s (% of B200 peak): fp32 sum/axpy at 2^28 elems, 1/2/4/8 GPUs"

A step is one pass of the hot path over one batch per GPU:
    z = a*x + y      ElementwiseKernel (NVRTC glue on the FLAT tiler), 2^28 float32   [12 B/elem]
    s = x.sum()      single-pass full reduction, 2^28 float32                         [ 4 B/elem]
    (N > 1) all-reduce of the 0-d partial sum over NCCL / NVLink -- the path's one exchange step
Per-GPU work is fixed (weak scaling); `value` is the whole-job aggregate: algorithmic bytes
of all ranks / max-over-ranks device time.  Inputs (3 GiB per GPU) are far larger than the
126 MB L2, so no L2 flush is needed between iterations.

  python bench.py [--gpus N --steps K --warmup W]            our arm
  python bench.py --impl reference [...]                      the reference's CPU path (NumPy on the host cores)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ELEMS = 1 << 28
BYTES_AXPY = 12 * N_ELEMS
BYTES_SUM = 4 * N_ELEMS
BYTES_STEP = BYTES_AXPY + BYTES_SUM
METRIC = 'achieved HBM GB/s, fp32 axpy + sum at 2^28 elements per GPU'
UNIT = 'GB/s'
A = np.float32(1.5)


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# --------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region through NVML (in-process; nvidia-smi -lms is too
# coarse for a sub-second region)
# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8): 'hw_slowdown',
            getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
            getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4): 'sw_power_cap',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# --------------------------------------------------------------------------------------------
# the reference's CPU path: NumPy on the host cores (bounded sample of the same workload)
# --------------------------------------------------------------------------------------------
def cpu_step_factory(n, threads):
    """One CPU 'step' = NumPy axpy + sum over n float32, split over `threads` chunks
    (NumPy releases the GIL inside ufunc / reduction inner loops)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    rs = np.random.RandomState(0)
    x = (rs.rand(n) * 2 - 1).astype(np.float32)
    y = (rs.rand(n) * 2 - 1).astype(np.float32)
    z = np.empty_like(x)
    bounds = [(i * n // threads, (i + 1) * n // threads) for i in range(threads)]
    pool = ThreadPoolExecutor(threads) if threads > 1 else None

    def chunk(b):
        lo, hi = b
        oracle.numpy_axpy(A, x[lo:hi], y[lo:hi], z[lo:hi])
        return oracle.numpy_sum(x[lo:hi])

    def step():
        if pool is None:
            return chunk(bounds[0])
        return float(np.sum(list(pool.map(chunk, bounds))))
    return step


def cpu_baseline(threads, n=1 << 26, min_seconds=10.0, max_steps=200):
    step = cpu_step_factory(n, threads)
    step()
    t0 = time.perf_counter()
    k = 0
    while True:
        step()
        k += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or k >= max_steps:
            break
    gbs = 16.0 * n * k / dt / 1e9
    return {'value': round(gbs, 3), 'unit': UNIT, 'cores': threads,
            'kind': 'port',
            'sample': 'NumPy %s axpy (multiply+add, out=) + sum over 2^%d float32, %d steps in %.1f s, %d thread(s)'
                      % (np.__version__, int(np.log2(n)), k, dt, threads)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 1 << 26
    step = cpu_step_factory(n, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    gbs = 16.0 * n * args.steps / dt / 1e9
    base = {'value': round(gbs, 3), 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'sample': 'each step = NumPy axpy + sum over 2^26 float32 (1/4 of the per-GPU batch), '
                      'split over %d host threads' % threads}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': round(gbs, 3), 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(1e3 * dt / args.steps, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'fp32 axpy (z=a*x+y) + sum, 2^28 elements per GPU (CPU arm: 2^26-element sample per step)',
                   'l2': 'inputs exceed L2'},
        'cpu_baseline': base,
        'e2e': {'value': round(gbs, 3), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    import cupy_b200 as cp
    from cupy_b200 import distributed as cdist
    comm = None
    if world > 1:
        comm = cdist.init_process_group(world, rank, backend='nccl')

    # synthetic shard of this rank, generated in place on the device (seeded)
    g = torch.Generator(device='cuda')
    g.manual_seed(1234 + rank)
    tx = torch.rand(N_ELEMS, device='cuda', dtype=torch.float32, generator=g) * 2 - 1
    ty = torch.rand(N_ELEMS, device='cuda', dtype=torch.float32, generator=g) * 2 - 1
    x, y = cp.from_torch(tx), cp.from_torch(ty)
    z = cp.empty((N_ELEMS,), np.float32)
    s = cp.empty((), np.float32)
    axpy = cp.ElementwiseKernel('T a, T x, T y', 'T z', 'z = a * x + y', 'axpy')

    # N > 1: the partial sums are combined across the GPUs inside the reduction kernel's last block, through
    # NVLink peer memory (cupy_b200.distributed.sharded_sum -> b200_reduce_run_sharded); if the box cannot map
    # peer memory, sharded_sum falls back to one NCCL all-reduce of the 0-d partial
    fused = bool(comm is not None and comm.peer_exchange() is not None)

    def reduce_step():
        if comm is None:
            x.sum(out=s)
        else:
            cdist.sharded_sum(x, comm, out=s)

    def step():
        axpy(A, x, y, z)
        reduce_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    k = args.steps
    ev_a0 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    ev_a1 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    ev_s1 = [torch.cuda.Event(enable_timing=True) for _ in range(k)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for i in range(k):
        ev_a0[i].record()
        axpy(A, x, y, z)
        ev_a1[i].record()
        reduce_step()
        ev_s1[i].record()
    end.record()
    barrier()
    clocks = sampler.result()
    total_ms = start.elapsed_time(end)
    axpy_ms = float(np.mean([ev_a0[i].elapsed_time(ev_a1[i]) for i in range(k)]))
    sum_ms = float(np.mean([ev_a1[i].elapsed_time(ev_s1[i]) for i in range(k)]))

    t = torch.tensor([total_ms], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / k
    value = BYTES_STEP * world / (ms_per_step * 1e-3) / 1e9

    # ---- correctness guard inside the bench: the timed path produced the right numbers
    chk = float(s.get())
    want = float(tx.double().sum().item())
    if world > 1:
        w = torch.tensor([want], device='cuda', dtype=torch.float64)
        dist.all_reduce(w)
        want = float(w.item())
    assert abs(chk - want) <= 1e-5 * max(1.0, abs(want)) + 1e-2 * world, (chk, want)

    # ---- end to end through the public API with HOST buffers (pinned): every step copies its inputs
    #      host -> device, runs axpy + sum, and copies the result z (and the sum) device -> host.  The batch
    #      travels in chunks on three streams (cupy_b200.cuda.Stream: copy-in, compute, copy-out) so the two
    #      PCIe directions and the kernels overlap; device buffers are allocated once.
    e2e_steps = max(1, min(k, args.e2e_steps))
    hx, hy = cp.empty_pinned((N_ELEMS,), np.float32), cp.empty_pinned((N_ELEMS,), np.float32)
    hz = cp.empty_pinned((N_ELEMS,), np.float32)
    hx[:] = 0.25
    hy[:] = 0.5
    n_chunks = args.e2e_chunks
    cn = N_ELEMS // n_chunks
    dx, dy, dz = (cp.empty((N_ELEMS,), np.float32) for _ in range(3))
    parts = cp.empty((n_chunks,), np.float32)
    ds = cp.empty((), np.float32)
    s_in, s_cmp, s_out = cp.cuda.Stream(non_blocking=True), cp.cuda.Stream(non_blocking=True), cp.cuda.Stream(non_blocking=True)
    hsum = cp.empty_pinned((1,), np.float32)

    def e2e_step():
        for c in range(n_chunks):
            sl = slice(c * cn, (c + 1) * cn)
            dx[sl].set(hx[sl], stream=s_in)                      # H2D from pinned host memory
            dy[sl].set(hy[sl], stream=s_in)
            s_cmp.wait_event(s_in.record())
            with s_cmp:
                axpy(A, dx[sl], dy[sl], dz[sl])
                dx[sl].sum(out=parts[c])
            s_out.wait_event(s_cmp.record())
            dz[sl].get(stream=s_out, out=hz[sl], blocking=False)  # D2H of the step's result
        with s_cmp:
            parts.sum(out=ds)
            if comm is not None:
                comm.all_reduce(ds, ds, 'sum')
            ds.reshape(1).get(out=hsum, blocking=True)            # D2H of the reduction
        s_out.synchronize()
        return float(hsum[0])

    torch.cuda.synchronize()
    got = e2e_step()
    assert abs(got - 0.25 * N_ELEMS * world) <= 1e-5 * 0.25 * N_ELEMS * world, got
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_val = BYTES_STEP * world * e2e_steps / e2e_s / 1e9
    assert abs(float(hz[12345]) - float(np.float32(1.5) * np.float32(0.25) + np.float32(0.5))) < 1e-6
    assert abs(float(hz[N_ELEMS - 1]) - float(np.float32(1.5) * np.float32(0.25) + np.float32(0.5))) < 1e-6

    peak, peak_src = measured_peak()
    traffic = traffic_src = None          # DRAM bytes of one axpy launch from the committed ncu --set full capture
    try:
        with open(os.path.join(ROOT, 'profiles', 'axpy_traffic.json')) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
    except Exception:
        pass

    if rank == 0:
        line = {
            'metric': METRIC, 'value': round(value, 2), 'unit': UNIT, 'n_gpus': world, 'steps': k,
            'warmup': max(args.warmup, 3), 'ms_per_step': round(ms_per_step, 5), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {
                'workload': 'ElementwiseKernel axpy z=a*x+y (12 B/elem) + full sum (4 B/elem), float32, '
                            '2^28 elements per GPU' + (
                                '' if world == 1 else
                                '; partial sums combined across GPUs inside the reduction kernel over NVLink peer memory'
                                if fused else '; partial sums all-reduced over NCCL'),
                'elements_per_gpu': N_ELEMS, 'algorithmic_bytes_per_step_per_gpu': BYTES_STEP,
                'l2': 'inputs (3 GiB per GPU) exceed the 126 MB L2; no flush needed',
                'sharding': 'contiguous 1-D shards, one process per GPU' if world > 1 else 'single GPU',
            },
            'pct_of_peak': {'measured': round(100 * value / world / peak, 2),
                            'nominal_8000': round(100 * value / world / 8000.0, 2)},
            'elements_per_s': round(value * 1e9 / 16.0 * 2, 1),   # axpy + sum elements
            'roofline': {'bound': 'hbm', 'kernel': 'axpy (FlatTiler<4,4,4,256> + user op via NVRTC)',
                         'achieved': round(BYTES_AXPY / (axpy_ms * 1e-3) / 1e9, 2), 'peak': peak, 'unit': 'GB/s',
                         'frac': round(BYTES_AXPY / (axpy_ms * 1e-3) / 1e9 / peak, 4), 'traffic': traffic,
                         'traffic_source': traffic_src,
                         'peak_source': peak_src, 'avg_launch_ms': round(axpy_ms, 5),
                         'algorithmic_bytes_per_launch': BYTES_AXPY},
            'roofline_sum': {'bound': 'hbm', 'kernel': 'reduce_full_kernel<SumOp<float>,4,4>',
                             'achieved': round(BYTES_SUM / (sum_ms * 1e-3) / 1e9, 2), 'peak': peak, 'unit': 'GB/s',
                             'frac': round(BYTES_SUM / (sum_ms * 1e-3) / 1e9 / peak, 4),
                             'avg_launch_ms': round(sum_ms, 5), 'algorithmic_bytes_per_launch': BYTES_SUM},
            'clocks': clocks,
            'e2e': {'value': round(e2e_val, 3), 'unit': UNIT, 'h2d_bytes_per_step': 8 * N_ELEMS,
                    'd2h_bytes_per_step': 4 * N_ELEMS + 4, 'steps': e2e_steps,
                    'chunks': n_chunks,
                    'path': 'pinned host -> ndarray.set(stream) -> ElementwiseKernel axpy + sum -> ndarray.get(stream, out=pinned), '
                            '%d chunks pipelined over copy-in / compute / copy-out streams' % n_chunks},
            'gpu_launches': 2 * k,
        }
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline(1, min_seconds=args.cpu_seconds)
    # ---- every other BASELINE.json config, device-timed with an in-bench parity check, and the reference's own
    #      GPU kernels beside them (bench_configs.py).  Not part of `value`; the headline stays axpy + sum.
    if not args.no_configs:
        import bench_configs
        del x, y, z, tx, ty, dx, dy, dz, hx, hy, hz
        torch.cuda.empty_cache()
        c5 = bench_configs.run_c5(world, rank, comm, peak)
        darr = bench_configs.run_darray_check(world, rank, comm)
        if rank == 0:
            line['c5'] = c5
            line['darray'] = darr
        if world == 1:
            line['c1_cpu'] = bench_configs.c1_numpy()
            line['configs'], note = bench_configs.run_configs(peak, with_ref_gpu=not args.no_ref_gpu)
            if note:
                line['reference_gpu_unavailable'] = note
            bad = [e['name'] for e in line['configs'] if 'MISMATCH' in str(e.get('check'))]
            bad += ['c5 ' + k for k in ('sum', 'var') if 'MISMATCH' in c5[k]['check']]
            line['configs_parity'] = 'green' if not bad else {'failed': bad}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--e2e-chunks', type=int, default=16)
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='headline only: skip the configs table / c5 record')
    ap.add_argument('--no-ref-gpu', action='store_true', help='skip the reference-GPU baseline leg of the configs table')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
