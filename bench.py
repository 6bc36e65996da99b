#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE config 2.

    metric   leapfrog-steps x chains / sec   (plain HMC, D-dim isotropic Gaussian)
    workload config 2: D=1024 isotropic N(0,I), 256 chains per GPU, L=10, eps=0.05, S=1000 iterations, burn=0
    step     one pass of the hot path over one batch: ONE persistent-kernel launch advancing all chains of the rank
             through all S iterations (gibbs -> H -> L leapfrog steps -> H -> MH -> sample write), = C*S*L chain-steps

    python bench.py --gpus N --steps K --warmup W            (under torchrun for N > 1: one rank per GPU)
    python bench.py --impl reference ...                      (the reference's algorithm on the host cores)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D, C_PER_GPU, S, L, EPS = 1024, 256, 1000, 10, 0.05
WARP_INST_PER_LAUNCH = 1076063104          # config 2, E=4 K=1 geometry: ncu smsp__inst_executed.sum of one launch
                                           # (profiles/r1i_prof_hmc_run.summary.txt)
METRIC = 'leapfrog-steps x chains / sec'
UNIT = 'chain-steps/s'


def workload_config(world):
    """The `config` object both arms report: BASELINE config 2."""
    return {'workload': 'BASELINE config 2: D=1024 isotropic Gaussian, plain HMC, 256 chains/GPU, L=10, '
                        'eps=0.05, S=1000 iterations per step',
            'chains_per_gpu': C_PER_GPU, 'dim': D, 'L': L, 'iterations_per_step': S,
            'parallelism': 'chains sharded over %d GPU(s), no data-path collective' % world}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port: same Python loop + autograd as hamiltorch.sample) on host cores
# ----------------------------------------------------------------------------------------------------------
def _cpu_chain(args):
    seed, n_iter = args
    import torch as _t
    _t.set_num_threads(1)
    from hamiltorch_b200 import targets as T
    from oracle import hmc_oracle as O
    tgt = T.GaussianIso(D)
    _t.manual_seed(seed)
    init = 0.1 * _t.randn(D)
    t0 = time.perf_counter()
    O.sample_hmc(tgt, init, num_samples=n_iter, num_steps_per_sample=L, step_size=EPS)
    return time.perf_counter() - t0


def cpu_reference_rate(n_iter, procs):
    """`procs` independent chains (one process per host core, 1 torch thread each -- intra-op threads do not help at
    D=1024), each running n_iter iterations of config 2.  Returns chain-steps/s and wall."""
    import multiprocessing as mp
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_chain, [(1000 + i, n_iter) for i in range(procs)])
    wall = time.perf_counter() - t0
    return procs * n_iter * L / wall, wall


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_iter = args.cpu_iters
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_rate(max(2, n_iter // 10), cores)
    t_tot, steps_tot = 0.0, 0
    for _ in range(args.steps):
        rate, wall = cpu_reference_rate(n_iter, cores)
        t_tot += wall
        steps_tot += cores * n_iter * L
    value = steps_tot / t_tot
    sample = '%d chains (1 per core) x %d iterations x L=%d of config 2 per step' % (cores, n_iter, L)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # same workload as the B200 arm; each reference step is the bounded sample named in `sample` (independent chains
        # of config 2 through the reference's per-chain Python loop; the rate is per chain-step, so it extrapolates
        # linearly to the full 256 x 1000 job)
        'config': dict(workload_config(args.gpus), sample=sample),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every 20 ms while the timed regions run.  In-process NVML (nvidia_ml_py) on a
    daemon thread: one nvmlInit before the warm-up, then ~50 us queries -- no subprocess attaching to the driver while
    kernels are being launched (an `nvidia-smi -lms` loop did stall launches for milliseconds now and then).  Falls back
    to that loop only if NVML cannot be imported."""
    REASONS = (('hw_slowdown', 'nvmlClocksEventReasonHwSlowdown'),
               ('hw_thermal_slowdown', 'nvmlClocksEventReasonHwThermalSlowdown'),
               ('sw_thermal_slowdown', 'nvmlClocksEventReasonSwThermalSlowdown'),
               ('sw_power_cap', 'nvmlClocksEventReasonSwPowerCap'))
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.thread = index, None, None
        self.sm, self.mx, self.reasons, self.stop_flag = [], [], set(), False

    def _visible_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            ids = [v.strip() for v in vis.split(',') if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _loop(self, nv, h):
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for name, const in self.REASONS:
                    if r & getattr(nv, const):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(self.mx) if self.mx else None,
                    'reasons': sorted(self.reasons), 'samples': len(sm), 'source': 'nvml'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = [n for n, _ in self.REASONS]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi'}


# ----------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------
def run_b200_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    import hamiltorch_b200 as hb
    from hamiltorch_b200 import engine, targets as T, _native as N

    N.require_cuda()
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        # fork the CPU workers BEFORE this process touches CUDA or spins up torch's intra-op pool
        cores = os.cpu_count() or 1
        rate, wall = cpu_reference_rate(args.cpu_iters, cores)
        cpu_base = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                    'sample': '%d chains (1 per core) x %d iterations x L=%d of config 2, %.1f s wall'
                              % (cores, args.cpu_iters, L, wall)}
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)

    C = C_PER_GPU                                   # weak scaling: every rank owns 256 chains
    chain_offset = rank * C
    tgt = engine.NativeTarget(T.GaussianIso(D), dev)
    ld = N.padded_ld(D)
    g = torch.Generator().manual_seed(1234 + rank)
    q0_host = (0.1 * torch.randn(C, D, generator=g)).pin_memory()
    q0 = q0_host.to(dev)
    out = torch.empty((C, S, ld), dtype=torch.float32, device=dev)           # 1 GiB: 8x the 126 MB L2
    host_out = torch.empty((C, S, ld), dtype=torch.float32).pin_memory()
    stats_local = torch.zeros((max(args.steps, args.warmup, 1), C, 2), dtype=torch.float32, device=dev)
    stats = torch.empty((world,) + tuple(stats_local.shape), dtype=torch.float32, device=dev)

    def step(seed):
        return engine.hmc_run(tgt, q0, S, L, EPS, seed=seed, chain_offset=chain_offset, out=out, device=dev,
                              tuning=int(os.environ.get('HMCX_TUNING', '0')))

    def keep_stats(k, res):               # per-chain summary of step k (reject count, final step size), device side
        stats_local[k, :, 0].copy_(res.num_rejected)
        stats_local[k, :, 1].copy_(res.step_size)

    def gather_stats():                   # the run's single (tiny) collective: every rank's per-chain summaries
        if world > 1:
            dist.all_gather_into_tensor(stats.view(-1), stats_local.view(-1))
        else:
            stats[0].copy_(stats_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The clock sampler is started BEFORE the warm-up (NVML initialisation must not land inside a timed region); its
    # 20 ms polls then run through the warm-up and all timed regions.
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.2)
    barrier()

    # ---- warm-up: the W requested steps, then (still untimed) until the step time has settled ----
    # A fresh box runs its first launches several times slower for up to a few seconds (observed on B200 boxes right
    # after start-up: 5-9 ms instead of 1.5 ms per step, clocks already at maximum); W = 3 steps do not outlast that.
    for w in range(args.warmup):
        keep_stats(w, step(w))
    gather_stats()
    barrier()
    extra_warmup, best, stable, t_w = 0, float('inf'), 0, time.time()
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while extra_warmup < 400 and time.time() - t_w < 4.0 and stable < 8:
        w0.record()
        step(1000 + extra_warmup)
        w1.record()
        torch.cuda.synchronize()
        ms = w0.elapsed_time(w1)
        extra_warmup += 1
        best = min(best, ms)
        stable = stable + 1 if ms <= 1.1 * best else 0
    barrier()

    # ---- device-timed region: EXACTLY K steps, inputs resident in HBM ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 2)]
    ev[0].record()
    for k in range(args.steps):
        ev[1 + 2 * k].record()
        res = step(100 + k)
        ev[2 + 2 * k].record()
        keep_stats(k, res)
    gather_stats()
    ev[-1].record()
    barrier()
    t_total_ms = ev[0].elapsed_time(ev[-1])
    t_kernel_ms = sum(ev[1 + 2 * k].elapsed_time(ev[2 + 2 * k]) for k in range(args.steps)) / args.steps
    rejected = stats[:, :args.steps, :, 0].sum().item()

    # ---- e2e: public API, HOST buffers, H2D of the inputs and D2H of the result inside the timed region ----
    def e2e_step(seed):
        r = hb.sample_chains(T.GaussianIso(D), q0_host, num_samples=S, num_steps_per_sample=L, step_size=EPS,
                             rng='philox', seed=seed, chain_offset=chain_offset, out=out)
        host_out.copy_(r.samples_padded, non_blocking=True)
        return r

    # the same call with the reference's store_on_GPU=False contract (samplers.py:1008-1012): `out` is the pinned host
    # block, the kernel's retained-row stores go over PCIe while the chains run -- no device sample buffer, no D2H copy
    def e2e_stream_step(seed):
        return hb.sample_chains(T.GaussianIso(D), q0_host, num_samples=S, num_steps_per_sample=L, step_size=EPS,
                                rng='philox', seed=seed, chain_offset=chain_offset, out=host_out)

    def time_e2e(fn):
        fn(7)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = max(1, min(args.steps, 3))
        e0.record()
        for k in range(n):
            fn(200 + k)
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n

    t_e2e_copy_ms = time_e2e(e2e_step)
    t_e2e_stream_ms = time_e2e(e2e_stream_step)
    # both paths return the same bytes to the host; check it once on rank-local data (outside the timed regions)
    e2e_step(999)
    torch.cuda.synchronize()
    ref_rows = host_out[:, -1].clone()
    e2e_stream_step(999)
    torch.cuda.synchronize()
    assert torch.equal(ref_rows, host_out[:, -1]), 'streamed samples differ from the copied ones'
    t_e2e_ms = min(t_e2e_copy_ms, t_e2e_stream_ms)

    # ---- optional: cost of collecting every rank's samples with one NCCL all-gather (reported, not in `value`) ----
    allgather_ms = None
    if world > 1 and args.gather_samples:
        big = torch.empty((world, C, S, ld), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(big.view(-1), out.view(-1))
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        dist.all_gather_into_tensor(big.view(-1), out.view(-1))
        a1.record()
        barrier()
        allgather_ms = a0.elapsed_time(a1)
        del big

    # ---- streaming leapfrog kernel (the HBM-roofline form of samplers.leapfrog): state >> L2, L=1 ----
    Cs = 32768                                           # 32768 x 1024 fp32 = 128 MiB per array, 4 arrays
    qs = torch.randn(Cs, D, device=dev)
    ps = torch.randn(Cs, D, device=dev)
    qo, po = torch.empty_like(qs), torch.empty_like(ps)
    eps_vec = torch.full((Cs,), EPS, device=dev)
    lib, mass0 = N.load_library(), engine.NativeMass(None, D, dev)

    def stream_launch():          # straight through the C ABI with pre-allocated buffers: no host work between launches
        rc = lib.hmcx_leapfrog(tgt.ref(), mass0.ref(), N.ptr(qs), N.ptr(ps), N.ptr(eps_vec), Cs, ld, 1, N.ptr(qo),
                               N.ptr(po), None, None, N.stream_ptr(dev))
        N.check(rc, 'hmcx_leapfrog')

    for _ in range(3):
        stream_launch()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_s = 20
    s0.record()
    for _ in range(n_s):
        stream_launch()
    s1.record()
    torch.cuda.synchronize()
    t_stream_ms = s0.elapsed_time(s1) / n_s
    del qs, ps, qo, po
    clk = clocks.stop() if rank == 0 else None        # sampled across all three timed regions (all under load)

    # max over ranks of every timing
    t = torch.tensor([t_total_ms, t_kernel_ms, t_e2e_copy_ms, t_stream_ms, t_e2e_stream_ms], dtype=torch.float64,
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_total_ms, t_kernel_ms, t_e2e_copy_ms, t_stream_ms, t_e2e_stream_ms = t.tolist()
    t_e2e_ms = min(t_e2e_copy_ms, t_e2e_stream_ms)

    if rank == 0:
        units_per_step = world * C * S * L
        ms_per_step = t_total_ms / args.steps
        value = units_per_step / (ms_per_step * 1e-3)
        peak, peak_src = measured_peak_hbm()
        algo_bytes = C * S * D * 4 + C * D * 4          # per launch: every retained sample written once + init read once
        achieved = algo_bytes / (t_kernel_ms * 1e-3) / 1e9
        stream_bytes = Cs * D * 16                       # B_step = 16*D B per leapfrog-step x chain
        stream_gbs = stream_bytes / (t_stream_ms * 1e-3) / 1e9
        h2d, d2h = C * D * 4, C * S * ld * 4
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_mhz = float((clk or {}).get('sm_mhz') or 1965.0)
        issue_peak = n_sm * 4 * sm_mhz * 1e6 / 1e9
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': dict(workload_config(world), rng='in-kernel Philox4x32-10',
                           l2_policy='each step streams 1.0 GiB of samples (8x the 126 MB L2); no explicit flush',
                           extra_untimed_warmup_steps=extra_warmup),
            'roofline': {'bound': 'hbm', 'kernel': 'hmc_run_kernel<ISO,NONE,E=4,K=1,PHILOX,NUTS=0>', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture
                         # profiles/r1i_prof_hmc_run.summary.txt (1.14 MB read + 990.40 MB written)
                         'traffic': 991.5e6, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': algo_bytes, 'kernel_ms': t_kernel_ms,
                         'note': 'fused trajectory kernel: L=10 steps per 4*D bytes written, fp32-issue bound by design; '
                                 'see roofline_streaming for the HBM-bound form'},
            # what actually bounds the fused kernel: warp-instruction issue.  Instructions per launch are static for this
            # geometry (ncu smsp__inst_executed.sum, profiles/r1i_prof_hmc_run.summary.txt); time is measured live.
            'roofline_issue': {'bound': 'issue', 'kernel': 'hmc_run_kernel<ISO,NONE,E=4,K=1>',
                               'warp_instructions_per_launch': WARP_INST_PER_LAUNCH,
                               'achieved': WARP_INST_PER_LAUNCH / (t_kernel_ms * 1e-3) / 1e9,
                               'peak': issue_peak, 'unit': 'G warp-inst/s',
                               'frac': WARP_INST_PER_LAUNCH / (t_kernel_ms * 1e-3) / 1e9 / issue_peak,
                               'peak_source': '%d SMs x 4 schedulers x %.0f MHz (sampled under load)' % (n_sm, sm_mhz)},
            'roofline_streaming': {'bound': 'hbm', 'kernel': 'leapfrog_kernel<ISO,NONE> L=1, 32768x1024 state',
                                   'achieved': stream_gbs, 'peak': peak, 'unit': 'GB/s', 'frac': stream_gbs / peak,
                                   'algorithmic_bytes_per_launch': stream_bytes, 'kernel_ms': t_stream_ms,
                                   'chain_steps_per_s': Cs / (t_stream_ms * 1e-3)},
            'e2e': {'value': units_per_step / (t_e2e_ms * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': t_e2e_ms,
                    'path': ('hb.sample_chains(out=<pinned host block>): kernel streams the samples to the host'
                             if t_e2e_stream_ms <= t_e2e_copy_ms else
                             'hb.sample_chains(out=<device block>) + D2H copy of the samples'),
                    'ms_per_step_copy_path': t_e2e_copy_ms, 'ms_per_step_stream_path': t_e2e_stream_ms},
            'gpu_launches': args.steps,
            'accept_rate': 1.0 - rejected / (world * C * S * args.steps),
            'clocks': clk,
        }
        if allgather_ms is not None:
            line['allgather_samples_ms'] = allgather_ms
        if cpu_base is not None:
            line['cpu_baseline'] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-iters', type=int, default=2000, help='iterations per chain of the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--gather-samples', action='store_true', help='also time one NCCL all-gather of all samples')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit('launch with torchrun --nproc-per-node %d for --gpus %d' % (args.gpus, args.gpus))
    run_b200_arm(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
