#!/usr/bin/env python
"""Headline benchmark: Gibbs sweeps/sec of Gaussian Bayesian Tensor Filtering.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|small]

One "step" = one Gibbs sweep = one ``resample(data)`` of the reference (nu2 -> sigma2 ->
Tau2 -> lam2 -> W -> V; factor.py:306-311, 112-128).  Workload at N=1: BASELINE.json
configs[1] ("c2": 4096 x 1024 x 64 x 3 replicates, nembeds=16, tf_order=2, 20 % NaN).

* ``value``  : sweeps/s with the data resident in HBM, timed with CUDA events on the
  engine's stream (inputs 2.4 GB >> 126 MB L2, so no L2 flush is needed).
* ``e2e``    : the same metric through the public ``run_gibbs`` path from HOST buffers:
  upload of Y, pre-reduction, K sweeps, device->host copy of every saved sample.
* ``roofline``: the dominant kernel of the sweep.  With the integer-tensor-core statistics path
  (stats_i8.cu, the default for Gaussian data at this size) that is ``i8gemm_kernel`` (tcgen05 int8):
  executed int8 operations / its CUDA-event time against twice the measured dense bf16 rate of
  MEASURED_PEAKS.json, with the FP64-equivalent rate of the product block and the HBM rate of the FP64
  linear block beside it; with ``BTF_STATS_NO_I8=1`` the FP64 DMMA kernels against the FP64 peak measured
  by the library's own micro-benchmark (MEASURED_PEAKS.json has no FP64 entry).
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle port of the reference path
  (oracle/btf_oracle.py) timed on a bounded sample of the same workload and
  extrapolated linearly in rows / columns (stated in ``sample``).
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
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of i8gemm_kernel (mean of the row and the column launch)
# from the ncu --set full capture of the same command (profiles/); None until captured
I8_TRAFFIC = {('c2', 1): 4.325e8}     # profiles/r1_ncu_i8_summary.txt: rows 0.340 + 0.005 GB, columns 0.273 + 0.247 GB

WORKLOADS = {
    # name: N, M, T, R, K, order, nan_frac
    'c2': dict(N=4096, M=1024, T=64, R=3, K=16, order=2, nan=0.2,
               label='Gaussian BTF 4096x1024x64x3 reps, nembeds=16, tf_order=2, 20% NaN'),
    'c1': dict(N=11, M=12, T=20, R=1, K=3, order=2, nan=0.0,
               label='examples/gaussian_tensor_filtering.py shape 11x12x20x1, nembeds=3, tf_order=2'),
    # BASELINE.json configs[4]; rows scale with the GPU count (8192 per GPU): at --gpus 8 this IS the
    # 65536 x 8192 x 128 x 2 configuration, smaller counts run its weak-scaled slice.  The shard is
    # generated on the device piecewise (1.1 TB of raw FP64 never exists on the host).
    'c5': dict(N=8192, M=8192, T=128, R=2, K=32, order=2, nan=0.2, device_data=True, weak=True,
               label='Gaussian BTF 8192*G x 8192 x 128 x 2 reps, nembeds=32, tf_order=2, 20% NaN (C5 at G=8)'),
    'small': dict(N=512, M=128, T=32, R=3, K=16, order=2, nan=0.2,
                  label='Gaussian BTF 512x128x32x3 reps, nembeds=16, tf_order=2, 20% NaN (smoke size)'),
}


# ----------------------------------------------------------------------------- synthetic data
def truth(cfg, seed=2):
    rng = np.random.default_rng(seed)
    N, M, T, K = cfg['N'], cfg['M'], cfg['T'], cfg['K']
    W = rng.normal(size=(N, K))
    W[np.triu_indices(min(N, K), k=1, m=K)] = 0
    jumps = rng.normal(size=(M, T, K)) * (rng.random((M, T, 1)) < 0.3)
    V = np.cumsum(jumps[:, ::-1], axis=1)[:, ::-1] * 0.5
    return W, np.ascontiguousarray(V)


def fill_rows(out, W, V, r0, r1, cfg, seed):
    """Y[r0:r1] = Mu + N(0,1) with element-wise NaN (SURVEY.md 8d, C2 generator)."""
    rng = np.random.default_rng([seed, r0])
    Mu = np.einsum('nk,mtk->nmt', W[r0:r1], V)
    blk = out[r0:r1]
    blk[...] = rng.standard_normal(blk.shape)
    blk += Mu[..., None]
    if cfg['nan'] > 0:
        blk[rng.random(blk.shape) < cfg['nan']] = np.nan
    if cfg['N'] <= 16:
        blk[: min(3, r1 - r0), :3] = np.nan       # held-out block of the shipped example


def make_host_data(cfg, rows=None, pinned=True, seed=2):
    from functionalmf_b200.engine import pinned_empty
    W, V = truth(cfg, seed)
    r0, r1 = rows if rows is not None else (0, cfg['N'])
    shape = (r1 - r0, cfg['M'], cfg['T'], cfg['R'])
    Y = pinned_empty(shape) if pinned else np.empty(shape)
    step = max(1, (64 << 20) // max(1, cfg['M'] * cfg['T'] * cfg['R'] * 8))
    Wl = W[r0:r1]
    for a in range(0, r1 - r0, step):
        b = min(r1 - r0, a + step)
        fill_rows(Y, Wl, V, a, b, cfg, seed + 17 * r0)
    return Y, W, V


def fill_device_shard(eng, cfg, rows, device):
    """Generate this rank's rows on the GPU in pieces (torch is only the generator here) and
    stream them through the pre-reduction (btf_set_data_gaussian_rows)."""
    import torch
    dev = torch.device('cuda', device)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    N, M, T, R, K = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K']
    V = torch.randn(M, T, K, generator=g, device=dev, dtype=torch.float64)
    V = (V * (torch.rand(M, T, 1, generator=g, device=dev) < 0.3)).flip(1).cumsum(1).flip(1) * 0.5
    Vf = V.reshape(M * T, K)
    r0, r1 = rows
    g.manual_seed(99 + r0)
    piece = max(1, (1 << 30) // (M * T * R * 8))
    first = True
    for a in range(r0, r1, piece):
        b = min(r1, a + piece)
        W = torch.randn(b - a, K, generator=g, device=dev, dtype=torch.float64)
        Y = (W @ Vf.T).reshape(b - a, M, T, 1) + torch.randn(b - a, M, T, R, generator=g, device=dev, dtype=torch.float64)
        if cfg['nan'] > 0:
            Y[torch.rand(Y.shape, generator=g, device=dev) < cfg['nan']] = float('nan')
        torch.cuda.synchronize()
        eng.set_data_gaussian_rows_device(Y.data_ptr(), a - r0, b - a, R, first)
        first = False
        del Y, W
    torch.cuda.empty_cache()


# ----------------------------------------------------------------------------- clocks
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, device=0):
        self.device, self.samples, self._stop, self._thr = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in out.stdout.strip().split(',')]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]),
                'power_w_max': max(float(s[2]) for s in self.samples), 'reasons': reasons,
                'samples': len(self.samples)}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline(cfg, budget_s=20.0, seed=2):
    """Time the CPU oracle port on a bounded sample of the workload and extrapolate."""
    from oracle import btf_oracle as O
    N, M, T, R, K, order = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K'], cfg['order']
    rng = np.random.default_rng(seed)
    full = N * M * T <= 2_000_000
    nr = N if full else min(N, 16)
    nc = M if full else min(M, 2)
    W, V = truth(cfg, seed)
    Delta = O.delta_matrix(T, order)
    RD = Delta.shape[0]
    Tau2 = rng.gamma(2.0, 1.0, size=(M, RD)) + 0.05
    # row sample: nr full rows; column sample: nc full columns
    Yr = np.empty((nr, M, T, R))
    fill_rows(Yr, W[:nr], V, 0, nr, cfg, seed)
    if full:
        Yc = Yr
    else:
        Yc = np.empty((N, nc, T, R))
        sub = dict(cfg)
        for a in range(0, N, 512):
            b = min(N, a + 512)
            fill_rows(Yc, W, V[:nc], a, b, sub, seed + 1)
    t0 = time.perf_counter()
    cnt_r, S_r, _ = O.prereduce(Yr)
    cnt_c, S_c, _ = O.prereduce(Yc)
    t_pre = time.perf_counter() - t0
    nu2, sigma2, lam2 = 1.0, 0.5, 0.1
    reps, tW, tV, tH = 0, 0.0, 0.0, 0.0
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        g = rng.gamma(2.0, size=(M, 4, RD))
        O.step_sigma2(W, 0.1, 0.1, 1.0)
        O.step_tau2(V, Delta, lam2, Tau2, Tau2, Tau2, g, K)
        O.step_lam2(V, Delta, Tau2, 1.0, np.ones(2), K)
        O.step_nu2(W[:nr], V, Yr, 0.1, 0.1, 1.0)          # residual pass over the row sample
        t1 = time.perf_counter()
        cw, sw = O.gaussian_weights(cnt_r, S_r, nu2)
        O.step_W(W[:nr].copy(), V, cw, sw, sigma2, rng.standard_normal((nr, K)))
        t2 = time.perf_counter()
        cw, sw = O.gaussian_weights(cnt_c, S_c, nu2)
        O.step_V(W, V[:nc].copy(), cw, sw, Delta, lam2, Tau2[:nc], rng.standard_normal((nc, T, K)), order)
        t3 = time.perf_counter()
        tH += t1 - t0; tW += t2 - t1; tV += t3 - t2
        reps += 1
        if time.perf_counter() - t_start > budget_s or reps >= 50:
            break
    tH, tW, tV = tH / reps, tW / reps, tV / reps
    # the nu2 residual inside tH covers nr rows only: scale that share with the W step's factor
    sweep_s = tH + (tW + 0.0) * (N / nr) + tV * (M / nc)
    threads = os.cpu_count() or 1
    sample = ('oracle port (numpy/LAPACK): hyper-parameter steps on the full factors + W step on %d of %d rows '
              '+ V step on %d of %d columns at full cross-dimensions, %d repetitions, extrapolated linearly; '
              'BLAS threads = all %d host cores' % (nr, N, nc, M, reps, threads))
    return {'value': 1.0 / sweep_s, 'unit': 'sweeps/s', 'cores': threads, 'kind': 'port', 'sample': sample,
            'sweep_seconds_est': sweep_s, 'prereduce_seconds_sample': t_pre}


def bind_to_gpu_numa_node(device):
    """Best effort: run this process on the CPUs local to the GPU's PCIe root so that the pinned
    host buffers (first-touch) live on the near NUMA node -- the H2D copy of Y runs at ~51 GB/s
    from the near node and ~18 GB/s from the far one on these hosts."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(device)
        bus = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open('/sys/bus/pci/devices/%s/local_cpulist' % bus) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return spec
    except Exception:
        pass
    return None


# ----------------------------------------------------------------------------- our arm
def bench_ours(args):
    import torch
    import torch.distributed as dist
    from functionalmf_b200.engine import Engine, fp64_peak, hbm_copy_gbs, pinned_empty
    from functionalmf_b200.distributed import Shard, agree_unique_id

    cfg = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
        os.environ['NCCL_DEBUG'] = 'WARN'          # keep stdout to the one JSON line
    if cfg.get('weak'):
        cfg['N'] = cfg['N'] * world
    if world != args.gpus:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run' % (args.gpus, world))
    torch.cuda.set_device(local)
    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    N, M, T, R, K, order = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K'], cfg['order']
    opts = dict(seed=1234, device=local)
    shard = None
    if world > 1:
        shard = Shard(rank, world, N, M)
        opts.update(shard.engine_options())
    rows = shard.rows if shard else (0, N)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    device_data = bool(cfg.get('device_data'))
    eng = Engine(N, M, T, nembeds=K, tf_order=order, **opts)
    if world > 1:
        eng.nccl_init(agree_unique_id())
    RD = eng.RD
    eng.init_state(127)
    eng.set('sigma2', [0.5]); eng.set('lam2', [0.1]); eng.set('nu2', [1.0])
    K_e2e = args.steps
    e2e = None
    e2e_launches = 0
    t_gen = time.perf_counter()
    if device_data:
        fill_device_shard(eng, cfg, rows, local)
        t_gen = time.perf_counter() - t_gen
        e2e = {'value': None, 'unit': 'sweeps/s', 'h2d_bytes_per_step': None, 'd2h_bytes_per_step': None,
               'note': 'shard generated on the device; no host copy of this workload exists'}
    else:
        Y, Wt, Vt = make_host_data(cfg, rows=rows)
        t_gen = time.perf_counter() - t_gen
        # ------------ end-to-end leg: host buffers -> run_gibbs-style segment -> host samples
        res_W = pinned_empty((K_e2e, N, K)); res_V = pinned_empty((K_e2e, M, T, K))
        res_T = pinned_empty((K_e2e, M, RD)); res_S = pinned_empty((K_e2e, 4))
        # untimed warm-up pass of the same path (CUDA module loading, graph instantiation, staging
        # allocation), then the timed pass: upload + pre-reduction + K sweeps + per-sweep D2H
        eng.set_data_gaussian(Y)
        eng.run_segment(min(max(3, args.warmup), K_e2e), 0, 1, 0, W=res_W, V=res_V, Tau2=res_T, scalars=res_S)
        eng.synchronize()
        barrier()
        l0 = eng.kernel_launches
        t0 = time.perf_counter()
        eng.set_data_gaussian(Y)
        t_up = time.perf_counter() - t0
        eng.run_segment(K_e2e, 0, 1, 0, W=res_W, V=res_V, Tau2=res_T, scalars=res_S)
        eng.synchronize()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_launches = eng.kernel_launches - l0
        assert np.all(np.isfinite(res_S)) and np.all(np.isfinite(res_V[-1]))
        d2h = (res_W[0].nbytes + res_V[0].nbytes + res_T[0].nbytes + res_S[0].nbytes)
        e2e = {'value': K_e2e / e2e_s, 'unit': 'sweeps/s', 'h2d_bytes_per_step': Y.nbytes / float(K_e2e),
               'd2h_bytes_per_step': d2h, 'seconds': e2e_s, 'upload_seconds': t_up, 'cpu_affinity': numa,
               'host_buffers': sorted(set(how for _, how in __import__('functionalmf_b200.engine', fromlist=['x']).PINNED_LOG)),
               'includes': 'H2D of Y (%.2f GB per rank, once) + pre-reduction, %d sweeps, D2H of W,V,Tau2,scalars '
                           'every sweep' % (Y.nbytes / 1e9, K_e2e)}

    os.sched_setaffinity(0, affinity0)        # the CPU baseline below uses every host core again

    # ---------------- device-resident leg (data already in HBM from the e2e leg)
    for _ in range(max(3, args.warmup)):
        eng.sweep(1)
    barrier()
    with ClockSampler(local) as clk:
        l0 = eng.kernel_launches
        ms = eng.sweep_timed(args.steps)
        launches = eng.kernel_launches - l0
        barrier()
    ms = max_over_ranks(ms)
    clocks = clk.summary()
    phases = eng.time_phases(3) if world == 1 else eng.time_phases(2)
    st_final = dict(sigma2=eng.get_scalar('sigma2'), lam2=eng.get_scalar('lam2'), nu2=eng.get_scalar('nu2'))

    out = None
    if rank == 0:
        cells = float(N) * M * T
        Lp = K * (K + 1) // 2
        stats_ms = phases['row_stats'] + phases['col_stats']
        flops_stats = 4.0 * cells * (Lp + K) / world        # per rank: both contractions, 2 flop per FMA
        dmma = fp64_peak(local, 1, 20000)
        dfma = fp64_peak(local, 0, 20000)
        peak = max(dmma, dfma)
        achieved = flops_stats / (stats_ms * 1e-3) / 1e12 if stats_ms > 0 else 0.0
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm_peak = None
        if os.path.exists(peaks_file):
            try:
                hbm_peak = json.load(open(peaks_file)).get('hbm_gbs')
            except Exception:
                hbm_peak = None
        hbm_src = 'MEASURED_PEAKS.json' if hbm_peak else 'fallback (B200_PROFILING.md)'
        hbm_peak = hbm_peak or 6650.0
        bytes_stats = 2.0 * cells * 9.0 / world
        # DRAM traffic per launch of the dominant kernel from the ncu --set full capture of this
        # workload (profiles/r1_ncu_stats_zpre_summary.txt: rows 2.504 GB read + 0.029 GB written,
        # columns 2.436 GB + 0.144 GB); algorithmic bytes per launch are cells * 9 B = 2.416 GB
        traffic = 2.5566e9 if (args.workload == 'c2' and world == 1) else None
        roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak if peak > 0 else None, 'traffic': traffic,
                    'kernel': 'stats_kernel (row + column sufficient-statistic contractions, FP64 DMMA)',
                    'flops_per_sweep_per_gpu': flops_stats, 'kernel_ms_per_sweep': stats_ms,
                    'peak_source': 'btf_fp64_peak micro-benchmark in this run (DMMA %.2f, DFMA %.2f TFLOP/s); '
                                   'MEASURED_PEAKS.json has no FP64 entry' % (dmma, dfma),
                    'hbm': {'achieved': bytes_stats / (stats_ms * 1e-3) / 1e9 if stats_ms > 0 else None,
                            'peak': hbm_peak, 'unit': 'GB/s', 'peak_source': hbm_src,
                            'bytes_per_sweep_per_gpu': bytes_stats}}
        gemm_ms = phases.get('row_i8gemm', 0.0) + phases.get('col_i8gemm', 0.0)
        if gemm_ms > 0:
            # integer-tensor-core path (stats_i8.cu): the dominant kernel is i8gemm_kernel, two launches per
            # sweep.  Executed work: 8 digit planes x L product columns against the counts,
            # 2 * (8 L) * rows * (padded contraction length) int8 operations per launch.
            nloc, nloc_pad = eng.nloc, -(-eng.nloc // 128) * 128
            P, Ppad = M * T, -(-(M * T) // 256) * 256
            ops = 2.0 * (8 * Lp) * (float(nloc) * Ppad + float(P) * nloc_pad)
            bf16 = None
            if os.path.exists(peaks_file):
                try:
                    bf16 = json.load(open(peaks_file)).get('bf16_tflops')
                except Exception:
                    bf16 = None
            i8_peak = 2.0 * (bf16 or 1634.5)
            ach = ops / (gemm_ms * 1e-3) / 1e12
            lin_ms = phases.get('row_linear', 0.0) + phases.get('col_linear', 0.0)
            roofline = {
                'bound': 'tensor', 'achieved': ach, 'peak': i8_peak, 'unit': 'TFLOP/s', 'frac': ach / i8_peak,
                'traffic': I8_TRAFFIC.get((args.workload, world)),
                'kernel': 'i8gemm_kernel (tcgen05.mma.kind::i8: exact digit-plane contraction of the product block '
                          'of the row and the column statistics; int8 operations counted as executed)',
                'ops_per_sweep_per_gpu': ops, 'kernel_ms_per_sweep': gemm_ms,
                'peak_source': '2 x the measured dense bf16 rate of MEASURED_PEAKS.json (%s TFLOP/s burst; the int8 '
                               'tensor rate of this part is twice its bf16 rate); no int8 entry there' % (bf16 or 'fallback 1634.5'),
                'algorithmic_fp64': {
                    'flops_per_sweep_per_gpu': 4.0 * cells * Lp / world,
                    'tflops_at_gemm_time': 4.0 * cells * Lp / world / (gemm_ms * 1e-3) / 1e12,
                    'fp64_dmma_peak': dmma,
                    'note': 'FP64 flops of the product block (SURVEY 8d) divided by the int8 GEMM time: what the exact '
                            'fixed-point formulation delivers against the FP64 pipe it replaces'},
                'linear_block': {'kernel': 'sf_kernel (FP64 DMMA, reads S once per contraction)',
                                 'ms_per_sweep': lin_ms, 'bytes_per_sweep_per_gpu': 2.0 * cells * 8.0 / world,
                                 'achieved_gbs': 2.0 * cells * 8.0 / world / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else None,
                                 'peak_gbs': hbm_peak, 'peak_source': hbm_src},
                'statistics_ms_per_sweep': stats_ms,
            }
        out = {
            'metric': 'Gibbs sweeps/sec', 'value': args.steps / (ms * 1e-3), 'unit': 'sweeps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak' if cfg.get('weak') else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': cfg['label'], 'shape': [N, M, T, R], 'nembeds': K, 'tf_order': order,
                       'nan_frac': cfg['nan'], 'l2': 'inputs (%.2f GB compact) larger than L2' % (cells * 9 / 1e9),
                       'parallelism': 'rows+cols sharded x%d' % world if world > 1 else 'single GPU',
                       'sweep': 'nu2,sigma2,Tau2,lam2,W,V (ref_compat lam2)'},
            'e2e': e2e,
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': roofline,
            'phases_ms': phases,
            'state': st_final,
            'e2e_gpu_launches': int(e2e_launches),
            'datagen_seconds': t_gen,
        }
        if world == 1 and not args.no_cpu_baseline and not device_data:
            out['cpu_baseline'] = cpu_baseline(cfg, budget_s=args.cpu_budget)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return out


# ----------------------------------------------------------------------------- reference arm
def bench_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    cfg = WORKLOADS[args.workload]
    per_step = max(5.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(cfg, budget_s=per_step, seed=2 + i)
        if i >= args.warmup:
            vals.append(base['sweep_seconds_est'])
    sweep_s = float(np.mean(vals))
    N, M, T, R, K, order = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K'], cfg['order']
    base['value'] = 1.0 / sweep_s
    return {
        'impl': 'reference', 'metric': 'Gibbs sweeps/sec', 'value': 1.0 / sweep_s, 'unit': 'sweeps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sweep_s * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': cfg['label'], 'shape': [N, M, T, R], 'nembeds': K, 'tf_order': order,
                   'nan_frac': cfg['nan']},
        'cpu_baseline': base,
        'e2e': {'value': 1.0 / sweep_s, 'unit': 'sweeps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    out = bench_reference(args) if args.impl == 'reference' else bench_ours(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == '__main__':
    main()
