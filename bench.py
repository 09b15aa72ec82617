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


# ----------------------------------------------------------------------------- CPU baseline (the unmodified reference)
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, 'functionalmf', 'factor.py'))


def import_reference():
    """The UNMODIFIED reference package from baseline/_ref (installed there by __graft_entry__.build() with
    `pip install --no-deps --target baseline/_ref /root/reference`; git-ignored, travels with gpurun) under the three
    import shims of oracle/shims: scikit-sparse (CHOLMOD -> banded LAPACK Cholesky), pypolyagamma, SharedArray are not
    installable offline.  Only this arm may execute anything under oracle/."""
    import warnings
    warnings.filterwarnings('ignore')
    for pth in (os.path.join(ROOT, 'oracle', 'shims'), REF_DIR):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    import functionalmf.factor as F
    assert os.path.abspath(F.__file__).startswith(REF_DIR), F.__file__
    return F


class ReferenceSample(object):
    """The reference's own sampler steps on a bounded sample of the workload:
    GaussianBayesianTensorFiltering._resample_W (factor.py:313-362) on `nr` rows at full M, T;
    _resample_V (factor.py:364-409) on `nc` columns at full N, T; _resample_nu2 on the row sample; sigma2 / Tau2 / lam2
    on the full column count.  Per-row and per-column costs do not depend on the other dimension's count (Python
    loops over rows / columns), so one sweep of the full tensor costs
        t_hyper + (t_nu2 + t_W) N / nr + t_V M / nc.
    Small workloads (C1) run model.resample(Y) on the whole tensor instead."""

    def __init__(self, cfg, nr=16, nc=1, seed=2):
        F = import_reference()
        self.cfg = cfg
        N, M, T, R, K, order = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K'], cfg['order']
        self.full = N * M * T <= 2_000_000
        W, V = truth(cfg, seed)
        np.random.seed(seed)
        kw = dict(nembeds=K, tf_order=order, sigma2_init=0.5, lam2_init=0.1, nu2_init=1.0)
        if self.full:
            self.nr, self.nc = N, M
            self.Y = np.empty((N, M, T, R))
            fill_rows(self.Y, W, V, 0, N, cfg, seed)
            self.model = F.GaussianBayesianTensorFiltering(N, M, T, **kw)
            return
        self.nr, self.nc = max(K, min(N, nr)), max(1, min(M, nc))        # factor.py:233 needs nrows >= nembeds
        self.Yr = np.empty((self.nr, M, T, R))
        fill_rows(self.Yr, W[:self.nr], V, 0, self.nr, cfg, seed)
        self.Yc = np.empty((N, self.nc, T, R))
        for a in range(0, N, 512):
            fill_rows(self.Yc, W, V[:self.nc], a, min(N, a + 512), cfg, seed + 1)
        self.rows = F.GaussianBayesianTensorFiltering(self.nr, M, T, **kw)
        self.cols = F.GaussianBayesianTensorFiltering(N, self.nc, T, **kw)

    def step(self):
        """One sampled sweep; returns (wall seconds of the sample, estimated seconds of the full sweep)."""
        N, M = self.cfg['N'], self.cfg['M']
        t0 = time.perf_counter()
        if self.full:
            self.model.resample(self.Y)
            dt = time.perf_counter() - t0
            return dt, dt
        m = self.rows
        m._resample_nu2(self.Yr)
        t1 = time.perf_counter()
        m._resample_sigma2(); m._resample_Tau2(); m._resample_lam2()
        t2 = time.perf_counter()
        m._resample_W(self.Yr)
        t3 = time.perf_counter()
        self.cols._resample_V(self.Yc)
        t4 = time.perf_counter()
        est = (t2 - t1) + ((t1 - t0) + (t3 - t2)) * (N / float(self.nr)) + (t4 - t3) * (M / float(self.nc))
        return t4 - t0, est

    def describe(self, reps):
        if self.full:
            return ('unmodified reference (baseline/_ref, functionalmf.factor.GaussianBayesianTensorFiltering.resample) on the '
                    'whole tensor, %d sweeps; shims: CHOLMOD -> banded LAPACK Cholesky' % reps)
        return ('unmodified reference (baseline/_ref): _resample_W on %d of %d rows and _resample_nu2 on the same rows at full '
                'M,T; _resample_V on %d of %d columns at full N,T; sigma2/Tau2/lam2 on all %d columns; %d repetitions; '
                'sweep = t_hyper + (t_nu2 + t_W) N/nr + t_V M/nc (extrapolated); single Python thread like the reference '
                '(nthreads is ignored, factor.py:36-37); shims: CHOLMOD -> banded LAPACK Cholesky'
                % (self.nr, self.cfg['N'], self.nc, self.cfg['M'], self.cfg['M'], reps))


def cpu_baseline_reference(cfg, budget_s=20.0, seed=2):
    """cpu_baseline for our arm: a few sampled sweeps of the unmodified reference within `budget_s` seconds."""
    smp = ReferenceSample(cfg, seed=seed)
    reps, est_sum, t_start = 0, 0.0, time.perf_counter()
    while True:
        _, est = smp.step()
        est_sum += est
        reps += 1
        if time.perf_counter() - t_start > budget_s or reps >= 200:
            break
    sweep_s = est_sum / reps
    return {'value': 1.0 / sweep_s, 'unit': 'sweeps/s', 'cores': 1, 'kind': 'reference', 'sample': smp.describe(reps),
            'sweep_seconds_est': sweep_s, 'host_cores_available': os.cpu_count()}


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


# ----------------------------------------------------------------------------- roofline
def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        return {}


def _ncu_traffic(kernel_key, workload, world):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel from the committed ncu
    --set full capture of this command (profiles/r2_traffic.json, written by tools/ncu_summary.py), else None."""
    try:
        tab = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
        return tab.get('%s/g%d' % (workload, world), {}).get(kernel_key)
    except Exception:
        return None


def build_roofline(args, cfg, eng, phases, ms_step, world, device):
    """`roofline` = the kernel family with the largest time share in THIS run (serialised per-phase CUDA-event times),
    plus the whole sweep against the HBM and FP64 bounds of BASELINE.md section 4, plus the other large families."""
    from functionalmf_b200.engine import fp64_peak, i8_peak
    N, M, T, R, K, order = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K'], cfg['order']
    cells = float(N) * M * T
    Lp = K * (K + 1) // 2
    n, kd = T * K, (order + 1) * K
    pk = _measured_peaks()
    hbm_peak = pk.get('hbm_gbs') or 6650.0
    hbm_src = 'MEASURED_PEAKS.json hbm_gbs' if pk.get('hbm_gbs') else 'fallback (B200_PROFILING.md)'
    dmma = fp64_peak(device, 1, 20000)
    dfma = fp64_peak(device, 0, 20000)
    fp64 = max(dmma, dfma)
    i8pk = i8_peak(device, 4000)
    gemm_ms = phases.get('row_i8gemm', 0.0) + phases.get('col_i8gemm', 0.0)
    lin_ms = phases.get('row_linear', 0.0) + phases.get('col_linear', 0.0)
    band_ms = phases.get('band_solve', 0.0)
    stats_ms = phases['row_stats'] + phases['col_stats']
    fam = {}
    # linear block (sf_kernel): reads S once per contraction, 8 B per cell
    if lin_ms > 0:
        byt = 2.0 * cells * 8.0 / world
        fam['sf_kernel'] = {'bound': 'hbm', 'achieved': byt / (lin_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                            'kernel': 'sf_kernel (FP64 linear block of the row and column statistics, 2 launches per sweep)',
                            'ms_per_sweep': lin_ms, 'algorithmic_bytes_per_sweep_per_gpu': byt,
                            'algorithmic_bytes_per_launch': byt / 2.0, 'peak_source': hbm_src,
                            'traffic': _ncu_traffic('sf_kernel', args.workload, world)}
    # product block (i8gemm2_kernel): executed int8 operations and the FP64 flops they replace
    if gemm_ms > 0:
        nloc, nloc_pad = eng.nloc, -(-eng.nloc // 128) * 128
        nall_pad = -(-N // 128) * 128 if world > 1 else nloc_pad
        P, Ppad = M * T, -(-(M * T) // 256) * 256
        ploc = eng.Mloc * T
        ntile = -(-Lp // 36)                       # 36 product columns x 7 digit planes = 252 (+4 idle) accumulator columns per tile
        r256 = lambda x: -(-x // 256) * 256
        ops = 2.0 * (256.0 * ntile) * (float(r256(nloc)) * Ppad + float(r256(ploc)) * nall_pad)
        alg = 4.0 * cells * Lp / world
        fam['i8gemm2_kernel'] = {'bound': 'tensor', 'achieved': ops / (gemm_ms * 1e-3) / 1e12, 'peak': i8pk, 'unit': 'TFLOP/s',
                                 'kernel': 'i8gemm2_kernel (tcgen05.mma.cta_group::2.kind::i8 + TMA, exact digit-plane contraction of the '
                                           'product block, 2 launches per sweep; achieved = int8 operations AS EXECUTED (7 digit planes, '
                                           '256 x 256 tiles incl. padding): an implementation choice - the algorithmic FP64 work is listed beside it',
                                 'ms_per_sweep': gemm_ms, 'executed_int8_ops_per_sweep_per_gpu': ops,
                                 'algorithmic_fp64_flops_per_sweep_per_gpu': alg,
                                 'algorithmic_fp64_tflops_at_this_time': alg / (gemm_ms * 1e-3) / 1e12, 'fp64_dmma_peak': dmma,
                                 'peak_source': 'btf_i8_peak in this run (resident-operand tcgen05 loop on every SM, %.0f Top/s); '
                                                'MEASURED_PEAKS.json has no int8 entry (2 x its bf16 burst = %.0f)'
                                                % (i8pk, 2.0 * (pk.get('bf16_tflops') or 1634.5)),
                                 'traffic': _ncu_traffic('i8gemm2_kernel', args.workload, world)}
    # band solve (band_lookahead_kernel): one launch, latency bound; factor traffic + FP64 flops
    if band_ms > 0:
        mloc = eng.Mloc
        byt = 3.0 * n * (kd + 1) * 8.0 * mloc + float(mloc) * T * (Lp + K) * 8.0
        flop = float(mloc) * n * kd * kd
        fam['band_lookahead_kernel'] = {'bound': 'hbm', 'achieved': byt / (band_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                        'kernel': 'band_lookahead_kernel (block-banded Cholesky + MVN draw per column; latency bound)',
                                        'ms_per_sweep': band_ms, 'algorithmic_bytes_per_launch': byt,
                                        'algorithmic_flops_per_launch': flop,
                                        'fp64_tflops': flop / (band_ms * 1e-3) / 1e12, 'fp64_peak': fp64,
                                        'fp64_frac': flop / (band_ms * 1e-3) / 1e12 / fp64 if fp64 > 0 else None,
                                        'peak_source': hbm_src, 'traffic': _ncu_traffic('band_lookahead_kernel', args.workload, world)}
    if not fam:
        # FP64 DMMA statistics path (BTF_STATS_NO_I8, PG models, small tensors)
        flop = 4.0 * cells * (Lp + K) / world
        fam['stats_kernel'] = {'bound': 'tensor', 'achieved': flop / (stats_ms * 1e-3) / 1e12 if stats_ms > 0 else 0.0, 'peak': fp64,
                               'unit': 'TFLOP/s', 'kernel': 'stats_kernel (FP64 DMMA sufficient statistics, 2 launches per sweep)',
                               'ms_per_sweep': stats_ms, 'algorithmic_flops_per_sweep_per_gpu': flop,
                               'peak_source': 'btf_fp64_peak in this run (DMMA %.2f, DFMA %.2f TFLOP/s)' % (dmma, dfma),
                               'traffic': _ncu_traffic('stats_kernel', args.workload, world)}
    for v in fam.values():
        v['frac'] = v['achieved'] / v['peak'] if v['peak'] else None
    top = max(fam, key=lambda k: fam[k]['ms_per_sweep'])
    roof = dict(fam[top])
    roof['dominant_by'] = 'largest serialised time share of the sweep in this run (%s: %.3f ms of %.3f ms phase total)' % (
        top, fam[top]['ms_per_sweep'], sum(v for k, v in phases.items() if k in
                                           ('nu2_or_pg', 'sigma2', 'tau2', 'lam2', 'row_stats', 'row_solve', 'col_stats', 'band_solve', 'comm')))
    roof['other_kernels'] = {k: v for k, v in fam.items() if k != top}
    # the whole sweep against BASELINE.md section 4
    B_alg = 2.0 * cells * 9.0 / world
    F_alg = (4.0 * cells * (Lp + K) + float(M) * n * kd * kd + float(N) * K ** 3 / 3.0) / world
    t = ms_step * 1e-3
    roof['sweep'] = {'ms_per_step': ms_step, 'B_alg_bytes_per_gpu': B_alg, 'F_alg_flops_per_gpu': F_alg,
                     'hbm_achieved_gbs': B_alg / t / 1e9, 'hbm_peak_gbs': hbm_peak, 'hbm_frac': B_alg / t / 1e9 / hbm_peak,
                     'fp64_equiv_tflops': F_alg / t / 1e12, 'fp64_peak_tflops': fp64,
                     'fp64_frac': F_alg / t / 1e12 / fp64 if fp64 > 0 else None,
                     'note': 'B_alg = 2 x cells x 9 B, F_alg = 4 cells (L+K) + M n kd^2 + N K^3/3 (BASELINE.md section 4); '
                             'fp64_frac > 1 is possible because the product block runs on the int8 tensor cores'}
    return roof


# ----------------------------------------------------------------------------- our arm
def bench_ours(args):
    import torch
    import torch.distributed as dist
    from functionalmf_b200.engine import Engine, pinned_empty
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
        out = {
            'metric': 'Gibbs sweeps/sec', 'value': args.steps / (ms * 1e-3), 'unit': 'sweeps/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
            'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak' if cfg.get('weak') else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(cfg, world),
            'parallelism': 'rows+cols sharded x%d' % world if world > 1 else 'single GPU',
            'e2e': e2e,
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': build_roofline(args, cfg, eng, phases, ms / args.steps, world, local),
            'phases_ms': phases,
            'phases_note': 'per-phase times are measured with the sweep serialised on one stream; in the timed run the '
                           'tensor-core product block, the HBM-bound linear block and the hyper-parameter steps overlap '
                           '(sum of phases > ms_per_step)',
            'state': st_final,
            'e2e_gpu_launches': int(e2e_launches),
            'datagen_seconds': t_gen,
        }
        if world == 1 and not args.no_cpu_baseline and not device_data:
            if reference_available():
                out['cpu_baseline'] = cpu_baseline_reference(cfg, budget_s=args.cpu_budget)
                # second, stronger CPU baseline: the vectorised numpy/LAPACK restatement
                out['cpu_baseline_port'] = cpu_baseline(cfg, budget_s=min(10.0, args.cpu_budget))
            else:
                out['cpu_baseline'] = cpu_baseline(cfg, budget_s=args.cpu_budget)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return out


# ----------------------------------------------------------------------------- reference arm
def workload_config(cfg, world):
    """The `config` dict both arms print (identical, so the driver's same_config check holds)."""
    N, M, T, R, K, order = cfg['N'], cfg['M'], cfg['T'], cfg['R'], cfg['K'], cfg['order']
    return {'workload': cfg['label'], 'shape': [N, M, T, R], 'nembeds': K, 'tf_order': order, 'nan_frac': cfg['nan'],
            'l2': ('inputs (%.2f GB compact) larger than L2' % (float(N) * M * T * 9 / 1e9)) if float(N) * M * T * 9 > 126e6
                  else 'inputs fit in L2 (latency-bound configuration: no flush makes it slower)',
            'sweep': 'nu2,sigma2,Tau2,lam2,W,V (ref_compat lam2)'}


def bench_reference(args):
    """--impl reference: the reference's own CPU implementation on the box's host cores.  Each step is one sampled
    sweep (see ReferenceSample); `ms_per_step` is the wall time of that bounded sample, `value` the sweeps/s of the
    FULL workload it implies (extrapolation stated in cpu_baseline.sample).  Without baseline/_ref the oracle port
    is timed instead (kind 'port')."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    cfg = dict(WORKLOADS[args.workload])
    if cfg.get('weak'):
        cfg['N'] = cfg['N'] * args.gpus
    if not reference_available():
        per_step = max(5.0, min(30.0, 120.0 / max(1, args.steps + args.warmup)))
        vals, wall, base = [], [], None
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            base = cpu_baseline(cfg, budget_s=per_step, seed=2 + i)
            if i >= args.warmup:
                vals.append(base['sweep_seconds_est']); wall.append(time.perf_counter() - t0)
        sweep_s, step_s = float(np.mean(vals)), float(np.mean(wall))
        base['value'] = 1.0 / sweep_s
    else:
        # size the sample so that warm-up + steps end within a few minutes
        total = max(1, args.steps + args.warmup)
        nc = 2 if total <= 12 else 1
        nr = 32 if total <= 12 else 16
        smp = ReferenceSample(cfg, nr=nr, nc=nc)
        vals, wall = [], []
        for i in range(total):
            dt, est = smp.step()
            if i >= args.warmup:
                vals.append(est); wall.append(dt)
        sweep_s, step_s = float(np.mean(vals)), float(np.mean(wall))
        base = {'value': 1.0 / sweep_s, 'unit': 'sweeps/s', 'cores': 1, 'kind': 'reference',
                'sample': smp.describe(args.steps), 'sweep_seconds_est': sweep_s, 'host_cores_available': os.cpu_count()}
    return {
        'impl': 'reference', 'metric': 'Gibbs sweeps/sec', 'value': 1.0 / sweep_s, 'unit': 'sweeps/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_s * 1e3,
        'ms_per_full_sweep_est': sweep_s * 1e3, 'extrapolated': sweep_s != step_s,
        'higher_is_better': True, 'scaling': 'weak' if cfg.get('weak') else 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(cfg, args.gpus),
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
