#!/usr/bin/env python
"""
bench.py -- one JSON line for the displacement-statistics hot path.

A "step" is one pass of the hot path over one synthetic trajectory: stage 1 (unwrap + cumulative sum +
framework drift), stage 2 (MSD moments over atoms x time origins x dt) and stage 3 (covariance build,
reconditioning, log-likelihood factor, 1500-step x 32-walker ensemble MCMC -> D).

Workload at N=1: BASELINE.json configs[1] -- LLZO-like cell, 1,536 atoms (448 Li) x 20,000 steps,
100 time intervals, MSD + covariance + D.  With N>1 every rank gets a cell of that size (atoms are
sharded, weak scaling); the per-frame framework sums and the per-dt moment sums are all-reduced.

  python bench.py --gpus N --steps K --warmup W            (own arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU algorithm, oracle port)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(n_mobile=448, n_framework=1088, n_frames=20000, box=26.0, sigma=0.1, n_steps=100, time_step=0.001,
           step_skip=1, start_idx=10, seed=1002)
METRIC = 'displacement terms/s (atom*origin*dt), MSD+cov+D per trajectory'
UNIT = 'terms/s'
WORKLOAD = 'configs[1]: LLZO-like 1536 atoms (448 Li) x 20000 steps, 100 dt, MSD+covariance+emcee D'


def bench_config(args):
    """`config` of the JSON line: the workload, the same dictionary in both arms (what differs between the arms -- the
    bounded sample of the reference arm, the joins and the kernel of the GPU arm -- is in `cpu_baseline.sample` / `run`)."""
    n_all, n_mob, T = CFG['n_mobile'] + CFG['n_framework'], CFG['n_mobile'], CFG['n_frames']
    return {
        'workload': WORKLOAD,
        'per_gpu': f'{n_all} atoms ({n_mob} mobile) x {T} frames, f32 wrapped fractional coordinates',
        'sharding': 'atoms' if args.gpus > 1 else 'none',
        'l2': f'inputs ({n_all * (T + 1) * 12 / 1e6:.0f} MB f32 + {n_mob * 3 * T * 8 / 1e6:.0f} MB f64 per step) '
              'exceed the 126 MB L2; no flush needed',
        'analyses_in_flight': f'GPU arm: {args.in_flight} analyses on {args.in_flight} streams (`value` is their throughput, '
                              '`latency` one analysis alone); reference arm: one analysis per host core',
    }


def k2_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the stage-2 kernel on this workload, from the
    ncu --set full capture summarised in profiles/ (None if the summary is missing)."""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'r02_k2_tile_traffic.json')))
        return float(d['dram_bytes_read'] + d['dram_bytes_write'])
    except Exception:
        return None


def lag_grid(n_frames, n_steps):
    return np.linspace(1, n_frames, n_steps, dtype=int)


def n_terms(n_atoms, n_frames, lags):
    """terms = N_sel * sum_k (T - k + 1), the first-observation term included."""
    return int(n_atoms * np.sum(n_frames - lags + 1))


def make_inputs(seed, n_mobile, n_framework, n_frames, box, sigma, dtype=np.float32):
    """Seeded 3-D Gaussian random walk stored WRAPPED as fractional coordinates [N_all, T+1, 3]
    (frame 0 duplicated as the reference's front-ends do, kinisi/parser.py:323-324) in a cubic cell;
    framework atoms jitter about fixed sites; every atom shares a slow common drift
    that the framework correction removes."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_all = n_mobile + n_framework
    frac = np.empty((n_all, n_frames + 1, 3), dtype=dtype)
    start = rng.uniform(0, box, size=(n_all, 1, 3))
    chunk = 64
    drift = 0.01 * sigma * np.arange(n_frames)[None, :, None]  # whole-cell drift, shared by every atom
    for a0 in range(0, n_mobile, chunk):
        a1 = min(n_mobile, a0 + chunk)
        steps = rng.normal(0.0, sigma, size=(a1 - a0, n_frames, 3))
        steps[:, 0] = 0.0
        pos = start[a0:a1] + np.cumsum(steps, axis=1) + drift
        frac[a0:a1, 1:] = (pos / box) % 1.0
    for a0 in range(n_mobile, n_all, chunk):
        a1 = min(n_all, a0 + chunk)
        pos = start[a0:a1] + rng.normal(0.0, 0.05, size=(a1 - a0, n_frames, 3)) + drift
        frac[a0:a1, 1:] = (pos / box) % 1.0
    frac[:, 0] = frac[:, 1]
    latt = (np.eye(3) * box)[None]
    return frac, latt


# ------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the UNMODIFIED reference (oracle/_ref, copied there by oracle/make_ref.py, with the
# stand-ins of oracle/shims for the three packages this image lacks); the numpy oracle port only if oracle/_ref is missing
# ------------------------------------------------------------------------------------------------
REF_SAMPLE = dict(n_mobile=32, n_framework=64)  # atoms of one reference analysis (fixed: independent of the host)


def _cpu_shard_sums(args):
    from oracle import kinisi_oracle as ko
    dc_sel, lags = args
    S1, S2, _, _ = ko.raw_sums_streaming(dc_sel, lags)
    return S1, S2


def cpu_pass(frac, latt, n_mobile, n_framework, lags_cfg, pool=None, n_shards=1):
    """One pass of the reference's algorithm (oracle port) on the host.  Returns (terms, seconds)."""
    from oracle import kinisi_oracle as ko
    t0 = time.perf_counter()
    latt_full = np.broadcast_to(latt, (frac.shape[1], 3, 3))
    dc = ko.get_disp(frac.astype(np.float64), latt_full)
    idx = list(range(n_mobile))
    fw = list(range(n_mobile, n_mobile + n_framework))
    parsed = ko.parse(dc, idx, fw, CFG['time_step'], CFG['step_skip'], n_steps=lags_cfg)
    if pool is None:
        stats = ko.bootstrap_streaming(parsed, 'msd')
    else:
        bounds = np.linspace(0, n_mobile, n_shards + 1).astype(int)
        parts = pool.map(_cpu_shard_sums, [(parsed['dc_sel'][bounds[i]:bounds[i + 1]], parsed['lags'])
                                           for i in range(n_shards) if bounds[i + 1] > bounds[i]])
        S1 = sum(p[0] for p in parts)
        S2 = sum(p[1] for p in parts)
        cnt = n_mobile * (dc.shape[1] - parsed['lags'] + 1.0)
        n = S1 / cnt
        v = (S2 - S1 * n) / (cnt - 1) / parsed['n_o']
        stats = {'dt': parsed['delta_t'][:len(parsed['lags'])], 'n': n, 'v': v, 'n_o': parsed['n_o']}
    res = ko.bootstrap_gls(stats, stats['dt'][CFG['start_idx']], random_state=np.random.RandomState(1))
    dt = time.perf_counter() - t0
    assert np.isfinite(res['gradient']).all()
    return n_terms(n_mobile, dc.shape[1], parsed['lags']), dt


def reference_modules():
    """(kinisi.parser, kinisi.diffusion) of the unmodified reference under oracle/_ref, or None."""
    try:
        from oracle.make_ref import import_reference
        return import_reference()
    except Exception:
        return None


def reference_pass(mods, frac, latt, n_mobile, n_framework, n_steps=None, start_idx=None):
    """One analysis by the UNMODIFIED reference: Parser.get_disp -> Parser -> MSDBootstrap -> .diffusion
    (kinisi/parser.py:53-222, kinisi/diffusion.py:552-602, 305-436).  Returns (terms, seconds, D median)."""
    import warnings
    rp, rd = mods
    n_steps = CFG['n_steps'] if n_steps is None else n_steps
    start_idx = CFG['start_idx'] if start_idx is None else start_idx
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        cells = list(np.broadcast_to(latt, (frac.shape[1], 3, 3)))
        disp = rp.Parser.get_disp([frac.astype(np.float64)], cells, progress=False)
        p = rp.Parser(disp, np.arange(n_mobile), np.arange(n_mobile, n_mobile + n_framework), CFG['time_step'],
                      CFG['step_skip'], n_steps=n_steps, memory_limit=64., progress=False)
        b = rd.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
        b.diffusion(float(p.delta_t[start_idx]), progress=False, random_state=np.random.RandomState(1))
    dt = time.perf_counter() - t0
    return n_terms(n_mobile, disp.shape[1], p.time_intervals), dt, float(b.D.n)


_REF = {}


def _ref_worker_init():
    _REF['mods'] = reference_modules()
    _REF['inputs'] = make_inputs(CFG['seed'] + os.getpid() % 1000, REF_SAMPLE['n_mobile'], REF_SAMPLE['n_framework'],
                                 CFG['n_frames'], CFG['box'], CFG['sigma'])


def _ref_worker_step(_):
    frac, latt = _REF['inputs']
    if _REF['mods'] is None:
        t, dt = cpu_pass(frac, latt, REF_SAMPLE['n_mobile'], REF_SAMPLE['n_framework'], CFG['n_steps'])
        return t, dt
    t, dt, _d = reference_pass(_REF['mods'], frac, latt, REF_SAMPLE['n_mobile'], REF_SAMPLE['n_framework'])
    return t, dt


def _host_workers():
    """One reference analysis per worker; the reference is single-threaded, so all host cores are used by running
    independent analyses side by side (about 2.5 GB each at the sample size)."""
    cores = os.cpu_count() or 1
    try:
        import psutil
        cores = max(1, min(cores, int(psutil.virtual_memory().available // (3 << 30))))
    except Exception:
        pass
    return cores


def run_reference(args):
    """--impl reference: the unmodified reference on the host cores.  One step = every worker analyses one bounded sample
    of configs[1] (REF_SAMPLE atoms, the full 20 000 frames and 100 time intervals, covariance + sampler)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    for v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ.setdefault(v, '1')  # one analysis per core
    workers = _host_workers()
    kind = 'reference' if reference_modules() is not None else 'port'
    ctx = mp.get_context('fork')
    with ctx.Pool(workers, initializer=_ref_worker_init) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker_step, range(workers), chunksize=1)
        t0 = time.perf_counter()
        terms = 0
        for _ in range(args.steps):
            terms += sum(t for t, _dt in pool.map(_ref_worker_step, range(workers), chunksize=1))
        total = time.perf_counter() - t0
    value = terms / total
    sample = (f'{workers} independent analyses per step, each {REF_SAMPLE["n_mobile"]} of {CFG["n_mobile"]} mobile + '
              f'{REF_SAMPLE["n_framework"]} framework atoms x {CFG["n_frames"]} frames, {CFG["n_steps"]} dt, covariance + '
              f'1500-step x 32-walker sampler; ' +
              ('unmodified kinisi 1.1.1 (oracle/_ref) with the oracle/shims stand-ins for uravu / emcee / statsmodels'
               if kind == 'reference' else 'numpy oracle port (oracle/_ref missing)'))
    # configs[0] whole, one process: 192 Li + 64 framework atoms x 5000 steps
    cfg0 = None
    mods = reference_modules()
    if mods is not None:
        f0, l0 = make_inputs(1001, 192, 64, 5000, 15.0, 0.1, dtype=np.float64)
        t_0, s_0, d_0 = reference_pass(mods, f0, l0, 192, 64)
        cfg0 = {'workload': 'configs[0]: 192 Li x 5000 steps, MSD + Bayesian D', 'terms_per_s': t_0 / s_0, 'seconds': s_0,
                'cores': 1, 'D_median_cm2_s': d_0}
    line = {
        'impl': 'reference',
        'metric': METRIC,
        'value': value,
        'unit': UNIT,
        'n_gpus': args.gpus,
        'steps': args.steps,
        'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True,
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f64',
        'data': 'synthetic',
        'config': bench_config(args),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': workers, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'configs0': cfg0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML on a background thread (a query
    costs tens of microseconds); the nvidia-smi subprocess it replaces takes driver-wide locks for milliseconds per
    query, which stalled kernel launches inside a timed region that is only tens of milliseconds long."""
    BAD = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40, 'sw_power_cap': 0x4}

    def __init__(self, index, period_s=None):
        import threading
        if period_s is None:
            period_s = float(os.environ.get('KB2_BENCH_CLOCK_PERIOD_MS', '10')) * 1e-3  # (4 ms cost the host-bound loop 2-6 %: GIL and driver locks)
        self.samples, self.max_clock, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._thread = None
        self._smi = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._thread = threading.Thread(target=self._run, args=(period_s, ), daemon=True)
            self._thread.start()
        except Exception:
            self._nv = None
            self._start_smi(index)

    def _run(self, period_s):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                if not self.max_clock:  # a constant of the board: asked once
                    self.max_clock.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.BAD.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(period_s)

    def _start_smi(self, index):
        fields = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
                  'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
                  'clocks_event_reasons.sw_power_cap')
        try:
            fd, self._path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self._smi = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={fields}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self._path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self._smi = None

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        elif self._smi is not None:
            self._smi.terminate()
            try:
                self._smi.wait(timeout=5)
            except Exception:
                self._smi.kill()
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            try:
                for line in open(self._path):
                    f = [x.strip() for x in line.split(',')]
                    if len(f) < 7:
                        continue
                    self.samples.append(float(f[0]))
                    self.max_clock.append(float(f[1]))
                    for name, val in zip(names, f[3:7]):
                        if val.lower().startswith('active'):
                            self.reasons.add(name)
                os.unlink(self._path)
            except Exception:
                pass
        if not self.samples:
            return None
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': float(max(self.max_clock)),
                'reasons': sorted(self.reasons), 'samples': len(self.samples),
                'source': 'nvml' if self._thread is not None else 'nvidia-smi'}


def synth_coords_device(torch, dev, n_atoms, n_frames, box, sigma, seed, chunk=256):
    """Wrapped fractional coordinates [n_atoms, n_frames + 1, 3] float32 of a 3-D Gaussian random walk, generated on the
    device in chunks of atoms (float64 cumulative sum, wrapped, then rounded to float32); frame 0 duplicated as the
    reference's front-ends do (kinisi/parser.py:323-324)."""
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    out = torch.empty((n_atoms, n_frames + 1, 3), dtype=torch.float32, device=dev)
    for a0 in range(0, n_atoms, chunk):
        a1 = min(n_atoms, a0 + chunk)
        steps = torch.randn((a1 - a0, n_frames, 3), dtype=torch.float64, device=dev, generator=gen) * (sigma / box)
        steps[:, 0] = torch.rand((a1 - a0, 3), dtype=torch.float64, device=dev, generator=gen)
        pos = torch.cumsum(steps, dim=1)
        pos -= torch.floor(pos)
        out[a0:a1, 1:] = pos.to(torch.float32)
        del steps, pos
    out[:, 0] = out[:, 1]
    out.clamp_(0.0, 0.99999994)  # (a float32 rounding of 0.99999999 is 1.0: keep the coordinates inside [0, 1))
    return out


K2_NAMES = {0: 'direct', 1: 'tma-ring', 2: 'arithmetic-runs', 3: 'lag-window', 4: 'tiled-lag-window'}


def named_shapes(args, torch, td, dev, world, rank, fp64_peak_tflops, hbm_peak_gbs):
    """Timed sub-records for the other BASELINE.json configurations, after the headline: configs[2] (MSTD, 10 000 ions x
    100 000 steps) and configs[4] (MSD, 100 000 atoms x 50 000 steps) STRONG-scaled over the ranks (atoms / N per rank,
    device-resident float32 input generated on the device, the cumulative displacements streamed in 2 GB blocks);
    configs[3] (MSCD, 1000 + 1000 ions, 1000 used time intervals -> 1000 x 1000 covariance) and configs[0] (192 Li x 5000
    steps through ASEParser) on rank 0.  Every record: wall time per analysis from CUDA events (max over ranks),
    stage times, terms/s, the FP64 fraction of the moment kernel and the HBM rate of the displacement kernel."""
    from kinisi_b200 import diffusion, engine
    from kinisi_b200 import parser as kparser
    from kinisi_b200.parser import Parser
    out = {}
    k1 = {'ms': [], 'bytes': 0}
    k2 = {'ms': []}
    raw_k2, raw_k1 = engine.self_moments, engine.unwrap_cumsum

    def t_k2(dc, n_frames, lg, mask=7, variant=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = raw_k2(dc, n_frames, lg, mask, variant)
        e1.record()
        k2['ms'].append((e0, e1))
        return r

    def t_k1(coords, mode, latt, latt_inv, latt_stride, idx, step_sums, step_scale, pitch, out=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = raw_k1(coords, mode, latt, latt_inv, latt_stride, idx, step_sums, step_scale, pitch, out=out)
        e1.record()
        k1['ms'].append((e0, e1))
        k1['bytes'] += int(idx.numel()) * (coords.shape[1] * 3 * coords.element_size() + 3 * pitch * 8)
        return r

    def sync_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0])

    def strong(name, kind, n_total, T, box, start_idx, passes=2):
        n_local = n_total // world + (1 if rank < n_total % world else 0)
        need = n_local * (T + 1) * 12 + kparser.STREAM_BLOCK_BYTES + (1 << 30)
        free, _total = torch.cuda.mem_get_info(dev)
        if need > free:
            out[name] = {'skipped': f'{need / 1e9:.0f} GB needed per GPU at {world} GPU(s), {free / 1e9:.0f} GB free'}
            return
        coords = synth_coords_device(torch, dev, n_local, T, box, 0.1, 5000 + 17 * rank)
        latt = (np.eye(3) * box)[None]
        lags = lag_grid(T, CFG['n_steps'])
        used = lags if kind == 'msd' else lags[:len(lags) // 2]
        terms = n_terms(n_total, T, used)
        start_dt = float(lags[start_idx] * CFG['time_step'])
        pkw = dict(n_steps=CFG['n_steps'], memory_limit=64., progress=False, distributed=world > 1,
                   global_counts=(n_total, 0) if world > 1 else None)
        engine.self_moments, engine.unwrap_cumsum = t_k2, t_k1
        try:
            res = []
            for it in range(passes + 1):
                k1['ms'].clear(), k2['ms'].clear()
                k1['bytes'] = 0
                if world > 1:
                    td.barrier()
                torch.cuda.synchronize()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
                p = Parser(Parser.get_disp(coords, latt, progress=False, device=dev), np.arange(n_local), [], CFG['time_step'], 1, **pkw)
                if kind == 'msd':
                    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
                else:
                    b = diffusion.MSTDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
                ev[1].record()
                if kind == 'msd':
                    b.diffusion(start_dt, progress=False)
                else:
                    b.jump_diffusion(start_dt, progress=False)
                _ = b.flatchain
                ev[2].record()
                torch.cuda.synchronize()
                if it == 0:
                    continue  # warm-up (plans, allocator)
                rec = {'ms': sync_max(ev[0].elapsed_time(ev[2])), 'stage12_ms': sync_max(ev[0].elapsed_time(ev[1])),
                       'stage3_ms': sync_max(ev[1].elapsed_time(ev[2])),
                       'k1_ms': sync_max(sum(a.elapsed_time(z) for a, z in k1['ms'])),
                       'k2_ms': sync_max(sum(a.elapsed_time(z) for a, z in k2['ms']))}
                res.append(rec)
            best = min(res, key=lambda r: r['ms'])
            streamed = bool(getattr(p, '_streamed', False))
            coef = b.D if kind == 'msd' else b.D_J
            out[name] = {
                'kind': kind, 'atoms': n_total, 'frames': T, 'atoms_per_gpu': n_local, 'scaling': 'strong',
                'dc_streamed': streamed, 'k2_kernel': K2_NAMES[engine.choose_variant(np.asarray(used, dtype=np.int32), T)],
                'ms_per_analysis': best['ms'], 'terms': terms,
                'terms_per_s': terms / (best['ms'] * 1e-3), 'stage_ms': {'stage1+2': best['stage12_ms'], 'stage3': best['stage3_ms']},
                'k2_ms': best['k2_ms'], 'k2_fp64_frac': 16.0 * terms / world / (best['k2_ms'] * 1e-3) / 1e12 / fp64_peak_tflops,
                'k1_ms': best['k1_ms'], 'k1_hbm_gbs': k1['bytes'] / (best['k1_ms'] * 1e-3) / 1e9,
                'k1_hbm_frac': k1['bytes'] / (best['k1_ms'] * 1e-3) / 1e9 / hbm_peak_gbs,
                'coefficient_median_cm2_s': float(coef.n),
            }
            if kind == 'msd':  # (the collective coefficient of one realisation of independent walkers is a noisy estimate)
                out[name]['true_cm2_s'] = 3 * 0.1**2 / CFG['time_step'] / (2e4 * 3)
        finally:
            engine.self_moments, engine.unwrap_cumsum = raw_k2, raw_k1
        del coords, p, b
        torch.cuda.empty_cache()

    strong('configs[2]', 'mstd', 10000, 100000, 60.0, 5)
    strong('configs[4]', 'msd', 100000, 50000, 120.0, 10)

    if rank == 0:
        # configs[3]: 1000 cations (+1) and 1000 anions (-1), 20 000 steps, n_steps = 2000 -> 1000 used time intervals
        from kinisi_b200.analyze import ConductivityAnalyzer, DiffusionAnalyzer
        n_ion, T, box = 2000, 20000, 40.0
        coords = synth_coords_device(torch, dev, n_ion, T, box, 0.1, 7003)
        latt = (np.eye(3) * box)[None]
        q = np.where(np.arange(n_ion) < n_ion // 2, 1.0, -1.0)
        recs = []
        for it in range(3):
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            p = Parser(Parser.get_disp(coords, latt, progress=False, device=dev), np.arange(n_ion), [], CFG['time_step'], 1,
                       n_steps=2000, memory_limit=64., progress=False)
            b = diffusion.MSCDBootstrap(p.delta_t, p.disp_3d, q, p._n_o, progress=False)
            ev[1].record()
            b.conductivity(float(b.dt[100]), 500.0, box**3, progress=False)
            _ = b.flatchain
            ev[2].record()
            torch.cuda.synchronize()
            recs.append((ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
        best = min(recs[1:])
        lags = np.asarray(p.disp_3d.lags[:len(p.disp_3d.lags) // 2])
        out['configs[3]'] = {'kind': 'mscd', 'atoms': n_ion, 'frames': T, 'time_intervals_used': int(b.dt.size),
                             'covariance_n': int(b.dt.size - 100),
                             'k2_kernel': K2_NAMES[engine.choose_variant(lags.astype(np.int32), T)], 'ms_per_analysis': best[0],
                             'stage_ms': {'stage1+2': best[1], 'stage3 (cov_build + eigh + gls_prepare + sampler)': best[2]},
                             'self_terms_per_s': n_terms(n_ion, T, lags) / (best[1] * 1e-3), 'sigma_median_mS_cm': float(b.sigma.n)}
        del coords, p, b
        torch.cuda.empty_cache()
        # configs[0]: 192 Li (+ 64 framework) x 5000 steps through ASEParser (host objects, float64) -> MSD -> D
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        from duck_atoms import trajectory_from_arrays
        f0, l0 = make_inputs(1001, 192, 64, 5000, 15.0, 0.1, dtype=np.float64)
        traj = trajectory_from_arrays(['Li'] * 192 + ['O'] * 64, f0[:, 1:], np.broadcast_to(l0, (5000, 3, 3)))
        walls = []
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            an = DiffusionAnalyzer.from_ase(traj, parser_params=dict(specie='Li', time_step=CFG['time_step'], step_skip=1,
                                                                     progress=False), uncertainty_params={'progress': False})
            an.diffusion(an.dt[10], {'progress': False})
            d_med = an.D.n
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t0)
        lags0 = lag_grid(5000, 100)
        out['configs[0]'] = {'kind': 'msd via ASEParser (host walk of 5000 frame objects included)', 'atoms': 192, 'frames': 5000,
                             'wall_ms_per_analysis': 1e3 * min(walls[1:]), 'terms_per_s': n_terms(192, 5000, lags0) / min(walls[1:]),
                             'D_median_cm2_s': float(d_med)}
    return out


def run_own(args):
    import torch
    import torch.distributed as td
    from kinisi_b200 import diffusion, engine
    from kinisi_b200 import dist as kdist
    from kinisi_b200.parser import Parser

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        td.init_process_group('nccl', device_id=dev)
    engine.require_cuda()
    if args.variant is not None:
        engine.DEFAULT_VARIANT = args.variant

    n_mob, n_fw, T = CFG['n_mobile'], CFG['n_framework'], CFG['n_frames']
    frac, latt = make_inputs(CFG['seed'] + rank, n_mob, n_fw, T, CFG['box'], CFG['sigma'])
    lags = lag_grid(T, CFG['n_steps'])
    terms_rank = n_terms(n_mob, T, lags)
    start_dt = float(lags[CFG['start_idx']] * CFG['time_step'] * CFG['step_skip'])
    idx = np.arange(n_mob)
    fw = np.arange(n_mob, n_mob + n_fw)
    host = torch.from_numpy(frac).pin_memory()
    resident = host.to(dev)
    pkw = dict(n_steps=CFG['n_steps'], memory_limit=64., progress=False, distributed=world > 1,
               global_counts=(world * n_mob, world * n_fw) if world > 1 else None)
    gkw = dict(progress=False)

    k2_events = []
    raw_self_moments = engine.self_moments

    def timed_self_moments(dc, n_frames, lg, mask=7, variant=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = raw_self_moments(dc, n_frames, lg, mask, variant)
        e1.record()
        k2_events.append((e0, e1, dc.shape[0]))
        return out

    engine.self_moments = timed_self_moments

    stage_events = []
    lanes = [torch.cuda.Stream(dev) for _ in range(max(1, args.in_flight))]

    def step(coords):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        traj = Parser.get_disp(coords, latt, progress=False, device=dev)
        p = Parser(traj, idx, fw, CFG['time_step'], CFG['step_skip'], **pkw)
        ev[1].record()
        b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
        ev[2].record()
        b.diffusion(start_dt, **gkw)  # ends with the device->host read of the flat chain
        ev[3].record()
        stage_events.append(ev)
        b._bench_parser = p
        return b

    def timed(coords, steps):
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if args.in_flight > 1:
            # consecutive analyses on alternating streams: everything of one analysis is in order on its stream, the
            # regression of one (a single CTA, then a single warp) runs beside stages 1-2 of the next
            main = torch.cuda.current_stream(dev)
            for st in lanes:
                st.wait_stream(main)
            held = [None] * args.in_flight
            for i in range(steps):
                with torch.cuda.stream(lanes[i % args.in_flight]):
                    held[i % args.in_flight] = step(coords)
            for st in lanes:
                main.wait_stream(st)
            b = held[(steps - 1) % args.in_flight]
        else:
            for _ in range(steps):
                b = step(coords)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            td.barrier()
        ms = e0.elapsed_time(e1)  # device time is the clock; the wall time is reported beside it
        t = torch.tensor([ms, 1e3 * wall], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0]), float(t[1]), b

    # warm-up (untimed), both paths
    for _ in range(max(args.warmup, 3)):
        step(resident)
    step(host)
    torch.cuda.synchronize()
    if args.in_flight > 1:  # the streams of the timed loops have their own allocator pools: warm them too
        timed(resident, 2 * args.in_flight)
        timed(host, 2 * args.in_flight)

    # FP64-pipe peak of this device, measured live (16 independent DFMA chains per thread)
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    info = engine.device_info()
    grid, iters = info['sm_count'] * 8, 4096
    engine.fp64_peak(grid, 64, sink)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.fp64_peak(grid, iters, sink)
    e1.record()
    torch.cuda.synchronize()
    fp64_peak_tflops = 2.0 * grid * 256 * iters * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12

    launches0 = engine.LAUNCHES['n']
    k2_events.clear()
    stage_events.clear()
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get('KB2_BENCH_NO_CLOCKS') else None
    ms_total, wall_total, b = timed(resident, args.steps)
    launches = engine.LAUNCHES['n'] - launches0
    k2_ms = [a.elapsed_time(z) for a, z, n in k2_events if n == n_mob]
    stages = np.array([[ev[i].elapsed_time(ev[i + 1]) for i in range(3)] for ev in stage_events]).mean(axis=0)
    # the stage-2 kernel timed back to back on the last step's displacement array (215 MB > L2, so every
    # launch streams it from HBM again); free of the host launch gaps the in-step interval can contain
    pz = b._bench_parser
    lg = np.asarray(pz.disp_3d.lags, dtype=np.int32)
    for _ in range(2):
        raw_self_moments(pz._dc_sel, T, lg, 7, None)
    k2_isolated = float('inf')
    for _ in range(3):  # the best of three trains of ten launches (see `isolated` below)
        torch.cuda.synchronize()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        for _ in range(10):
            raw_self_moments(pz._dc_sel, T, lg, 7, None)
        i1.record()
        torch.cuda.synchronize()
        k2_isolated = min(k2_isolated, i0.elapsed_time(i1) / 10)
    # the two stage-1 kernels the same way (HBM-bound: algorithmic bytes / time against the measured HBM peak)
    from kinisi_b200._lib import MODE_FRACTIONAL
    tr = pz._traj
    pitch = engine.pitch_for(T)
    idx_d = engine.small_to_device(idx, dev, np.int32)
    fw_d = engine.small_to_device(fw, dev, np.int32)

    def isolated(fn, n=10, reps=3):
        """Mean duration of n back-to-back launches (CUDA events on the launching stream), the best of `reps` such trains:
        the clock sampler's NVML queries (every 10 ms, driver locks) land inside some of the ~1 ms trains and stall a launch."""
        for _ in range(2):
            fn()
        best = float('inf')
        for _ in range(reps):
            torch.cuda.synchronize()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            z.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(z) / n)
        return best

    fs_ms = isolated(lambda: engine.frame_sums(resident, MODE_FRACTIONAL, tr.latt, tr.latt_inv, tr.latt_stride, fw_d, None, pitch))
    fsum = engine.frame_sums(resident, MODE_FRACTIONAL, tr.latt, tr.latt_inv, tr.latt_stride, fw_d, None, pitch)
    uw_out = torch.empty((n_mob, 3, pitch), dtype=torch.float64, device=dev)
    uw_ms = isolated(lambda: engine.unwrap_cumsum(resident, MODE_FRACTIONAL, tr.latt, tr.latt_inv, tr.latt_stride, idx_d, fsum,
                                                  1.0 / n_fw, pitch, out=uw_out))
    del uw_out
    e2e_ms, _, b2 = timed(host, args.steps)
    # the other half of BASELINE.json's metric, "wall time per trajectory": one analysis at a time on one stream, each waited
    # for before the next starts (device-resident input and pinned host input)
    def latency(coords, n=8):
        ms = []
        for i in range(n + 2):
            if world > 1:
                td.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            b1 = step(coords)
            _ = b1.flatchain
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(1e3 * (time.perf_counter() - t0))
        t = torch.tensor([float(np.median(ms))], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0])

    lat_resident_ms, lat_host_ms = latency(resident), latency(host)
    clocks = sampler.stop() if sampler is not None else None
    # the same bytes as a plain pinned host->device copy on every rank at once: the ceiling of `e2e` (one host feeds all GPUs)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        resident.copy_(host, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    tcopy = torch.tensor([c0.elapsed_time(c1) / 5], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(tcopy, op=td.ReduceOp.MAX)
    plain_h2d_ms = float(tcopy[0])
    engine.self_moments = raw_self_moments
    shapes = None
    if not args.no_named_shapes:
        try:
            peaks0 = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            peaks0 = {}
        del resident
        torch.cuda.empty_cache()
        shapes = named_shapes(args, torch, td, dev, world, rank, fp64_peak_tflops, peaks0.get('hbm_gbs', 6650.0))

    value = world * terms_rank * args.steps / (ms_total * 1e-3)
    e2e_value = world * terms_rank * args.steps / (e2e_ms * 1e-3)
    h2d = host.numel() * host.element_size() + latt.nbytes + 4 * (n_mob + n_fw) + 8 * (2 * CFG['n_steps'])
    d2h = b2.flatchain.nbytes

    if rank == 0:
        k2_in_step = float(np.mean(k2_ms)) if k2_ms else float('nan')
        k2 = k2_isolated
        achieved = 16.0 * terms_rank / (k2 * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_c, n_cf = 96, 64
            fc, lc = make_inputs(CFG['seed'], n_c, n_cf, T, CFG['box'], CFG['sigma'])
            mods = reference_modules()
            if mods is not None:
                t_c, s_c, _d = reference_pass(mods, fc, lc, n_c, n_cf)
                what, kind = 'unmodified kinisi 1.1.1 (oracle/_ref + oracle/shims)', 'reference'
            else:
                t_c, s_c = cpu_pass(fc, lc, n_c, n_cf, CFG['n_steps'])
                what, kind = 'numpy oracle port', 'port'
            cpu = {'value': t_c / s_c, 'unit': UNIT, 'cores': 1, 'kind': kind,
                   'sample': f'{n_c} of {n_mob} mobile + {n_cf} framework atoms x {T} frames, {CFG["n_steps"]} dt, '
                             f'full covariance + sampler; {what}, {s_c:.1f} s'}
        line = {
            'metric': METRIC,
            'value': value,
            'unit': UNIT,
            'n_gpus': world,
            'steps': args.steps,
            'warmup': max(args.warmup, 3),
            'ms_per_step': ms_total / args.steps,
            'higher_is_better': True,
            'scaling': 'weak',
            'vs_baseline': None,
            'dtype': 'f64',
            'data': 'synthetic',
            'config': bench_config(args),
            'run': {
                'joins': ('none' if world == 1 else
                          'peer memory, no launch of their own: frame sums joined inside kb2_frame_sums_joined, moment sums inside kb2_moment_stats_joined'
                          if any(kdist._PEER.values()) and os.environ.get('KB2_FUSED_JOIN', '1') != '0' else
                          'peer-memory kernel (kb2_peer_allreduce_f64), 2 per step' if any(kdist._PEER.values()) else
                          'NCCL all-reduce, 2 per step'),
                'self_moment_variant': {0: 'direct', 1: 'tma-ring', 2: 'arithmetic-runs', 3: 'lag-window', 4: 'tiled-lag-window'}[
                    (engine.choose_variant(lags, T) if engine.DEFAULT_VARIANT == engine.VARIANT_AUTO else engine.DEFAULT_VARIANT) & 0xff]
                + (' (chosen by the lag grid)' if engine.DEFAULT_VARIANT == engine.VARIANT_AUTO else ' (forced)'),
                'wall_ms_per_step': wall_total / args.steps,
                'analyses_in_flight': args.in_flight,
            },
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': e2e_ms / args.steps, 'plain_pinned_h2d_ms_per_step': plain_h2d_ms,
                    'plain_pinned_h2d_gbs_per_gpu': host.numel() * host.element_size() / (plain_h2d_ms * 1e-3) / 1e9,
                    'fraction_of_plain_h2d': plain_h2d_ms / (e2e_ms / args.steps)},
            'latency': {'what': 'wall time of ONE analysis (stages 1-3, result read back), nothing else in flight; median of 8',
                        'device_resident_input_ms': lat_resident_ms, 'pinned_host_input_ms': lat_host_ms},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {
                'kernel': 'self_moments (stage 2)',
                'bound': 'fp64',
                'achieved': achieved,
                'peak': fp64_peak_tflops,
                'unit': 'TFLOP/s',
                'frac': achieved / fp64_peak_tflops,
                'traffic': k2_traffic_bytes(),
                'note': '8 FP64 instructions per term counted as 2 flop each; peak = DFMA micro-benchmark measured '
                        'in this run (MEASURED_PEAKS.json has no FP64 figure); kernel time = CUDA events on the '
                        'launching stream, 10 back-to-back launches on the last timed step\'s data, best of three such trains (in-step interval '
                        'reported beside it)',
                'kernel_ms': k2,
                'kernel_ms_in_step': k2_in_step,
                'hbm_gbs_algorithmic': n_mob * 3 * T * 8 / (k2 * 1e-3) / 1e9,
                'hbm_peak_gbs': peaks.get('hbm_gbs', 6650.0),
                'hbm_peak_source': 'measured' if 'hbm_gbs' in peaks else 'fallback',
            },
            'other_kernels': [
                {'kernel': 'frame_sums (stage 1, framework drift)', 'bound': 'hbm', 'ms': fs_ms,
                 'achieved_gbs': n_fw * (T + 1) * 12 / (fs_ms * 1e-3) / 1e9, 'peak_gbs': peaks.get('hbm_gbs', 6650.0),
                 'algorithmic_bytes': n_fw * (T + 1) * 12},
                {'kernel': 'unwrap_cumsum (stage 1, displacement)', 'bound': 'hbm', 'ms': uw_ms,
                 'achieved_gbs': (n_mob * (T + 1) * 12 + n_mob * 3 * pitch * 8) / (uw_ms * 1e-3) / 1e9,
                 'peak_gbs': peaks.get('hbm_gbs', 6650.0),
                 'algorithmic_bytes': n_mob * (T + 1) * 12 + n_mob * 3 * pitch * 8},
                {'kernel': 'jacobi_project + ensemble_sample (stage 3)', 'bound': 'latency (one CTA / one warp)',
                 'ms': float(stages[2])},
            ],
            'stage_ms': {'stage1_parser': float(stages[0]), 'stage2_moments': float(stages[1]),
                         'stage3_cov_mcmc': float(stages[2])},
            'cpu_baseline': cpu,
            'named_shapes': shapes,
            'result': {'D_median_cm2_s': b.D.n, 'D_ci95': [float(x) for x in b.D.con_int()],
                       'D_true': 3 * CFG['sigma']**2 / CFG['time_step'] / (2e4 * 3)},
        }
        print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--variant', type=int, default=None, help='self-moment kernel: 0 direct, 1 TMA ring, 2 arithmetic runs, 3 lag window, 4 tiled lag window (default)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-named-shapes', action='store_true', help='skip the sub-records for configs[0], [2], [3], [4]')
    ap.add_argument('--in-flight', type=int, default=4,
                    help='analyses in flight: consecutive trajectories go round-robin to this many CUDA streams (1 = one stream)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
