"""A host-planned moment kernel (KB2_PROBE_VARIANT: 3 lag-window, 4 tiled lag-window (default)) against a torch float64
evaluation on a set of lag grids, then timings of variants 2, 3 and 4 on the BASELINE shapes.
Run on a B200: python tools/k2win_probe.py [quick]"""
import json
import os
import sys

sys.path.insert(0, '/root/repo')
import numpy as np
import torch

from kinisi_b200 import engine

dev = torch.device('cuda')
VAR = int(os.environ.get('KB2_PROBE_VARIANT', '4'))
gen = torch.Generator(device=dev)
gen.manual_seed(1)


def walk(n_atoms, T):
    pitch = engine.pitch_for(T)
    dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
    return dc


def torch_sums(dc, T, lags, comps=(0, 1, 2)):
    out = np.zeros((len(lags), 2))
    x = dc[:, list(comps), :T]
    for i, k in enumerate(lags):
        k = int(k)
        d2 = (x[:, :, k - 1]**2).sum(1)
        s1, s2 = d2.sum(), (d2 * d2).sum()
        if k < T:
            dd = ((x[:, :, k:] - x[:, :, :T - k])**2).sum(1)
            s1 = s1 + dd.sum()
            s2 = s2 + (dd * dd).sum()
        out[i] = (float(s1), float(s2))
    return out


fails = 0
cases = []
for T in (140, 1000, 5431, 20000):
    cases.append(('lin100_T%d' % T, T, np.linspace(1, T, 100, dtype=int)))
cases.append(('lin1000_T5431', 5431, np.linspace(1, 5431, 1000, dtype=int)))
cases.append(('log_T20000', 20000, np.unique(np.geomspace(1, 20000, 100, dtype=int))))
cases.append(('all_T300', 300, np.arange(1, 301)))
cases.append(('dups_T50', 50, np.linspace(1, 50, 100, dtype=int)))
cases.append(('min5_T4000', 4000, np.linspace(5, 3000, 37, dtype=int)))
cases.append(('single', 777, np.array([13])))
cases.append(('odd', 2048, np.array([1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2048])))
for name, T, lags in cases:
    for n_atoms in (1, 7, 70):
        dc = walk(n_atoms, T)
        ref = torch_sums(dc, T, lags)
        for mask, comps in ((7, (0, 1, 2)), (5, (0, 2)), (2, (1, ))):
            refm = ref if mask == 7 else torch_sums(dc, T, lags, comps)
            out = engine.self_moments(dc, T, lags.astype(np.int32), mask, VAR).cpu().numpy()
            err = np.max(np.abs(out - refm) / np.maximum(np.abs(refm), 1e-300))
            ok = err < 1e-11
            fails += not ok
            if not ok or mask == 7:
                print(f'{name:16s} atoms {n_atoms:3d} mask {mask} K {len(lags):4d} relerr {err:.2e} {"ok" if ok else "FAIL"}', flush=True)
print('FAILS', fails, flush=True)

res = {}
if len(sys.argv) > 1 and sys.argv[1] == 'quick':
    shapes = [(448, 20000)]
else:
    shapes = [(448, 20000), (2000, 50000), (1000, 100000), (3000, 5000)]
for n_atoms, T in shapes:
    dc = walk(n_atoms, T)
    grids = {'lin': np.linspace(1, T, 100, dtype=int)}
    if T == 20000:
        grids['log'] = np.unique(np.geomspace(1, T, 100, dtype=int))
    for gname, lags in grids.items():
        terms = n_atoms * float(np.sum(T - lags + 1))
        ref = None
        for name, v in (('runs', 2), ('win', 3), ('tile', 4)):
            for _ in range(3):
                out = engine.self_moments(dc, T, lags.astype(np.int32), 7, v)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                out = engine.self_moments(dc, T, lags.astype(np.int32), 7, v)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            if ref is None:
                ref = out.clone()
            err = float(((out - ref).abs() / ref.abs()).max())
            key = f'{name}_{gname}_{n_atoms}x{T}'
            res[key] = dict(ms=ms, terms_per_s=terms / ms * 1e3, fp64_tflops=16 * terms / ms * 1e3 / 1e12, relerr_vs_runs=err)
            print(key, res[key], flush=True)
tag = os.environ.get('KB2_PROBE_TAG', 'def')
json.dump(res, open(f'/root/repo/gpurun_out/k2win_probe_{tag}.json', 'w'), indent=1)
