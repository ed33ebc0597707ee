"""The four moment-kernel variants over a family of lag grids np.linspace(1, T, n_steps, dtype=int)[:n_used] (about 300 MB of
cumulative displacements each): one JSON line per grid with the time of every variant (5 launches back to back, CUDA
events) and the largest relative difference from the direct kernel.  python tools/k2grid_survey.py > out.jsonl"""
import json
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
dev = torch.device('cuda')
GRIDS = [(T, n, f) for T in (1000, 2000, 5000, 10000, 20000, 50000, 100000) for n in (50, 100, 200, 400, 1000, 2000)
         for f in (1.0, 0.5) if n <= T and not (f == 0.5 and n in (50, 400))]
if len(sys.argv) > 1:
    GRIDS = GRIDS[int(sys.argv[1])::int(sys.argv[2])]
for T, n_steps, frac in GRIDS:
    lags = np.unique(np.linspace(1, T, n_steps, dtype=int))
    lags = lags[:max(2, int(len(lags) * frac))].astype(np.int32)
    # enough atoms to fill the device, less for the grids with many lags (time ~ atoms x T x lags)
    n_atoms = int(max(64, min(3e8 / (24 * T), 6e10 / (T * len(lags)))))
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    pitch = engine.pitch_for(T)
    dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
    terms = n_atoms * float(np.sum(T - lags + 1))
    rec = {'T': T, 'n_steps': n_steps, 'n_used': len(lags), 'n_atoms': n_atoms, 'terms': terms, 'ms': {}, 'err': {}}
    ref = None
    for v in (0, 2, 3, 4):
        for _ in range(2):
            out = engine.self_moments(dc, T, lags, 7, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = engine.self_moments(dc, T, lags, 7, v)
        e1.record(); torch.cuda.synchronize()
        rec['ms'][v] = e0.elapsed_time(e1) / 3
        if ref is None:
            ref = out.clone()
        rec['err'][v] = float(((out - ref).abs() / ref.abs()).max())
    rec['auto'] = int(engine.choose_variant(lags, T)) if hasattr(engine, 'choose_variant') else None
    print(json.dumps(rec), flush=True)
    del dc
