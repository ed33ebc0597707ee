"""The moment-kernel variants on kinisi's logarithmic grid (np.geomspace(min_dt, max_dt, n_steps).astype(int), unique;
kinisi/parser.py:165): python tools/k2log_probe.py <n_atoms> <T> <n_steps>"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
n_atoms, T, n_steps = (int(a) for a in sys.argv[1:4])
dev = torch.device('cuda')
gen = torch.Generator(device=dev); gen.manual_seed(1)
pitch = engine.pitch_for(T)
dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
lags = np.unique(np.geomspace(1, T, n_steps).astype(int)).astype(np.int32)
terms = n_atoms * float(np.sum(T - lags + 1))
print('lags', len(lags), 'chosen', engine.choose_variant(lags, T, True))
ref = None
for v in (0, 3, 4, 2):
    for _ in range(2):
        out = engine.self_moments(dc, T, lags, 7, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = engine.self_moments(dc, T, lags, 7, v)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None:
        ref = out.clone()
    print(f'variant {v} {ms:8.3f} ms  {16 * terms / ms / 1e9 / 35.5:5.2f} of FP64 peak  max rel diff vs direct {float(((out - ref).abs() / ref.abs()).max()):.1e}')
