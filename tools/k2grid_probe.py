"""The moment-kernel variants on a given lag grid: python tools/k2grid_probe.py <n_atoms> <T> <n_steps> <n_used>
(lags = np.linspace(1, T, n_steps, dtype=int)[:n_used]); 5 launches back to back, CUDA events; all variants against variant 0."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
n_atoms, T, n_steps, n_used = (int(a) for a in sys.argv[1:5])
dev = torch.device('cuda')
gen = torch.Generator(device=dev); gen.manual_seed(1)
pitch = engine.pitch_for(T)
dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
lags = np.linspace(1, T, n_steps, dtype=int)[:n_used].astype(np.int32)
terms = n_atoms * float(np.sum(T - lags + 1))
ref = None
for v, name in ((0, 'direct'), (3, 'register windows'), (4, 'tiled windows'), (2, 'arithmetic runs')):
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
    err = float(((out - ref).abs() / ref.abs()).max())
    print(f'variant {v} ({name:16s}) {ms:8.3f} ms  {terms / ms / 1e6:7.1f} G terms/s  {16 * terms / ms / 1e9 / 35.5:5.2f} of FP64 peak  max rel diff vs direct {err:.1e}')
