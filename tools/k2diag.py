"""Every moment-kernel variant, launched five times back to back, against variant 3 launched once (dense grids of 1000 atoms);
with a few atoms against the numpy oracle."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
from oracle import kinisi_oracle as ko
dev = torch.device('cuda')
for n_atoms, T, n_steps, n_used in ((6, 20000, 1000, 500), (1000, 20000, 1000, 500), (1000, 20000, 400, 200)):
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    pitch = engine.pitch_for(T)
    dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
    lags = np.linspace(1, T, n_steps, dtype=int)[:n_used].astype(np.int32)
    if n_atoms <= 8:
        host = dc[:, :, :T].permute(0, 2, 1).contiguous().cpu().numpy()
        S1, S2, _, _ = ko.raw_sums_streaming(host, lags)
        ref, name = np.stack([S1, S2], axis=1), 'oracle'
    else:
        ref, name = engine.self_moments(dc, T, lags, 7, 3).cpu().numpy(), 'variant 3 once'
    for v in (0, 2, 3, 4):
        outs = [engine.self_moments(dc, T, lags, 7, v) for _ in range(5)]
        torch.cuda.synchronize()
        for j, o in enumerate(outs):
            rel = np.abs(o.cpu().numpy() - ref) / np.abs(ref)
            if rel.max() > 1e-9 or j == 4:
                i = np.unravel_index(np.argmax(rel), rel.shape)
                print(n_atoms, T, n_steps, n_used, 'variant', v, 'launch', j, 'vs', name, 'max rel', rel.max(), 'lag', lags[i[0]], 'n bad', int((rel > 1e-9).sum()))
