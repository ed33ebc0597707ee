"""One moment-kernel variant at one shape, a few launches (the target of an ncu capture).
usage: python tools/k2one.py <variant> <n_atoms> <T> [n_launches] [n_used lags of the 100-point grid]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
v, n_atoms, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device('cuda')
gen = torch.Generator(device=dev); gen.manual_seed(1)
pitch = engine.pitch_for(T)
dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
lags = np.linspace(1, T, 100, dtype=int).astype(np.int32)[:int(sys.argv[5]) if len(sys.argv) > 5 else 100]
for _ in range(n):
    out = engine.self_moments(dc, T, lags, 7, v)
torch.cuda.synchronize()
print(out[:3].cpu().numpy())
