"""BASELINE configs[4] per-GPU shard (12,500 atoms x 50,000 frames, f32) and configs[2] (10,000 ions x 100,000 frames,
MSTD): size-independent checks (MSD slope, NGP -> 0, D) and timings on one GPU."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import diffusion, engine
from kinisi_b200.parser import Parser
dev = torch.device('cuda')
which = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
n_atoms, T, kind = (12500, 50000, 'msd') if which == 'cfg4' else (10000, 100000, 'mstd')
sigma, box = 0.1, 60.0
gen = torch.Generator(device=dev); gen.manual_seed(11)
coords = torch.empty((n_atoms, T + 1, 3), dtype=torch.float32, device=dev)
for a0 in range(0, n_atoms, 500):  # wrapped random walk, built in blocks
    a1 = min(n_atoms, a0 + 500)
    st = torch.randn((a1 - a0, T + 1, 3), dtype=torch.float64, device=dev, generator=gen) * (sigma / box)
    st[:, 0] = torch.rand((a1 - a0, 3), dtype=torch.float64, device=dev, generator=gen)
    st[:, 1] = 0.0
    coords[a0:a1] = (torch.cumsum(st, dim=1) % 1.0).to(torch.float32)
    del st
latt = np.eye(3)[None] * box
torch.cuda.synchronize()
def run():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    p = Parser(Parser.get_disp(coords, latt), np.arange(n_atoms), [], 0.001, 1, n_steps=100, memory_limit=64., progress=False)
    ev[1].record()
    cls = diffusion.MSDBootstrap if kind == 'msd' else diffusion.MSTDBootstrap
    b = cls(p.delta_t, p.disp_3d, p._n_o, progress=False)
    ev[2].record()
    if kind == 'msd':
        b.diffusion(float(b.dt[10]), progress=False)
    else:
        b.jump_diffusion(float(b.dt[5]), progress=False)
    ev[3].record()
    torch.cuda.synchronize()
    return p, b, [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
p, b, ms = run()
p = b = None  # release the first run's arrays before the timed one: the allocator reuses them
p, b, ms = run()
lags = np.asarray(p.disp_3d.lags)
terms = n_atoms * float(np.sum(T - lags + 1))
print(which, 'stage ms', [round(x, 2) for x in ms], 'terms/s (stage 2)', '%.3e' % (terms / ms[1] * 1e3),
      'fp64 frac', round(16 * terms / (ms[1] * 1e-3) / 35.4e12, 3))
# physics: 3-D Gaussian walk, sigma per component and frame
if kind == 'msd':
    slope = np.polyfit(b.dt[:20], b.n[:20], 1)[0]
    print('MSD slope', slope, 'expected', 3 * sigma**2 / 0.001, 'ngp[:5]', b.ngp[:5], 'D', b.D.n, 'expected', 3 * sigma**2 / 0.001 / 6e4)
    assert abs(slope / (3 * sigma**2 / 0.001) - 1) < 0.01 and np.all(np.abs(b.ngp[:5]) < 0.01)
else:
    # one collective trajectory: the statistical error at lag k is ~sqrt(k / T), so compare with a direct evaluation
    S = p.disp_3d.dc[:, :, :T].sum(dim=0)  # [3, T]
    for i in (0, 3, 17, 40):
        k = int(lags[i])
        d = torch.cat([S[:, k - 1:k], S[:, k:] - S[:, :T - k]], dim=1)
        ref = float((d * d).sum(dim=0).mean())
        print('lag', k, 'mstd', b.n[i], 'direct', ref)
        assert abs(b.n[i] / ref - 1) < 1e-9
    print('MSTD / (N dt) at the first lags', (b.n[:4] / b.dt[:4] / n_atoms), 'expected', 3 * sigma**2 / 0.001, 'D_J', b.D_J.n)
print('ok')
