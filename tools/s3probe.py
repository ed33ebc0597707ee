"""Times the stage-3 kernels (Jacobi spectral step, sampler) on a configs[1]-shaped regression problem."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine, diffusion
from kinisi_b200.parser import Parser
dev = torch.device('cuda')
n_atoms, T = 64, 20000
rng = np.random.default_rng(3)
dc = np.cumsum(rng.normal(0, 0.1, size=(n_atoms, T, 3)), axis=1)
p = Parser(dc, list(range(n_atoms)), [], 0.001, 1, n_steps=100, progress=False, memory_limit=64.)
b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
start = int(sys.argv[1]) if len(sys.argv) > 1 else 10
b.diffusion(b.dt[start], progress=False)
print('n =', len(b.dt) - start, 'sweeps', int(b._sweeps.item()), 'D', b.D.n)
x = b._dev(b.dt[start:]); y = b._stats_dev[0, start:].contiguous()
cov0 = engine.cov_build(b._stats_dev[1, start:].contiguous(), b._dev(b._n_o[start:]))
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_j = timeit(lambda: engine.spectral_prepare(cov0, x, y, 1e16, True))
ev, proj, gls, ld, info = engine.spectral_prepare(cov0, x, y, 1e16, True)
t_s = timeit(lambda: engine.ensemble_sample(gls, 2, 32, 1500, ld, 12345))
print('jacobi ms %.3f  sampler ms %.3f' % (t_j, t_s))
lam = ev.cpu().numpy(); print('kept', int((proj[3] != 0).sum().item()), 'neg', int((lam < 0).sum()), 'lmax %.3e lmin %.3e' % (lam.max(), lam.min()))
