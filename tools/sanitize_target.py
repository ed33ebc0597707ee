"""Small invocations of the kernels with cross-thread / cross-CTA protocols, the target of the compute-sanitizer runs
(tools/sanitize.sh): unwrap_cumsum over several tiles per atom (totals exchanged between CTAs through self-flagging
slots), the three host-planned moment kernels (tile: mbarrier producer / consumer; win, runs: per-warp shared accumulators)
on plain and drifting grids, the Jacobi spectral kernel, the one-warp ensemble sampler and the resampling kernel."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from kinisi_b200 import engine, diffusion
from kinisi_b200.parser import Parser

dev = torch.device('cuda')
rng = np.random.default_rng(3)
n_mob, n_fw, T = 6, 4, 3100  # three tiles of 1280 frames per atom
frac = (rng.uniform(0, 1, (n_mob + n_fw, 1, 3)) + np.cumsum(rng.normal(0, 0.01, (n_mob + n_fw, T + 1, 3)), axis=1)) % 1.0
frac[:, 0] = frac[:, 1]
latt = (np.eye(3) * 11.0)[None]
for dtype in (np.float32, np.float64):
    p = Parser(Parser.get_disp(torch.from_numpy(frac.astype(dtype)).to(dev), latt), np.arange(n_mob), np.arange(n_mob, n_mob + n_fw),
               0.001, 1, n_steps=40, progress=False)
    dc = p.disp_3d.dc
    for grid in (np.linspace(1, T, 40, dtype=int), np.linspace(1, T, int(T / 12.09), dtype=int), np.arange(1, 80),
                 np.unique(np.geomspace(1, T, 30, dtype=int))):
        ref = engine.self_moments(dc, T, grid, 7, engine.VARIANT_DIRECT)
        for v in (engine.VARIANT_TILE, engine.VARIANT_WIN, engine.VARIANT_RUNS, engine.VARIANT_TMA):
            out = engine.self_moments(dc, T, grid, 7, v)
            assert torch.allclose(out, ref, rtol=1e-11), (v, len(grid))
    b = diffusion.MSTDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b.diffusion(float(b.dt[3]), progress=False, n_samples=60, n_burn=40)   # cov_build, jacobi_project, ensemble_sample_warp
    assert np.isfinite(b.D.n)
b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, bootstrap=True, n_resamples=50, max_resamples=100, progress=False,
                           random_state=np.random.RandomState(1))
torch.cuda.synchronize()
print('sanitize target ok')
