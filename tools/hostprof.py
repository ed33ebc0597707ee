"""cProfile of the host side of the bench step (resident input): where the Python time per step goes."""
import sys, cProfile, pstats
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from kinisi_b200 import diffusion
from kinisi_b200.parser import Parser
C = bench.CFG
dev = torch.device('cuda')
frac, latt = bench.make_inputs(C['seed'], C['n_mobile'], C['n_framework'], C['n_frames'], C['box'], C['sigma'])
lags = bench.lag_grid(C['n_frames'], C['n_steps'])
start_dt = float(lags[C['start_idx']] * C['time_step'] * C['step_skip'])
idx = np.arange(C['n_mobile']); fw = np.arange(C['n_mobile'], C['n_mobile'] + C['n_framework'])
resident = torch.from_numpy(frac).to(dev)
def step(coords):
    traj = Parser.get_disp(coords, latt, progress=False, device=dev)
    p = Parser(traj, idx, fw, C['time_step'], C['step_skip'], n_steps=C['n_steps'], memory_limit=64., progress=False)
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b.diffusion(start_dt, progress=False)
    return b
b = None
for _ in range(5): b = step(resident)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(100): b = step(resident)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats('cumulative').print_stats(45)
