"""bootstrap=True at BASELINE configs[1]: time of the resampled distributions (all lags) on the device, the gather rate of
the resampling kernel, and the reference's `_bootstrap` (oracle restatement, numpy) on a few resamples of the same lag."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from kinisi_b200 import diffusion, engine
from kinisi_b200.parser import Parser
from oracle import kinisi_oracle as ko

C = bench.CFG
dev = torch.device('cuda:0')
frac, latt = bench.make_inputs(C['seed'], C['n_mobile'], C['n_framework'], C['n_frames'], C['box'], C['sigma'])
idx = np.arange(C['n_mobile']); fw = np.arange(C['n_mobile'], C['n_mobile'] + C['n_framework'])
traj = Parser.get_disp(torch.from_numpy(frac).to(dev), latt, progress=False, device=dev)
p = Parser(traj, idx, fw, C['time_step'], C['step_skip'], n_steps=C['n_steps'], memory_limit=64., progress=False)
out = {}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, bootstrap=True, random_state=np.random.RandomState(0), progress=False)
    torch.cuda.synchronize(); out['msd_bootstrap_true_s'] = time.perf_counter() - t0
out['sizes'] = [int(d.size) for d in b._distributions[:5]]
out['rel_dev_n'] = float(np.max(np.abs(b._n_bootstrap / b.n - 1)))
out['v_ratio'] = [float(x) for x in (b._v_bootstrap / b.v)[:5]]
# the kernel alone on lag 1 (population 448 x 20000 = 72 MB, n_draws = n_o[0])
vals = engine.lag_squares(p.disp_3d.dc, p.disp_3d.n_frames, 1, 7)
nd = int(p._n_o[0])
for _ in range(2):
    engine.resample_means(vals, nd, 1000, 0, 1)
torch.cuda.synchronize()
a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); engine.resample_means(vals, nd, 1000, 0, 1); z.record(); torch.cuda.synchronize()
ms = a.elapsed_time(z)
out['lag1'] = {'population': int(vals.numel()), 'n_draws': nd, 'resamples': 1000, 'ms': ms, 'gathers_per_s': nd * 1000 / ms * 1e3}
vals2 = engine.lag_squares(p.disp_3d.dc, p.disp_3d.n_frames, int(p.disp_3d.lags[5]), 7)
nd2 = int(p._n_o[5])
engine.resample_means(vals2, nd2, 1000, 0, 1); torch.cuda.synchronize()
a.record(); engine.resample_means(vals2, nd2, 1000, 0, 1); z.record(); torch.cuda.synchronize()
out['lag5'] = {'population': int(vals2.numel()), 'n_draws': nd2, 'ms': a.elapsed_time(z), 'gathers_per_s': nd2 * 1000 / a.elapsed_time(z) * 1e3}
h = vals.cpu().numpy()
t0 = time.perf_counter(); ko.bootstrap_means(h, nd, 3, np.random.RandomState(0)); t = time.perf_counter() - t0
out['cpu_reference_lag1'] = {'resamples': 3, 's': t, 'gathers_per_s': 3 * nd / t, 'extrapolated_1000_s': t / 3 * 1000}
print(json.dumps(out, indent=1))
