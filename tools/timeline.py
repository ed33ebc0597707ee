"""Kernel timeline of one bench step (torch.profiler / CUPTI): start offsets, durations and the gaps between kernels."""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench
from kinisi_b200 import diffusion, engine
from kinisi_b200.parser import Parser
mode = sys.argv[1] if len(sys.argv) > 1 else 'resident'
C = bench.CFG
dev = torch.device('cuda')
frac, latt = bench.make_inputs(C['seed'], C['n_mobile'], C['n_framework'], C['n_frames'], C['box'], C['sigma'])
lags = bench.lag_grid(C['n_frames'], C['n_steps'])
start_dt = float(lags[C['start_idx']] * C['time_step'] * C['step_skip'])
idx = np.arange(C['n_mobile']); fw = np.arange(C['n_mobile'], C['n_mobile'] + C['n_framework'])
host = torch.from_numpy(frac).pin_memory()
resident = host.to(dev)
def step(coords):
    traj = Parser.get_disp(coords, latt, progress=False, device=dev)
    p = Parser(traj, idx, fw, C['time_step'], C['step_skip'], n_steps=C['n_steps'], memory_limit=64., progress=False)
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b.diffusion(start_dt, progress=False)
    return b
src = resident if mode == 'resident' else host
b = None
for _ in range(4): b = step(src)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4): b = step(src)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into steps at the invert_lattice kernel
starts = [i for i, e in enumerate(evs) if 'invert_lattice' in e.name]
a, b = starts[1], starts[2]
t0 = evs[a].time_range.start
prev_end = t0
print('%-58s %9s %9s %8s' % ('kernel / copy', 'start us', 'dur us', 'gap us'))
for e in evs[a:b]:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print('%-58s %9.1f %9.1f %8.1f' % (e.name[:58], s, d, e.time_range.start - prev_end))
    prev_end = max(prev_end, e.time_range.end)
print('step span us', evs[b].time_range.start - t0)
# the same loop timed with CUDA events, without the profiler
import time
for keep in (True, False):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20):
        if keep: b = step(src)
        else: step(src)
    e1.record(); t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    print('keep result alive' if keep else 'drop result at once', 'ms/step', e0.elapsed_time(e1) / 20, 'host enqueue ms/step', 1e3 * t_enq / 20)
# pure host enqueue time: keep every result (no waits at all), 20 steps
torch.cuda.synchronize()
keepall = []
t0 = time.perf_counter()
for _ in range(20):
    keepall.append(step(src))
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
print('pure host enqueue ms/step', 1e3 * t_enq / 20)
