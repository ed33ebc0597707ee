"""Why unwrap_cumsum times 8-15 % slower inside bench.py than in tools/k1probe.py: the same call timed (a) in a fresh process,
(b) on bench.py's own input, (c) after the analysis loop has run (allocator state, side streams, pinned buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from kinisi_b200 import diffusion, engine
from kinisi_b200.parser import Parser
from kinisi_b200._lib import MODE_FRACTIONAL
dev = torch.device('cuda')
C = bench.CFG
n_mob, n_fw, T = C['n_mobile'], C['n_framework'], C['n_frames']
pitch = engine.pitch_for(T)
latt_t = (torch.eye(3, dtype=torch.float64, device=dev) * C['box']).reshape(1, 3, 3)
inv = engine.invert_lattice(latt_t)
idx = torch.arange(n_mob, dtype=torch.int32, device=dev)
fw = torch.arange(n_mob, n_mob + n_fw, dtype=torch.int32, device=dev)

def timeit(fn, n=10, reps=3):
    for _ in range(3): fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3

def uw(coords, tag):
    sums = engine.frame_sums(coords, MODE_FRACTIONAL, latt_t, inv, 0, fw, None, pitch)
    out = torch.empty((n_mob, 3, pitch), dtype=torch.float64, device=dev)
    t = timeit(lambda: engine.unwrap_cumsum(coords, MODE_FRACTIONAL, latt_t, inv, 0, idx, sums, 1.0 / n_fw, pitch, out=out))
    print(f'{tag:60s} {t:6.1f} us')

gen = torch.Generator(device=dev); gen.manual_seed(1)
steps = torch.randn((n_mob + n_fw, T + 1, 3), dtype=torch.float32, device=dev, generator=gen) * 0.004
synth = (torch.cumsum(steps, dim=1) % 1.0).contiguous()
del steps
uw(synth, '(a) fresh process, k1probe input')
frac, latt = bench.make_inputs(C['seed'], n_mob, n_fw, T, C['box'], C['sigma'])
host = torch.from_numpy(frac).pin_memory()
resident = host.to(dev)
uw(resident, '(b) bench input (make_inputs), before any analysis')
uw(synth, '    k1probe input again')
lags = bench.lag_grid(T, C['n_steps'])
start_dt = float(lags[C['start_idx']] * C['time_step'] * C['step_skip'])
ia, fa = np.arange(n_mob), np.arange(n_mob, n_mob + n_fw)
lanes = [torch.cuda.Stream(dev) for _ in range(4)]
def step(coords):
    p = Parser(Parser.get_disp(coords, latt, progress=False, device=dev), ia, fa, C['time_step'], C['step_skip'], n_steps=C['n_steps'], memory_limit=64., progress=False)
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b.diffusion(start_dt, progress=False)
    return b
held = [None] * 4
for i in range(24):
    with torch.cuda.stream(lanes[i % 4]):
        held[i % 4] = step(resident)
torch.cuda.synchronize()
uw(resident, '(c) bench input after 24 resident analyses on four lanes')
for i in range(8):
    with torch.cuda.stream(lanes[i % 4]):
        held[i % 4] = step(host)
torch.cuda.synchronize()
uw(resident, '(d) ... and 8 analyses from pinned host memory')
uw(synth, '    k1probe input again')
print('allocated %.2f GB reserved %.2f GB' % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
