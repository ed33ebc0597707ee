"""Host-side issue rate of the analysis loop: a short trajectory (little device work) run alone or under torchrun with the
atoms sharded, so that the wall time per step is the host's (Python, launches, collectives' enqueue cost)."""
import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as td
import bench
from kinisi_b200 import diffusion
from kinisi_b200.parser import Parser
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
dev = torch.device('cuda', torch.cuda.current_device())
if world > 1:
    td.init_process_group('nccl')
C = dict(bench.CFG); T = int(os.environ.get('T', '2000'))
frac, latt = bench.make_inputs(C['seed'] + rank, C['n_mobile'], C['n_framework'], T, C['box'], C['sigma'])
lags = bench.lag_grid(T, C['n_steps'])
start_dt = float(lags[C['start_idx']] * C['time_step'] * C['step_skip'])
idx = np.arange(C['n_mobile']); fw = np.arange(C['n_mobile'], C['n_mobile'] + C['n_framework'])
resident = torch.from_numpy(frac).to(dev)
pkw = dict(n_steps=C['n_steps'], memory_limit=64., progress=False, distributed=world > 1,
           global_counts=(world * C['n_mobile'], world * C['n_framework']) if world > 1 else None)
lanes = [torch.cuda.Stream(dev) for _ in range(4)]
def step():
    p = Parser(Parser.get_disp(resident, latt, progress=False, device=dev), idx, fw, C['time_step'], C['step_skip'], **pkw)
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b.diffusion(start_dt, progress=False, n_samples=100, n_burn=50)
    return b
def loop(n):
    held = [None] * 4
    for i in range(n):
        with torch.cuda.stream(lanes[i % 4]):
            held[i % 4] = step()
    torch.cuda.synchronize()
loop(16)
t0 = time.perf_counter(); loop(200); wall = (time.perf_counter() - t0) / 200
if rank == 0:
    print(f'world {world} T {T}: wall {wall * 1e3:.3f} ms per step')
    pr = cProfile.Profile(); pr.enable(); loop(100); pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(60)
else:
    loop(100)
if world > 1:
    td.destroy_process_group()
