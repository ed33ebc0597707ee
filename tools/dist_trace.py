"""Kernel timeline of the bench loop on rank 0 (torch profiler / CUPTI), run under torchrun or alone: where the device
idles between the kernels of the analyses in flight.  Writes gpurun_out/dist_trace.txt."""
import json, os, sys, time
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
C = bench.CFG
frac, latt = bench.make_inputs(C['seed'] + rank, C['n_mobile'], C['n_framework'], C['n_frames'], C['box'], C['sigma'])
lags = bench.lag_grid(C['n_frames'], C['n_steps'])
start_dt = float(lags[C['start_idx']] * C['time_step'] * C['step_skip'])
idx = np.arange(C['n_mobile']); fw = np.arange(C['n_mobile'], C['n_mobile'] + C['n_framework'])
resident = torch.from_numpy(frac).to(dev)
pkw = dict(n_steps=C['n_steps'], memory_limit=64., progress=False, distributed=world > 1,
           global_counts=(world * C['n_mobile'], world * C['n_framework']) if world > 1 else None)
L = int(os.environ.get('LANES', '4'))
lanes = [torch.cuda.Stream(dev) for _ in range(L)]

def step():
    p = Parser(Parser.get_disp(resident, latt, progress=False, device=dev), idx, fw, C['time_step'], C['step_skip'], **pkw)
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    b.diffusion(start_dt, progress=False)
    return b

def loop(n):
    held = [None] * L
    for i in range(n):
        with torch.cuda.stream(lanes[i % L]):
            held[i % L] = step()
    torch.cuda.synchronize()

for _ in range(3): step()
loop(3 * L)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter(); loop(16); wall = time.perf_counter() - t0
if rank == 0:
    path = 'gpurun_out/dist_trace.json'
    os.makedirs('gpurun_out', exist_ok=True)
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memcpy', 'gpu_memset')]
    ev.sort(key=lambda e: e['ts'])
    t_first, t_last = ev[0]['ts'], max(e['ts'] + e['dur'] for e in ev)
    out = [f'world {world} lanes {L}: wall {wall / 16 * 1e3:.3f} ms/step, device span {(t_last - t_first) / 16 / 1e3:.3f} ms/step, {len(ev)} device events']
    # busy time = union of intervals
    busy, cur_s, cur_e = 0.0, None, None
    gaps = []
    for e in ev:
        s, f = e['ts'], e['ts'] + e['dur']
        if cur_e is None: cur_s, cur_e = s, f
        elif s <= cur_e: cur_e = max(cur_e, f)
        else:
            busy += cur_e - cur_s; gaps.append((s - cur_e, cur_e, s)); cur_s, cur_e = s, f
    busy += cur_e - cur_s
    out.append(f'device busy (union) {busy / 16 / 1e3:.3f} ms/step, idle {(t_last - t_first - busy) / 16 / 1e3:.3f} ms/step in {len(gaps)} gaps')
    tot = {}
    for e in ev:
        k = e['name'][:60]; tot[k] = tot.get(k, 0.0) + e['dur']
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:12]:
        out.append(f'  {v / 16:9.1f} us/step  {k}')
    out.append('largest gaps (us): what ended before / what started after')
    for g, a, b in sorted(gaps, reverse=True)[:14]:
        before = max((e for e in ev if e['ts'] + e['dur'] <= a + 0.5), key=lambda e: e['ts'] + e['dur'])
        after = min((e for e in ev if e['ts'] >= b - 0.5), key=lambda e: e['ts'])
        out.append(f'  {g:7.1f}  {before["name"][:44]:44s} -> {after["name"][:44]}')
    out.append('timeline (us from the start of the 5th step; stream, start, duration, name), kernels longer than 15 us')
    big = [e for e in ev if e['dur'] > 15]
    k2 = [e for e in big if 'self_moments' in e['name']]
    t_ref = k2[4]['ts'] if len(k2) > 4 else t_first
    for e in big:
        if t_ref - 200 <= e['ts'] <= t_ref + 2600:
            out.append(f"  s{e['args'].get('stream', '?'):>3}  {e['ts'] - t_ref:8.1f}  {e['dur']:7.1f}  {e['name'][:50]}")
    open(f'gpurun_out/dist_trace_w{world}.txt', 'w').write('\n'.join(out) + '\n')
    print('\n'.join(out))
    os.remove(path)
if world > 1:
    td.destroy_process_group()
