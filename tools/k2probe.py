import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
dev = torch.device('cuda')
n_atoms, T = 448, 20000
if len(sys.argv) > 2:
    n_atoms, T = int(sys.argv[1]), int(sys.argv[2])
gen = torch.Generator(device=dev); gen.manual_seed(1)
pitch = engine.pitch_for(T)
dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
lags = torch.as_tensor(np.linspace(1, T, 100, dtype=int).astype(np.int32)).to(dev)
terms = n_atoms * float(np.sum(T - np.linspace(1, T, 100, dtype=int) + 1))
res = {}
ref = None
print('plan', engine.self_moments_plan(lags.cpu().numpy(), T))
import os
VARIANTS = (('direct_r8', 0), ('tma_2cta', 1 | (2 << 8)), ('runs', 2))
if os.environ.get('K2_ONLY'):
    VARIANTS = tuple(x for x in VARIANTS if x[0] == os.environ['K2_ONLY'])
for name, v in VARIANTS:
    for _ in range(3):
        out = engine.self_moments(dc, T, lags, 7, v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = engine.self_moments(dc, T, lags, 7, v)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # single launch, GPU idle before
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = engine.self_moments(dc, T, lags, 7, v); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    if ref is None: ref = out.clone()
    err = float(((out - ref).abs() / ref.abs()).max())
    res[name] = dict(ms_loop=ms, host_launch_ms=(t1 - t0) * 1e3, wall_single_ms=(t2 - t0) * 1e3, terms_per_s=terms / ms * 1e3, relerr=err,
                     fp64_tflops=16 * terms / ms * 1e3 / 1e12)
    print(name, res[name], flush=True)
json.dump(res, open('/root/repo/gpurun_out/k2probe.json', 'w'), indent=1)
