"""Times the stage-1 kernels (frame sums of the framework, unwrap + cumulative sum of the mobile atoms) on configs[1]."""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
from kinisi_b200._lib import MODE_FRACTIONAL
dev = torch.device('cuda')
n_mob, n_fw, T = 448, 1088, 20000
if len(sys.argv) > 3:
    n_mob, n_fw, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
gen = torch.Generator(device=dev); gen.manual_seed(1)
tdt = torch.float64 if os.environ.get('K1_DTYPE') == 'f64' else torch.float32
steps = torch.randn((n_mob + n_fw, T + 1, 3), dtype=tdt, device=dev, generator=gen) * 0.004
coords = (torch.cumsum(steps, dim=1) % 1.0).contiguous()
esz = coords.element_size()
latt = (torch.eye(3, dtype=torch.float64, device=dev) * 26.0).reshape(1, 3, 3)
inv = engine.invert_lattice(latt)
pitch = engine.pitch_for(T)
idx = torch.arange(n_mob, dtype=torch.int32, device=dev)
fw = torch.arange(n_mob, n_mob + n_fw, dtype=torch.int32, device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
if n_fw:
    sums = engine.frame_sums(coords, MODE_FRACTIONAL, latt, inv, 0, fw, None, pitch)
    t_fs = timeit(lambda: engine.frame_sums(coords, MODE_FRACTIONAL, latt, inv, 0, fw, None, pitch))
else:
    sums, t_fs = None, float('nan')
out = torch.empty((n_mob, 3, pitch), dtype=torch.float64, device=dev)
b_fs = n_fw * (T + 1) * 3 * esz
b_uw = n_mob * (T + 1) * 3 * esz + n_mob * 3 * pitch * 8
line = 'frame_sums %.1f us  %.0f GB/s' % (t_fs * 1e3, b_fs / t_fs / 1e6)
for name, mode in (('cp.async tile kernel', '0'), ('warp-per-atom kernel', '1'), ('bulk-copy tile kernel', '2')):
    if len(sys.argv) > 4 and sys.argv[4] != mode:
        continue
    os.environ['KB2_UW_KERNEL'] = mode.split(':')[0]
    if ':' in mode:
        os.environ['KB2_UB_OPTS'] = mode.split(':')[1]
    t_uw = timeit(lambda: engine.unwrap_cumsum(coords, MODE_FRACTIONAL, latt, inv, 0, idx, sums, 1.0 / max(n_fw, 1), pitch, out=out))
    line += '   unwrap_cumsum (%s) %.1f us  %.0f GB/s' % (name, t_uw * 1e3, b_uw / t_uw / 1e6)
print(line)
