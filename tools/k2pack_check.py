"""The tiled moment kernel with ragged last strips packed (several atoms side by side) against the direct kernel: atom
counts around the batch and pair boundaries, all dimension masks, even and odd row lengths; then the time at configs[1]."""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from kinisi_b200 import engine
dev = torch.device('cuda')
bad = 0
def make(n_atoms, T, seed=1):
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    pitch = engine.pitch_for(T)
    dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dc[:, :, :T] = torch.cumsum(torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1, dim=2)
    return dc
cases = [(T, n) for T in (20000, 6500, 4200, 3333, 1300, 707) for n in (100, 33, 20)] + [(20000, 2000), (8000, 250)]
for T, n in cases:
    lags = np.unique(np.linspace(1, T, n, dtype=int)).astype(np.int32)[:300]
    wins, _ = engine.self_moments_tiles(lags, T)
    Ds = sorted(set(w[1] for w in wins if w[2] > 1))
    packed = [D for D in Ds if 0 < D % 32 <= 12]
    for n_atoms in (1, 2, 3, 50, 51, 101):
        dc = make(n_atoms, T, seed=n_atoms)
        for mask in (7, 5, 2):
            a = engine.self_moments(dc, T, lags, mask, 4)
            b = engine.self_moments(dc, T, lags, mask, 0)
            err = float(((a - b).abs() / b.abs()).max())
            if not err < 1e-11:
                bad += 1
                print('MISMATCH', T, n, n_atoms, mask, err, 'D', Ds)
    print('ok' if not bad else 'BAD', T, n, 'row lengths', Ds, 'packed', packed, flush=True)
print('mismatches', bad)
T = 20000
lags = np.linspace(1, T, 100, dtype=int).astype(np.int32)
terms_per_atom = float(np.sum(T - lags + 1))
for n_atoms in (448, 447, 2000):
    dc = make(n_atoms, T)
    for _ in range(3):
        out = engine.self_moments(dc, T, lags, 7, 4)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = engine.self_moments(dc, T, lags, 7, 4)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    ref = engine.self_moments(dc, T, lags, 7, 0)
    print(f'{n_atoms} x {T}: {best:.4f} ms  {16 * n_atoms * terms_per_atom / best / 1e9 / 35.5:.3f} of FP64 peak  pack={os.environ.get("KB2_TILE_PACK", "1")}  err {float(((out - ref).abs() / ref.abs()).max()):.1e}')
