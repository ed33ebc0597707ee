"""The tensor-core Gram route of BASELINE.json's north_star, measured against the direct FP64 kernel at configs[1]
(448 atoms x 20 000 frames, 100 lags): the cross term of |x(t+k) - x(t)|^2 as a time-Gram contraction over atoms x 3,
G = X X^T with X [T, 3N], in split-bf16 precision on the tensor cores (torch.matmul = cuBLAS, fp32 accumulation):

    sum_a |x_a(j+k) - x_a(j)|^2 = G[j+k, j+k] + G[j, j] - 2 G[j+k, j]

Only the MEAN squared displacement comes out of G: the variance needs sum_j (sum over components of one atom)^2, i.e. the
per-atom squares BEFORE the atom sum, which the contraction over atoms has already lost.  Reported: time of the GEMMs (the
whole T x T Gram: the 100 lags of the grid are 100 diagonals spread over the whole matrix, there is no band to restrict
to), and the largest relative error of the mean MSD over the grid, against the FP64 kernel's time and its own error level.
Run on a B200: python tools/gram_probe.py > profiles/r02_gram_probe.txt"""
import json
import sys

sys.path.insert(0, '/root/repo')
import numpy as np
import torch

from kinisi_b200 import engine

dev = torch.device('cuda')
gen = torch.Generator(device=dev)
gen.manual_seed(1002)
n_atoms, T = 448, 20000
pitch = engine.pitch_for(T)
dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
steps = torch.randn((n_atoms, 3, T), dtype=torch.float64, device=dev, generator=gen) * 0.1
steps[:, :, 0] = 0
dc[:, :, :T] = torch.cumsum(steps, dim=2)
lags = np.linspace(1, T, 100, dtype=int)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


# reference: the FP64 kernel (first-observation terms removed so that both sides are the plain multi-origin sum)
t_direct, sums = timeit(lambda: engine.self_moments(dc, T, lags, 7), 10)
x = dc[:, :, :T]
first = torch.stack([(x[:, :, int(k) - 1]**2).sum() for k in lags])
s1_ref = (sums[:, 0] - first).cpu().numpy()
cnt = n_atoms * (T - lags).astype(np.float64)
ok = cnt > 0
msd_ref = np.where(ok, s1_ref / np.maximum(cnt, 1), 0.0)

X = x.permute(2, 0, 1).reshape(T, 3 * n_atoms).contiguous()  # [T, 3N] float64


def split_bf16(a, parts):
    out, r = [], a.clone()
    for _ in range(parts):
        h = r.to(torch.bfloat16)
        out.append(h)
        r = r - h.to(torch.float64)
    return out


def msd_from_gram(G):  # G [T, T] (float32 or float64)
    d = torch.diagonal(G).to(torch.float64)
    out = []
    for k in lags:
        k = int(k)
        if k >= T:
            out.append(0.0)
            continue
        cross = torch.diagonal(G, offset=-k).to(torch.float64)
        out.append(float((d[k:] + d[:T - k] - 2.0 * cross).sum()))
    return np.array(out) / np.maximum(cnt, 1)


res = {'direct_fp64_kernel_ms': t_direct, 'note': 'mean MSD only from the Gram routes (no fourth moment / variance)'}
for name, parts in (('bf16 x1 (1 GEMM)', 1), ('bf16 split 2 (3 GEMMs)', 2), ('bf16 split 3 (6 GEMMs)', 3)):
    P = split_bf16(X, parts)
    pairs = [(i, j) for i in range(parts) for j in range(parts) if i + j < parts]

    def gram():
        G = None
        for i, j in pairs:
            g = torch.matmul(P[i], P[j].t()).to(torch.float32)
            G = g if G is None else G.add_(g)
        return G

    ms, G = timeit(gram, 3)
    msd = msd_from_gram(G)
    err = float(np.max(np.abs(msd[ok] - msd_ref[ok]) / msd_ref[ok]))
    res[name] = {'gemm_ms': ms, 'n_gemms': len(pairs), 'tflops_bf16': 2.0 * T * T * 3 * n_atoms * len(pairs) / (ms * 1e-3) / 1e12,
                 'max_rel_err_mean_msd': err}
    del G, P
    torch.cuda.empty_cache()
for name, dt, flag in (('tf32 (1 GEMM)', torch.float32, True), ('fp32 (1 GEMM, no tensor cores)', torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = flag
    Xf = X.to(dt)
    ms, G = timeit(lambda: torch.matmul(Xf, Xf.t()), 3)
    msd = msd_from_gram(G)
    res[name] = {'gemm_ms': ms, 'max_rel_err_mean_msd': float(np.max(np.abs(msd[ok] - msd_ref[ok]) / msd_ref[ok]))}
    del G
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
