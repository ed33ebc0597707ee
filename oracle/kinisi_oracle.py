"""
oracle/kinisi_oracle.py -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline, never the product path).

A CPU (numpy, float64) restatement of the displacement-statistics hot path of bjmorgan/kinisi 1.1.1.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module.  `kinisi_b200/` never does.

Parity status
-------------
* PINNED against the unmodified reference run in the build container (`tests/golden/make_golden.py`
  imports `/root/reference` and dumps `tests/golden/*.npz`; `tests/test_oracle_golden.py` compares)
  for: unwrapping, drift correction, the dt grid, per-dt displacement construction, n_o,
  MSD/MSTD/MSCD mean / variance / std / NGP, the model covariance, the minimum-eigenvalue
  reconditioning, slogdet + pinvh and the log-likelihood values.
* PINNED against the reference's own known-answer tests: NGP([1,2,3]) = -0.3
  (kinisi/tests/test_diffusion.py:105-107), the drift KAT (kinisi/tests/test_parser.py:151-168), the
  documentation fixture docs/source/_static/displacements.npz.
* PARITY UNPINNED for two pieces whose arithmetic lives in third-party packages that are neither
  installed here nor vendored in /root/reference and whose results the reference's tests do not pin:
  `emcee.EnsembleSampler` (stretch move; unpinned version, the RNG is unseedable from kinisi,
  diffusion.py:386-389) and `statsmodels.stats.correlation_tools.cov_nearest` (statsmodels>=0.14).
  Both are restated from their published algorithms (see oracle/shims/README.md).  The same holds for
  `pyblock.blocking.reblock` / `find_optimal_block` behind `block=True` (diffusion.py:584-595): pyblock is absent, no
  reference test pins a blocked result; `reblock` / `find_optimal_block` below restate Flyvbjerg & Petersen (1989).
* PINNED bit for bit for `bootstrap=True`: `bootstrap_means` / `sample_until_normal` against the reference +
  scikit-learn with seeded RandomStates (tests/golden/resample_reference.npz).

Every function cites the reference file:line it follows (paths relative to /root/reference).
The restatement is *streaming*: it evaluates the same numpy expressions one dt at a time and never
retains the O(N*T*n_dt) displacement list, so that it can serve as a CPU baseline for sizes where the
reference raises MemoryError (parser.py:196-200).
"""
import numpy as np
import scipy.constants as const
from scipy.linalg import pinvh
from scipy.optimize import minimize
from scipy.stats import linregress

# kinisi/diffusion.py:23-38
DIMENSIONALITY = {
    'x': np.s_[0],
    'y': np.s_[1],
    'z': np.s_[2],
    'xy': np.s_[:2],
    'xz': np.s_[::2],
    'yz': np.s_[1:],
    'xyz': np.s_[:],
}


# --------------------------------------------------------------------------------------------
# Stage 1: displacement construction
# --------------------------------------------------------------------------------------------
def get_disp(coords, latt):
    """kinisi/parser.py:101-129 (Parser.get_disp).

    coords: fractional coordinates [N, F, 3] (F frames, the caller has already duplicated frame 0
    as the front-ends do, parser.py:323-324); latt: [F, 3, 3].  Returns dc [N, F-1, 3], dc[:, 0] = 0.
    """
    coords = np.asarray(coords, dtype=np.float64)
    latt = np.asarray(latt, dtype=np.float64)
    latt_inv = np.linalg.inv(latt)
    wrapped = np.einsum('ijk,jkl->ijl', coords, latt)
    wrapped_diff = np.diff(wrapped, axis=1)
    images = np.floor(np.einsum('ijk,jkl->ijl', wrapped_diff, latt_inv[1:]) + 0.5)
    unwrapped = wrapped_diff - np.einsum('ijk,jkl->ijl', images, latt[1:])
    return np.cumsum(unwrapped, axis=1)


def universe_coords_latt(positions, cells):
    """kinisi/parser.py:620-638 (MDAnalysisParser.get_structure_coords_latt): per frame, Cartesian positions times the
    inverse cell, in the arrays' own precision (MDAnalysis hands out float32: numpy then inverts and multiplies in
    single precision); the first frame is duplicated.  positions [F, N, 3], cells [F, 3, 3] -> (coords [N, F + 1, 3],
    latt [F + 1, 3, 3])."""
    coords, latt = [], []
    for pos, matrix in zip(positions, cells):
        inv_matrix = np.linalg.inv(matrix)
        coords.append(np.array(np.dot(pos, inv_matrix))[:, None])
        latt.append(matrix)
    coords.insert(0, coords[0])
    latt.insert(0, latt[0])
    return np.concatenate(coords, axis=1), np.array(latt)


def correct_drift(drift_indices, disp):
    """kinisi/parser.py:131-148 (Parser.correct_drift)."""
    if len(drift_indices) > 0:
        return disp - np.average(disp[drift_indices], axis=0)[None, :, :]
    return disp


def molecule_centres(coords, members, masses=None):
    """kinisi/parser.py:706-724 (_get_molecules): pseudo-centre-of-mass of groups of atoms (arXiv:2501.14578).

    coords: fractional coordinates [N, F, 3] (site-major, this module's layout); members: 0-based [n_mol, m];
    masses: [m] or None.  Returns the centres [n_mol, F, 3] in [0, 1)."""
    coords = np.asarray(coords, dtype=np.float64)
    members = np.asarray(members)
    s_coords = np.transpose(coords[members], (2, 0, 1, 3))  # [F, n_mol, m, 3] as sq_coords[:, indices]
    theta = s_coords * (2 * np.pi)
    xi = np.cos(theta)
    zeta = np.sin(theta)
    xi_bar = np.average(xi, axis=-2, weights=masses)
    zeta_bar = np.average(zeta, axis=-2, weights=masses)
    theta_bar = np.arctan2(-zeta_bar, -xi_bar) + np.pi
    new_s_coords = theta_bar / (2 * np.pi)
    recentred = (s_coords - (new_s_coords + 0.5)[:, :, np.newaxis]) % 1
    com_pseudo_space = np.average(recentred, weights=masses, axis=2)
    corrected = (com_pseudo_space + (new_s_coords + 0.5)) % 1
    return np.transpose(corrected, (1, 0, 2))


def time_intervals(n_frames, time_step, step_skip, min_dt=None, max_dt=None, n_steps=100, spacing='linear'):
    """kinisi/parser.py:80-90 (defaults and clamps) + :150-168 (get_time_intervals).

    Returns the integer lag grid.  n_frames is drift_corrected.shape[1].
    """
    unit = step_skip * time_step
    if max_dt is None:
        max_dt = n_frames * unit
    if min_dt is None:
        min_dt = unit
    min_dt_int = int(min_dt / unit)
    if min_dt_int >= n_frames:
        raise ValueError('The passed min_dt is greater than or equal to the maximum simulation length.')
    if n_steps > (n_frames - min_dt_int) + 1:
        n_steps = int(n_frames - min_dt_int) + 1
    lo = int(min_dt / unit)
    hi = int(max_dt / unit)
    if lo == 0:
        lo = 1
    if spacing == 'linear':
        return np.linspace(lo, hi, n_steps, dtype=int)
    if spacing == 'logarithmic':
        return np.unique(np.geomspace(lo, hi, n_steps, dtype=int))
    raise ValueError("Only linear or logarithmic spacing is allowed.")


def displacements_for_lag(dc_sel, k):
    """kinisi/parser.py:208-213: one entry of disp_3d for lag k (multi-origin sampling).

    Observation 0 is dc[:, k-1] (the (k-1)-step displacement from the duplicated first frame);
    observations 1.. are dc[:, j+k] - dc[:, j].
    """
    return np.concatenate([dc_sel[:, np.newaxis, k - 1], np.subtract(dc_sel[:, k:], dc_sel[:, :-k])], axis=1)


def lag_is_kept(n_atoms, n_obs, k):
    """kinisi/parser.py:214-215: a lag is dropped from disp_3d / n_o when
    n_atoms * ceil(n_obs / k) <= 1."""
    return n_atoms * len(range(0, n_obs, k)) > 1


def memory_estimate_gb(n_atoms, n_frames, lags, itemsize=8):
    """kinisi/parser.py:190-195 (the MemoryError guard's estimate)."""
    mem = 0
    for k in lags:
        mem += n_atoms * max(n_frames - int(k), 0) * 3 * itemsize
        mem += n_atoms * 3 * itemsize
    return mem * 1e-9


def parse(dc, indices, drift_indices, time_step, step_skip, min_dt=None, max_dt=None, n_steps=100, spacing='linear'):
    """kinisi/parser.py:53-92 (Parser.__init__) without materialising disp_3d.

    Returns dict(dc_sel, lags (kept only), delta_t (ALL lags, parser.py:183), n_o (kept only)).
    """
    dc = correct_drift(drift_indices, np.asarray(dc, dtype=np.float64))
    lags_all = time_intervals(dc.shape[1], time_step, step_skip, min_dt, max_dt, n_steps, spacing)
    delta_t = lags_all * time_step * step_skip
    dc_sel = dc[indices]
    n_atoms, n_frames = dc_sel.shape[0], dc_sel.shape[1]
    kept, n_o = [], []
    for k in lags_all:
        n_obs = n_frames - k + 1
        if not lag_is_kept(n_atoms, n_obs, k):
            continue
        kept.append(int(k))
        n_o.append(n_atoms * lags_all[-1] / k)  # parser.py:221
    return {
        'dc': dc,
        'dc_sel': dc_sel,
        'lags_all': lags_all,
        'lags': np.array(kept, dtype=int),
        'delta_t': delta_t,
        'n_o': np.array(n_o, dtype=float)
    }


# --------------------------------------------------------------------------------------------
# Stage 2: moment reductions
# --------------------------------------------------------------------------------------------
def ngp_calculation(d_squared):
    """kinisi/diffusion.py:291-303."""
    top = np.mean(d_squared * d_squared) * 3
    bottom = np.square(np.mean(d_squared)) * 5
    return top / bottom - 1


def _slice_dims(disp, dimension):
    sl = DIMENSIONALITY[dimension.lower()]
    dims = len(dimension)
    return disp[:, :, sl].reshape(disp.shape[0], disp.shape[1], dims)


def moments_from_disp(disp, n_o_i, dimension='xyz', kind='msd', ionic_charge=None):
    """One iteration of the default (bootstrap=False, block=False) branch of
    MSDBootstrap.__init__ (kinisi/diffusion.py:571-601), MSTDBootstrap.__init__ (:655-687) or
    MSCDBootstrap.__init__ (:747-779) on one materialised displacement array [N, M, 3].

    Returns (n, v, ngp) or None when the reference `continue`s (size <= 1).
    """
    disp_slice = _slice_dims(disp, dimension)
    d_squared = np.sum(disp_slice**2, axis=-1)
    if kind == 'msd':
        if d_squared.size <= 1:
            return None
        n = d_squared.mean()
        v = np.var(d_squared, ddof=1) / n_o_i
        return n, v, ngp_calculation(d_squared)
    if kind == 'mstd':
        coll = np.sum(np.sum(disp_slice, axis=0)**2, axis=-1)  # diffusion.py:659
    elif kind == 'mscd':
        coll = np.sum(np.sum((ionic_charge * disp.T).T, axis=0)**2, axis=-1)  # diffusion.py:751 (unsliced!)
    else:
        raise ValueError(kind)
    if coll.size <= 1:
        return None
    n = coll.mean()
    v = np.var(coll, ddof=1) / (n_o_i / d_squared.shape[0])
    return n, v, ngp_calculation(d_squared.flatten())


def bootstrap_from_list(delta_t, disp_3d, n_o, kind='msd', ionic_charge=None, sub_sample_dt=1, dimension='xyz'):
    """The *Bootstrap constructors on a materialised list (the reference seam, diffusion.py:552-602,
    636-688, 723-780)."""
    disps = disp_3d[::sub_sample_dt]
    delta_t = np.array(delta_t[::sub_sample_dt])
    n_iter = len(disps) if kind == 'msd' else int(len(disps) / 2)  # diffusion.py:650, 738
    if kind == 'mscd':
        try:
            len(ionic_charge)
            ionic_charge = np.asarray(ionic_charge, dtype=float)
        except TypeError:
            ionic_charge = np.ones(disps[0].shape[0]) * ionic_charge
    out = {k: [] for k in ('dt', 'n', 'v', 's', 'ngp')}
    for i in range(n_iter):
        r = moments_from_disp(disps[i], n_o[i], dimension, kind, ionic_charge)
        if r is None:
            continue
        out['n'].append(r[0])
        out['v'].append(r[1])
        out['s'].append(np.sqrt(r[1]))
        out['ngp'].append(r[2])
        out['dt'].append(delta_t[i])
    res = {k: np.array(v, dtype=float) for k, v in out.items()}
    res['n_o'] = np.asarray(n_o)[:res['n'].size]  # diffusion.py:602
    return res


def bootstrap_streaming(parsed, kind='msd', ionic_charge=None, sub_sample_dt=1, dimension='xyz'):
    """Stages 1->2 fused on the CPU: same numpy expressions as `bootstrap_from_list`, one lag at a
    time, nothing retained.  `parsed` is the dict from `parse`."""
    dc_sel = parsed['dc_sel']
    lags = parsed['lags'][::sub_sample_dt]
    delta_t = np.array(parsed['delta_t'][::sub_sample_dt])
    n_o = parsed['n_o']
    n_iter = len(lags) if kind == 'msd' else int(len(lags) / 2)
    if kind == 'mscd':
        try:
            len(ionic_charge)
            ionic_charge = np.asarray(ionic_charge, dtype=float)
        except TypeError:
            ionic_charge = np.ones(dc_sel.shape[0]) * ionic_charge
    out = {k: [] for k in ('dt', 'n', 'v', 's', 'ngp')}
    for i in range(n_iter):
        disp = displacements_for_lag(dc_sel, int(lags[i]))
        r = moments_from_disp(disp, n_o[i], dimension, kind, ionic_charge)
        if r is None:
            continue
        out['n'].append(r[0])
        out['v'].append(r[1])
        out['s'].append(np.sqrt(r[1]))
        out['ngp'].append(r[2])
        out['dt'].append(delta_t[i])
    res = {k: np.array(v, dtype=float) for k, v in out.items()}
    res['n_o'] = np.asarray(n_o)[:res['n'].size]
    return res


def raw_sums_streaming(dc_sel, lags, dimension='xyz', weights=None, collective=False):
    """Sufficient statistics per lag, for checking kernel outputs directly: returns S1, S2 (sums of
    d^2 and d^4 over atoms x observations) and, if `collective`, C1, C2 (sums of c and c^2 with
    c_j = |sum_a w_a disp_a(j)|^2, all three components when weights is given as the MSCD quirk
    diffusion.py:751 demands, else the sliced components, diffusion.py:659)."""
    S1, S2, C1, C2 = [], [], [], []
    for k in lags:
        disp = displacements_for_lag(dc_sel, int(k))
        d2 = np.sum(_slice_dims(disp, dimension)**2, axis=-1)
        S1.append(d2.sum())
        S2.append((d2 * d2).sum())
        if collective:
            if weights is None:
                c = np.sum(np.sum(_slice_dims(disp, dimension), axis=0)**2, axis=-1)
            else:
                c = np.sum(np.sum((np.asarray(weights, dtype=float) * disp.T).T, axis=0)**2, axis=-1)
            C1.append(c.sum())
            C2.append((c * c).sum())
    return tuple(np.array(x, dtype=float) for x in (S1, S2, C1, C2))


# --------------------------------------------------------------------------------------------
# bootstrap=True: resampled means
# --------------------------------------------------------------------------------------------
def bootstrap_means(array, n_samples, n_resamples, random_state=None):
    """kinisi/diffusion.py:783-801 (`_bootstrap`).  The reference calls sklearn.utils.resample(values,
    n_samples=int(n_samples), random_state=rs) per resample, which for replace=True takes
    `rs.randint(0, len(values), size=(n,))` and indexes (scikit-learn, sklearn/utils/_indexing.py `resample`;
    pinned bit for bit against the reference + scikit-learn 1.9 in tests/golden/resample_reference.npz)."""
    rs = np.random.mtrand._rand if random_state is None else random_state
    values = np.asarray(array).flatten()
    return [float(np.mean(values[rs.randint(0, values.shape[0], size=(int(n_samples), ))])) for _ in range(n_resamples)]


def sample_until_normal(array, n_samples, n_resamples, max_resamples, alpha=1e-3, random_state=None):
    """kinisi/diffusion.py:258-289: resample, then add 100 at a time until scipy.stats.normaltest passes or
    max_resamples is reached.  Returns (values, reached_max)."""
    from scipy.stats import normaltest
    values = bootstrap_means(array, n_samples, n_resamples, random_state)
    p_value = normaltest(values)[1]
    while p_value < alpha and len(values) < max_resamples:
        values += bootstrap_means(array, n_samples, 100, random_state)
        p_value = normaltest(values)[1]
    return np.array(values), len(values) >= max_resamples


def reblock(data, ddof=1):
    """pyblock.blocking.reblock for one variable (the call at kinisi/diffusion.py:585, 671, 763).  PARITY UNPINNED: pyblock
    (unpinned in the reference's pyproject.toml, imported lazily) is neither installed here nor vendored under
    /root/reference and no reference test pins a `block=True` result; restated from the published algorithm
    (Flyvbjerg & Petersen, J. Chem. Phys. 91, 461 (1989); pyblock 0.6 documentation): at every level the mean, the
    ddof=1 variance, std_err = sqrt(var / n) and std_err_err = std_err / sqrt(2 (n - ddof)); the next level averages
    neighbouring pairs, dropping the last point of an odd level.  Returns [(block, ndata, mean, cov, std_err, std_err_err)]."""
    data = np.asarray(data, dtype=np.float64).flatten()
    stats, iblock = [], 0
    while data.shape[0] >= 2:
        n = data.shape[0]
        mean = data.mean()
        cov = float(np.cov(data, ddof=ddof))
        std_err = np.sqrt(cov / n)
        stats.append((iblock, n, mean, cov, std_err, std_err / np.sqrt(2 * (n - ddof))))
        last = 2 * (n // 2)
        data = (data[:last:2] + data[1:last:2]) / 2
        iblock += 1
    return stats


def find_optimal_block(ndata, stats):
    """pyblock.blocking.find_optimal_block for one variable (same status as `reblock`): the smallest level whose block
    size B = 2^level satisfies B^3 > 2 n (std_err(B) / std_err(0))^4 (Wolff; Lee et al.), nan when no level does."""
    optimal = float('inf')
    first = stats[0][4]
    with np.errstate(all='ignore'):
        for (iblock, _, _, _, std_err, _) in reversed(stats):
            if 2**(3 * iblock) > 2 * ndata * (std_err / first)**4:
                optimal = iblock
    return float('nan') if np.isinf(optimal) else optimal


def blocked_mean_and_variance(values):
    """kinisi/diffusion.py:584-595: mean and squared standard error at the optimal block, the last level when there is
    none (the reference's `except TypeError` on a nan index)."""
    stats = reblock(values)
    opt = find_optimal_block(np.asarray(values).size, stats)
    lvl = stats[-1] if np.isnan(opt) else stats[int(opt)]
    return lvl[2], lvl[4]**2


_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
_U32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), vectorised over uint64-held 32-bit words: the published
    generator the device resampler draws from (no counterpart in the reference, whose indices come from numpy's
    Mersenne twister; the two can only agree in distribution)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _U32 for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _PHILOX_M0 * c0
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _U32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _U32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _U32, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _U32, lo0
        k0, k1 = (k0 + _PHILOX_W0) & 0xFFFFFFFF, (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _mulhi64(u, m):
    """floor(u * m / 2^64) for uint64 arrays u and an integer m < 2^64, by 32-bit limbs."""
    m = int(m)
    ul, uh = u & _U32, u >> np.uint64(32)
    ml, mh = np.uint64(m & 0xFFFFFFFF), np.uint64(m >> 32)
    ll, lh, hl, hh = ul * ml, ul * mh, uh * ml, uh * mh
    mid = (ll >> np.uint64(32)) + (lh & _U32) + (hl & _U32)
    return hh + (lh >> np.uint64(32)) + (hl >> np.uint64(32)) + (mid >> np.uint64(32))


def counter_resample_indices(n_values, n_draws, resample, seed, stream_id=0):
    """The indices draw 0 .. n_draws - 1 of one resample take in kb2_resample_means (include/kinisi_b200.h): draw d
    uses half (d & 1) of Philox4x32-10(counter {d >> 1 lo, d >> 1 hi, resample, stream_id}, key seed)."""
    pairs = np.arange((n_draws + 1) // 2, dtype=np.uint64)
    x = philox4x32_10(pairs & _U32, pairs >> np.uint64(32), np.full_like(pairs, resample), np.full_like(pairs, stream_id),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = np.empty(2 * pairs.size, dtype=np.uint64)
    u[0::2] = (x[1] << np.uint64(32)) | x[0]
    u[1::2] = (x[3] << np.uint64(32)) | x[2]
    return _mulhi64(u[:n_draws], n_values).astype(np.int64)


def counter_resample_means(values, n_draws, n_resamples, first_resample=0, seed=0, stream_id=0):
    """`_bootstrap` (kinisi/diffusion.py:783-801) with the counter-based indices above instead of numpy's."""
    values = np.asarray(values, dtype=np.float64).flatten()
    return np.array([
        np.mean(values[counter_resample_indices(values.size, int(n_draws), first_resample + r, seed, stream_id)])
        for r in range(n_resamples)
    ])


# --------------------------------------------------------------------------------------------
# Stage 3: covariance + Bayesian regression
# --------------------------------------------------------------------------------------------
def populate_covariance_matrix(variances, n_samples):
    """kinisi/diffusion.py:838-854: C[i,j] = C[j,i] = (n_o[i] / n_o[j]) * v[i] for i <= j."""
    v = np.asarray(variances, dtype=float)
    no = np.asarray(n_samples, dtype=float)
    ratio = no[:, None] / no[None, :]
    upper = ratio * v[:, None]
    iu = np.triu(np.ones((v.size, v.size), dtype=bool))
    return np.where(iu, upper, upper.T)


def minimum_eigenvalue_method(matrix, cond_max=1e16):
    """kinisi/diffusion.py:870-888.  Uses the general `np.linalg.eig` and thresholds on the FIRST
    returned eigenvalue, exactly as the reference does."""
    eigenthings = np.linalg.eig(matrix)
    eigenvalues = eigenthings.eigenvalues
    new_eigenvalues = np.copy(eigenvalues)
    T = eigenvalues[0] / cond_max
    new_eigenvalues[np.where(new_eigenvalues < T)] = T
    return np.real(eigenthings.eigenvectors @ np.diag(new_eigenvalues) @ eigenthings.eigenvectors.T)


def cov_nearest(cov, threshold=1e-15):
    """statsmodels.stats.correlation_tools.cov_nearest(method='clipped') as called at
    kinisi/diffusion.py:425.  Restated from the published algorithm -- PARITY UNPINNED."""
    cov = np.asarray(cov)
    std = np.sqrt(np.diag(cov))
    corr = cov / np.outer(std, std)
    evals, evecs = np.linalg.eigh(corr)
    if np.any(evals < threshold):
        x_new = np.dot(evecs * np.maximum(evals, threshold), evecs.T)
        x_std = np.sqrt(np.diag(x_new))
        corr = x_new / x_std / x_std[:, None]
    return corr * np.outer(std, std)


def model_variance(dt, v, n_o, diff_regime):
    """kinisi/diffusion.py:406-420 with model=True: v ~ a / n_o * dt**2, `a` fitted with curve_fit under a >= 0."""
    from scipy.optimize import curve_fit

    def _model_variance(x, a):
        return a / n_o[diff_regime:] * x**2

    popt, _ = curve_fit(_model_variance, dt[diff_regime:], v[diff_regime:], bounds=(0, np.inf))
    return _model_variance(dt[diff_regime:], *popt)


def generate_covariance_matrix(v, n_o, diff_regime, cond_max=1e16, dt=None, model=False):
    """kinisi/diffusion.py:397-425."""
    model_v = model_variance(dt, v, n_o, diff_regime) if model else v[diff_regime:]
    cov = populate_covariance_matrix(model_v, n_o[diff_regime:])
    return cov_nearest(minimum_eigenvalue_method(cov, cond_max))


def straight_line(abscissa, gradient, intercept=0.0):
    """kinisi/diffusion.py:857-867."""
    return gradient * abscissa + intercept


def make_log_likelihood(dt, n, cov):
    """kinisi/diffusion.py:350-365: slogdet + n ln 2pi, pinvh, and the closure."""
    _, logdet = np.linalg.slogdet(cov)
    logdet += np.log(2 * np.pi) * n.size
    inv = pinvh(cov)

    def log_likelihood(theta):
        if theta[0] < 0:
            return -np.inf
        diff = straight_line(dt, *theta) - n
        return -0.5 * (logdet + np.matmul(diff.T, np.matmul(inv, diff)))

    return log_likelihood, logdet, inv


def gls_posterior(dt, n, inv, fit_intercept=True):
    """Closed-form Gaussian posterior of theta under the (flat-prior) likelihood above, ignoring the
    slope >= 0 truncation: mean = (X^T A X)^-1 X^T A n, cov = (X^T A X)^-1 with A = pinvh(C).  Used as
    an exact check for 'within Monte-Carlo error'."""
    X = np.stack([dt, np.ones_like(dt)], axis=1) if fit_intercept else dt[:, None]
    XtA = X.T @ inv
    prec = XtA @ X
    cov_theta = np.linalg.inv(prec)
    return cov_theta @ (XtA @ n), cov_theta


def stretch_move_sampler(log_prob_fn, pos, nsteps, rng, a=2.0):
    """emcee 3.x EnsembleSampler with the default StretchMove, as called at diffusion.py:385-390.
    Restated from the published algorithm -- PARITY UNPINNED.  Returns chain [nsteps, W, p]."""
    x = np.array(pos, dtype=float)
    nw, nd = x.shape
    lp = np.array([log_prob_fn(xi) for xi in x])
    chain = np.empty((nsteps, nw, nd))
    for t in range(nsteps):
        inds = np.arange(nw) % 2
        rng.shuffle(inds)
        for split in range(2):
            active = np.where(inds == split)[0]
            comp = x[inds != split]
            zz = ((a - 1.0) * rng.rand(active.size) + 1)**2.0 / a
            rint = rng.randint(comp.shape[0], size=active.size)
            q = comp[rint] - (comp[rint] - x[active]) * zz[:, None]
            new_lp = np.array([log_prob_fn(qi) for qi in q])
            acc = ((nd - 1.0) * np.log(zz) + new_lp - lp[active]) > np.log(rng.rand(active.size))
            x[active[acc]] = q[acc]
            lp[active[acc]] = new_lp[acc]
        chain[t] = x
    return chain


def get_chain(chain, flat=True, thin=1, discard=0):
    """emcee backend `get_chain` as called at diffusion.py:391."""
    v = chain[discard + thin - 1::thin]
    return v.reshape(-1, chain.shape[-1]) if flat else v


def bootstrap_gls(stats,
                  start_dt,
                  cond_max=1e16,
                  fit_intercept=True,
                  n_samples=1000,
                  n_walkers=32,
                  n_burn=500,
                  thin=10,
                  random_state=None):
    """kinisi/diffusion.py:305-395 (bootstrap_GLS, model=False).  `stats` is a dict with dt, n, v, n_o."""
    rng = random_state if random_state is not None else np.random.RandomState()
    dt, n, v, n_o = stats['dt'], stats['n'], stats['v'], stats['n_o']
    diff_regime = np.argwhere(dt >= start_dt)[0][0]
    cov = generate_covariance_matrix(v, n_o, diff_regime, cond_max)
    x, y = dt[diff_regime:], n[diff_regime:]
    log_likelihood, logdet, inv = make_log_likelihood(x, y, cov)
    slope = linregress(x, y).slope
    if slope < 0:
        slope = 1e-20
    start = np.array([slope, 1e-20]) if fit_intercept else np.array([slope])
    max_likelihood = minimize(lambda th: -log_likelihood(th), start).x
    pos = max_likelihood + max_likelihood * 1e-3 * rng.randn(n_walkers, max_likelihood.size)
    chain = stretch_move_sampler(log_likelihood, pos, n_samples + n_burn, rng)
    flatchain = get_chain(chain, True, thin, n_burn)
    return {
        'covariance_matrix': cov,
        'logdet': logdet,
        'inv': inv,
        'max_likelihood': max_likelihood,
        'flatchain': flatchain,
        'gradient': flatchain[:, 0],
        'intercept': flatchain[:, 1] if fit_intercept else None,
        'diff_regime': diff_regime
    }


def posterior_predictive(gradient, intercept, dt, cov, n_posterior_samples, n_predictive_samples, seed):
    """kinisi/diffusion.py:509-521: for each of `n_posterior_samples` draws (with replacement) of (gradient,
    intercept), `n_predictive_samples` draws of a multivariate normal about the straight line with the reconditioned
    covariance (`allow_singular=True`).  Global numpy RNG as in the reference, seeded here.  Returns (ppd, drawn)."""
    from scipy.stats import multivariate_normal
    np.random.seed(seed)
    ppd = np.zeros((n_posterior_samples, n_predictive_samples, dt.size))
    drawn = np.random.randint(0, gradient.size, size=n_posterior_samples)
    for i, n in enumerate(drawn):
        mu = gradient[n] * dt + intercept[n]
        mv = multivariate_normal(mean=mu, cov=cov, allow_singular=True)
        ppd[i] = mv.rvs(n_predictive_samples)
    return ppd.reshape(-1, dt.size), drawn


def diffusion_coefficient(gradient, dims):
    """kinisi/diffusion.py:436: D [cm^2/s] from the MSD gradient [A^2/ps]."""
    return gradient / (2e4 * dims)


def jump_diffusion_coefficient(gradient, dims, n_atoms):
    """kinisi/diffusion.py:456-457."""
    return gradient / (2e4 * dims * n_atoms)


def conductivity(gradient, dims, volume, temperature):
    """kinisi/diffusion.py:479-482: sigma [mS/cm]."""
    volume = volume * 1e-24
    D = gradient / (2e4 * dims)
    conversion = 1000 / (volume * const.N_A) * (const.N_A * const.e)**2 / (const.R * temperature)
    return D * conversion


# --------------------------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench (seeded, reproducible)
# --------------------------------------------------------------------------------------------
def synthetic_walk(n_atoms, n_frames, box=15.0, sigma=0.1, seed=0, n_framework=0, fw_jitter=0.05, fw_drift=0.01,
                   dtype=np.float64):
    """Seeded 3-D Gaussian random walk stored WRAPPED into a cubic cell.

    Returns (frac [N_all, n_frames+1, 3] with frame 0 duplicated as the front-ends do
    (parser.py:323-324), latt [n_frames+1, 3, 3]).  Mobile atoms come first, then framework atoms
    that jitter about fixed sites and share a common linear drift.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n_all = n_atoms + n_framework
    start = rng.uniform(0, box, size=(n_all, 1, 3))
    steps = rng.normal(0.0, sigma, size=(n_atoms, n_frames, 3))
    steps[:, 0] = 0.0
    pos = np.empty((n_all, n_frames, 3))
    pos[:n_atoms] = start[:n_atoms] + np.cumsum(steps, axis=1)
    if n_framework:
        jitter = rng.normal(0.0, fw_jitter, size=(n_framework, n_frames, 3))
        drift = fw_drift * np.arange(n_frames)[None, :, None]
        pos[n_atoms:] = start[n_atoms:] + jitter + drift
    frac = (pos / box) % 1.0
    frac = np.concatenate([frac[:, :1], frac], axis=1).astype(dtype)
    latt = np.tile(np.eye(3) * box, (n_frames + 1, 1, 1))
    return frac, latt
