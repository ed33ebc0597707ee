"""
Host-side mirror of kinisi/diffusion.py for the displacement-statistics path: `Bootstrap`,
`MSDBootstrap`, `MSTDBootstrap`, `MSCDBootstrap` with the reference's constructor signatures,
attributes (`dt n s v ngp n_o covariance_matrix gradient intercept flatchain D D_J sigma`) and error
behaviour; the arithmetic runs in the sm_100a kernels behind include/kinisi_b200.h.

`disp_3d` may be the lazy sequence a :py:class:`kinisi_b200.parser.Parser` produces (fused path: the
displacement list is never materialised) or a plain list of `[atom, observation, axis]` arrays (the
reference's seam, kinisi/diffusion.py:552-564; each entry is reduced by one kernel).

`bootstrap=True` (diffusion.py:258-289, 783-801) resamples and `block=True` (diffusion.py:584-595) reblocks on the
device (csrc/k6_resample.cu).
"""
import warnings
from typing import Union

import numpy as np
import scipy.constants as const
import torch

from . import dist, engine
from .distribution import Distribution
from .parser import LazyDisplacements, SingleOriginDisplacements

# True: run the reference's sequence literally with library eigensolvers (eigh for the minimum-eigenvalue
# method, a second eigh inside cov_nearest, slogdet, and a third eigh for pinvh).  False (default): one
# spectral decomposition (a Jacobi kernel for n <= 128, csrc/k4_eigen.cu), reused for all of them -- the same
# result to rounding (see Bootstrap.generate_covariance_matrix and bootstrap_GLS).
LITERAL_RECONDITIONING = False

DIMENSIONALITY = {
    'x': np.s_[0],
    'y': np.s_[1],
    'z': np.s_[2],
    'xy': np.s_[:2],
    'xz': np.s_[::2],
    'yz': np.s_[1:],
    'xyz': np.s_[:],
    b'x': np.s_[0],
    b'y': np.s_[1],
    b'z': np.s_[2],
    b'xy': np.s_[:2],
    b'xz': np.s_[::2],
    b'yz': np.s_[1:],
    b'xyz': np.s_[:]
}


class _LazyEuclidian:
    """`euclidian_displacements` (kinisi/diffusion.py:577): sqrt(d^2) of every observation at each dt,
    as Distributions built on demand (O(N*T) values per dt: never on the hot path)."""

    def __init__(self, disps, used, mask_slice, dims):
        self._disps, self._used, self._slice, self._dims = disps, list(used), mask_slice, dims

    def __len__(self):
        return len(self._used)

    def __getitem__(self, i):
        d = np.asarray(self._disps[self._used[i]])
        sl = d[:, :, self._slice].reshape(d.shape[0], d.shape[1], self._dims)
        return Distribution(np.sqrt(np.sum(sl**2, axis=-1).flatten()))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _seed_from(random_state):
    """64 seed bits for the device resampler, drawn from the caller's RandomState (numpy's global one when None, as
    sklearn.utils.check_random_state does for the reference, kinisi/diffusion.py:799)."""
    rs = np.random.mtrand._rand if random_state is None else random_state
    w = rs.randint(0, 2**32, size=2, dtype=np.uint64)
    return int(w[0]) | (int(w[1]) << 32)


def _resample_until_normal(values, n_samples, n_resamples, max_resamples, alpha, seed):
    """kinisi/diffusion.py:280-289 on a device population: n_resamples means, then 100 more at a time while
    normaltest rejects and max_resamples is not reached.  The sequence is indexed by the resample number, so the
    batches made ahead of the test (up to 1000 per launch) are the values one-hundred-at-a-time calls would give."""
    from scipy.stats import normaltest
    n_draws = int(n_samples)

    def batch(first, count):
        if n_draws < 1 or values.numel() == 0:
            return np.full(count, np.nan)  # the mean of an empty resample
        return engine.resample_means(values, n_draws, count, first, seed).cpu().numpy()

    have = batch(0, int(n_resamples))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        p_value = normaltest(have)[1]
        while p_value < alpha and have.size < max_resamples:
            ahead = int(min(1000, 100 * -(-(max_resamples - have.size) // 100)))
            extra = batch(have.size, ahead)
            for u in range(0, ahead, 100):
                have = np.concatenate([have, extra[u:u + 100]])
                p_value = normaltest(have)[1]
                if not (p_value < alpha and have.size < max_resamples):
                    break
    if have.size >= max_resamples:
        warnings.warn("The maximum number of resamples has been reached, and the distribution is not yet normal.")
    return Distribution(have)


class Bootstrap:
    """
    The top-level class for the moment statistics and the Bayesian regression (kinisi/diffusion.py:41-521).

    :param delta_t: An array of the time interval values.
    :param disp_3d: The displacements: a lazy sequence from the parser or a list of arrays
        `[atom, displacement observation, dimension]`, one per delta_t.
    :param n_o: Number of statistically independent observations at each time interval.
    :param sub_sample_dt: The frequency in observations to be sampled. Optional, default is 1.
    :param dimension: Subset of 'xyz'. Optional, defaults to 'xyz'.
    """

    def __init__(self, delta_t, disp_3d, n_o, sub_sample_dt: int = 1, dimension: str = 'xyz'):
        engine.require_cuda()
        self._displacements = disp_3d[::sub_sample_dt]
        self._delta_t = np.array(delta_t[::sub_sample_dt])
        self._lazy = isinstance(self._displacements, LazyDisplacements)
        self._max_obs = self._shape_of(0)[1]
        self._distributions = []
        self._n_o = n_o
        self._sub_sample_dt = sub_sample_dt
        self._dimension = dimension
        self._euclidian_displacements = []
        self._diffusion_coefficient = None
        self._jump_diffusion_coefficient = None
        self._sigma = None
        self._intercept_dist = None
        self._gradient = None
        self._flatchain = None
        self._flat_pending = None  # (pinned buffer, event, shape, fit_intercept): the flat chain on its way to the host
        self._coefficient = None   # (attribute, factor): unit conversion applied to the gradient on first access
        self._covariance_matrix_dev = None
        self._npd_covariance_matrix_dev = None
        self._eig = None
        self._cov_regime = None
        self._slice = DIMENSIONALITY[dimension.lower()]
        self._mask = engine.dim_mask(dimension)
        self._start_dt = None
        self.dims = len(dimension.lower())
        self._model = None
        self._cond_max = None
        self._n_bootstrap = np.array([])
        self._s_bootstrap = np.array([])
        self._v_bootstrap = np.array([])
        self._stats_dev = None  # [4, n] = n, v, s, ngp on the device
        self._stats_host = None
        self._dt = np.array([])
        self._device = self._displacements.device if self._lazy else torch.device('cuda',
                                                                                     torch.cuda.current_device())

    # ---- helpers -----------------------------------------------------------------------------
    def _shape_of(self, i):
        if self._lazy:
            return self._displacements.shape_of(i)
        return tuple(self._displacements[i].shape)

    def _n_atoms(self):
        """disp.shape[0] in the reference: the (global) number of atoms of entry 0."""
        return self._displacements.n_atoms_global if self._lazy else self._shape_of(0)[0]

    def _resampling(self, bootstrap, block, n_resamples, max_resamples, alpha, random_state):
        """`bootstrap=True`: device resampler (see sample_until_normal); `block=True`: device reblocking (_blocked)."""
        if (bootstrap or block) and self._lazy and self._displacements.distributed:
            raise NotImplementedError('bootstrap=True / block=True work on one population: gather the atoms on one rank')
        if block:
            print('You are using the blocking method to estimate variances, please cite '
                  'doi:10.1063/1.457480 and the pyblock pacakge.')  # the reference's message, diffusion.py:572-573
        self._resample = (int(n_resamples), int(max_resamples), alpha, random_state) if bootstrap else None
        self._block = bool(block)
        self._atom_sum_cache = None

    @staticmethod
    def _blocked(values):
        """Mean and variance of the mean by the blocking method (kinisi/diffusion.py:584-595; pyblock.blocking.reblock
        and find_optimal_block restated -- pyblock itself is not available): every reblocking level is
        reduced on the device (kb2_reblock); the optimal level is the smallest one whose block size B = 2^level
        satisfies B^3 > 2 n (std_err(B) / std_err(0))^4, the last level when none does."""
        n0 = values.numel()
        acc = engine.reblock(values).cpu().numpy()
        ndata = np.array([n0 >> l for l in range(acc.shape[0])], dtype=np.float64)
        mean = acc[:, 0] / ndata
        with np.errstate(all='ignore'):
            std_err = np.sqrt(acc[:, 1] / (ndata - 1) / ndata)
            ok = np.power(2.0, 3 * np.arange(acc.shape[0])) > 2 * n0 * (std_err / std_err[0])**4
        lvl = int(np.flatnonzero(ok)[0]) if ok.any() else acc.shape[0] - 1
        return mean[lvl], std_err[lvl]**2

    def _apply_blocking(self, populations):
        """Replace n, v, s by the blocked estimates (the `if block:` branch; the NGP stays, diffusion.py:600)."""
        est = np.array([self._blocked(v) for v in populations], dtype=np.float64).reshape(-1, 2)
        rows = np.stack([est[:, 0], est[:, 1], np.sqrt(est[:, 1])])
        self._stats_dev[0:3] = self._dev(rows)
        self._stats_host = None

    def _population(self, i, mask, weights=None, collective=False):
        """Device array of the squared displacements of entry i: d_squared (kinisi/diffusion.py:574) or, collective,
        coll_motion (:659, 751)."""
        d = self._displacements
        if self._lazy and not isinstance(d, SingleOriginDisplacements):
            lag = int(d.lags[i])
            if not collective:
                return engine.lag_squares(d.dc, d.n_frames, lag, mask)
            if self._atom_sum_cache is None:
                self._atom_sum_cache = engine.atom_sum(d.dc, None if weights is None else self._dev(weights))
            return engine.lag_squares(self._atom_sum_cache.unsqueeze(0), d.n_frames, lag, mask)
        arr = d.device_entry(i) if isinstance(d, SingleOriginDisplacements) else engine.to_device(
            d[i], self._device, torch.float64)
        w = None if (weights is None or not collective) else self._dev(weights)
        return engine.disp_squares(arr, mask, w, collective)

    def _resample_entry(self, values, n_samples):
        """The `if bootstrap:` block of kinisi/diffusion.py:578-583, 663-669, 755-761."""
        n_resamples, max_resamples, alpha, random_state = self._resample
        distro = _resample_until_normal(values, n_samples, n_resamples, max_resamples, alpha, _seed_from(random_state))
        self._distributions.append(distro)
        self._n_bootstrap = np.append(self._n_bootstrap, np.mean(distro.samples))
        self._v_bootstrap = np.append(self._v_bootstrap, np.var(distro.samples, ddof=1))
        self._s_bootstrap = np.append(self._s_bootstrap, np.std(distro.samples, ddof=1))
        return values

    @staticmethod
    def sample_until_normal(array: np.ndarray,
                            n_samples: float,
                            n_resamples: int,
                            max_resamples: int,
                            alpha: float = 1e-3,
                            random_state: np.random.mtrand.RandomState = None) -> Distribution:
        """
        Resample from the distribution until a normal distribution is obtained or a maximum is reached
        (kinisi/diffusion.py:258-289).  Every resample is the mean of int(n_samples) values drawn with replacement;
        all resamples of a batch are made by one kernel launch (csrc/k6_resample.cu) from a counter-based generator
        seeded from `random_state` (or numpy's global state), so a RandomState in the same state gives the same
        samples, though not the values the reference's Mersenne-twister indices would give.

        :param array: The array to sample from (numpy array or CUDA tensor).
        :param n_samples: Number of samples.
        :param n_resamples: Number of resamples to perform initially.
        :param max_resamples: The maximum number of resamples to perform.
        :param alpha: Level that the p-value of :py:func:`scipy.stats.normaltest` should be below for more
            resamples to be drawn. Optional, default is :py:attr:`1e-3`.
        :param random_state: A :py:attr:`RandomState` object to be used to ensure reproducibility.

        :return: The resampled distribution.
        """
        engine.require_cuda()
        if isinstance(array, torch.Tensor):
            values = array.to(torch.float64).reshape(-1).contiguous()
        else:
            values = engine.to_device(np.asarray(array, dtype=np.float64).reshape(-1),
                                      torch.device('cuda', torch.cuda.current_device()), torch.float64)
        return _resample_until_normal(values, n_samples, n_resamples, max_resamples, alpha, _seed_from(random_state))

    def _dev(self, a):
        return engine.small_to_device(a, self._device, np.float64)

    def _const(self, a):
        """Upload of an array the kernels only read (content-addressed: see engine.const_to_device)."""
        return engine.const_to_device(a, self._device, np.float64)

    def _self_sums(self, used, mask):
        """[len(used), 2] sums of d^2, d^4 over atoms x observations for the entries in `used`."""
        d = self._displacements
        if self._lazy and not isinstance(d, SingleOriginDisplacements):
            lags = d.lags[used]
            order = np.argsort(lags, kind='stable')
            if d.n_atoms > 0:
                # atoms are additive: with a staged ingest the kernel follows the transfer block by block, and a
                # streamed sequence rebuilds its displacements block by block (parser.LazyDisplacements.blocks)
                sums = None
                want_coll = getattr(self, '_coll_request', None)
                for r0, r1, rows in d.blocks():
                    part = engine.self_moments(rows, d.n_frames, lags[order], mask)
                    sums = part if sums is None else sums.add_(part)
                    if want_coll is not None:  # the collective sum of the same pass (one stage-1 run per block)
                        w = want_coll['weights']
                        S = engine.atom_sum(rows, None if w is None else w[r0:r1].contiguous())
                        want_coll['S'] = S if want_coll['S'] is None else want_coll['S'].add_(S)
            else:
                sums = torch.zeros((len(used), 2), dtype=torch.float64, device=self._device)
            if d.distributed:
                # the sum over ranks is done by the kernel that consumes the sums (_finish -> kb2_moment_stats_joined) where
                # the grid is short enough for one slice; by a kernel of its own otherwise
                self._moments_join = dist.moments_channel(sums.numel(), sums.device)
                if self._moments_join is None:
                    dist.allreduce_sum_(sums, 'moments')
            if not np.array_equal(order, np.arange(len(order))):
                inv = np.empty_like(order)
                inv[order] = np.arange(len(order))
                sums = sums[engine.small_to_device(inv, self._device, np.int64)]
            return sums
        rows = []
        for i in used:
            arr = d.device_entry(i) if isinstance(d, SingleOriginDisplacements) else engine.to_device(
                d[i], self._device, torch.float64)
            rows.append(engine.disp_moments(arr, mask)[:2])
        sums = torch.stack(rows)
        if self._lazy and d.distributed:
            dist.allreduce_sum_(sums, 'moments')
        return sums

    def _collective_sums(self, used, weights, coll_mask):
        """[len(used), 2] sums of c, c^2 with c_j = |sum_a w_a disp_a(j)|^2 over observations."""
        d = self._displacements
        if self._lazy and not isinstance(d, SingleOriginDisplacements):
            w = None if weights is None else self._dev(weights)
            got = getattr(self, '_coll_request', None)
            if got is not None and got['S'] is not None:
                S = got['S']  # summed over the blocks of a streamed sequence by _self_sums
            elif d.n_atoms > 0:
                S = None
                for r0, r1, rows in d.blocks():
                    part = engine.atom_sum(rows, None if w is None else w[r0:r1].contiguous())
                    S = part if S is None else S.add_(part)
            else:
                S = torch.zeros((3, engine.pitch_for(d.n_frames)), dtype=torch.float64, device=self._device)
            if d.distributed:
                dist.allreduce_sum_(S, 'frames')
            return engine.self_moments(S.unsqueeze(0), d.n_frames, d.lags[used], coll_mask)
        w = None if weights is None else self._dev(weights)
        rows = []
        for i in used:
            arr = d.device_entry(i) if isinstance(d, SingleOriginDisplacements) else engine.to_device(
                d[i], self._device, torch.float64)
            rows.append(engine.disp_moments(arr, self._mask, w, coll_mask)[2:])
        return torch.stack(rows)

    def _finish(self, used, main_sums, cnt, divisor, ngp_sums, gcnt):
        used = list(used)
        if len(used) == 0:
            self._stats_dev = torch.zeros((4, 0), dtype=torch.float64, device=self._device)
        else:
            packed = self._const(np.stack([np.asarray(cnt, dtype=np.float64), np.asarray(divisor, dtype=np.float64),
                                           np.asarray(gcnt, dtype=np.float64)]))  # one upload: counts, divisors, NGP counts
            join, self._moments_join = getattr(self, '_moments_join', None), None
            self._stats_dev = engine.moment_stats(main_sums, packed[0], packed[1], ngp_sums, packed[2], join=join)
            if join is not None:
                dist.check_peer_status()
        self._dt = np.asarray(self._delta_t, dtype=float)[np.asarray(used, dtype=np.int64)] if len(used) else np.array([])
        self._n_o = self._n_o[:len(used)]  # diffusion.py:602
        self._euclidian_displacements = _LazyEuclidian(self._displacements, used, self._slice, self.dims)

    def _host_stats(self):
        if self._stats_host is None:
            self._stats_host = self._stats_dev.cpu().numpy()
        return self._stats_host

    # ---- the reference's attribute surface ---------------------------------------------------
    @property
    def _n(self):
        return self._host_stats()[0]

    @property
    def _v(self):
        return self._host_stats()[1]

    @property
    def _s(self):
        return self._host_stats()[2]

    @property
    def _ngp(self):
        return self._host_stats()[3]

    @property
    def dt(self) -> np.ndarray:
        """:return: Timestep values that were sampled."""
        return self._dt

    @property
    def n(self) -> np.ndarray:
        """:return: The mean MSD/MSTD/MSCD, in units Å^2."""
        return self._n

    @property
    def s(self) -> np.ndarray:
        """:return: The MSD/MSTD/MSCD standard deviation, in units Å^2."""
        return self._s

    @property
    def v(self) -> np.ndarray:
        """:return: The MSD/MSTD/MSCD variance, in units Å^4."""
        return self._v

    @property
    def euclidian_displacements(self):
        """:return: Displacements between particles at each dt."""
        return self._euclidian_displacements

    @property
    def ngp(self) -> np.ndarray:
        """:return: Non-Gaussian parameter as a function of dt."""
        return self._ngp

    # ---- results of the regression: copied to the host asynchronously, materialised on first access -----------
    def _materialise(self):
        if self._flat_pending is not None:
            pinned, event, shape, fit_intercept = self._flat_pending
            self._flat_pending = None
            event.synchronize()
            n = int(np.prod(shape))
            self._flatchain = pinned[:n * 8].view(torch.float64).reshape(shape).numpy().copy()
            engine.pinned_give(pinned)
            self._gradient = Distribution(self._flatchain[:, 0])
            self._intercept_dist = Distribution(self._flatchain[:, 1]) if fit_intercept else None

    def __del__(self):
        # results that were never read: the buffer goes back to the pool once its copy has landed (waiting here also
        # keeps a caller that drops its results from queueing more than one analysis ahead of the device)
        pending = getattr(self, '_flat_pending', None)
        if pending is not None:
            try:
                pending[1].synchronize()
                engine.pinned_give(pending[0])
            except Exception:
                pass

    @property
    def flatchain(self):
        """:return: The flat MCMC chain `[(n_samples / thin) * n_walkers, ndim]` (None before the regression)."""
        self._materialise()
        return self._flatchain

    @flatchain.setter
    def flatchain(self, value):
        self._flat_pending = None
        self._flatchain = value

    @property
    def gradient(self) -> Union[Distribution, None]:
        """:return: The samples of the gradient of the straight line."""
        self._materialise()
        return self._gradient

    @gradient.setter
    def gradient(self, value):
        self._gradient = value

    @property
    def _intercept(self):
        self._materialise()
        return self._intercept_dist

    @_intercept.setter
    def _intercept(self, value):
        self._intercept_dist = value

    @property
    def intercept(self) -> Union[Distribution, None]:
        """:return: The estimated intercept (None if fit_intercept was False)."""
        return self._intercept

    def _converted(self, attribute):
        """D / D_J / sigma: the gradient samples times the unit conversion of the last regression call."""
        if self._coefficient is not None and self._coefficient[0] == attribute:
            g = self.gradient
            self._coefficient_cache = Distribution(g.samples * self._coefficient[1])
            setattr(self, attribute, self._coefficient_cache)
            self._coefficient = None
        return getattr(self, attribute)

    def _covariance_device(self):
        """The reconditioned covariance on the device; on the default path the sampler does not need the
        matrix itself (only its spectrum), so it is rebuilt here on first use."""
        if self._covariance_matrix_dev is None and self._cov_regime is not None:
            self._covariance_matrix_dev = self.generate_covariance_matrix(self._cov_regime)
        host = getattr(self, '_covariance_matrix_host', None)
        if self._covariance_matrix_dev is None and host is not None:  # an object restored by from_dict
            self._covariance_matrix_dev = engine.to_device(np.asarray(host), self._device, torch.float64)
        return self._covariance_matrix_dev

    @property
    def covariance_matrix(self) -> np.ndarray:
        """:return: The covariance matrix for the trajectories."""
        host = getattr(self, '_covariance_matrix_host', None)
        if host is not None and self._covariance_matrix_dev is None:
            return host
        cov = self._covariance_device()
        return None if cov is None else cov.cpu().numpy()

    @property
    def _covariance_matrix(self):
        return self.covariance_matrix

    @property
    def _npd_covariance_matrix(self):
        return None if self._npd_covariance_matrix_dev is None else self._npd_covariance_matrix_dev.cpu().numpy()

    @staticmethod
    def ngp_calculation(d_squared: np.ndarray) -> float:
        """
        Non-Gaussian parameter (kinisi/diffusion.py:291-303) of an array of squared displacements, through
        the same moment kernels: 3 <d^4> / (5 <d^2>^2) - 1.
        """
        d = np.sqrt(np.asarray(d_squared, dtype=np.float64).reshape(-1, 1, 1))
        dev = torch.device('cuda', torch.cuda.current_device())
        arr = engine.to_device(np.concatenate([d, np.zeros_like(d), np.zeros_like(d)], axis=2), dev)
        sums = engine.disp_moments(arr, 7)[:2].reshape(1, 2)
        cnt = torch.full((1, ), float(d.shape[0]), dtype=torch.float64, device=dev)
        one = torch.ones(1, dtype=torch.float64, device=dev)
        return float(engine.moment_stats(sums, cnt + 1.0, one, sums, cnt)[3, 0].item())

    # ---- stage 3 -----------------------------------------------------------------------------
    def generate_covariance_matrix(self, diff_regime: int):
        """
        The model covariance, reconditioned and repaired (kinisi/diffusion.py:397-425), on the device:
        `_populate_covariance_matrix` (:838-854) -> `minimum_eigenvalue_method` (:870-888) ->
        `cov_nearest` (statsmodels, clipped-correlation repair).
        """
        v = self._variances_for_covariance(diff_regime)
        n_o = self._dev(self._n_o[diff_regime:])
        cov = engine.cov_build(v, n_o)
        self._npd_covariance_matrix_dev = cov
        # minimum eigenvalue method: clip the spectrum at lambda_max / cond_max and rebuild
        lam, vec = torch.linalg.eigh(cov)
        floor = lam.max() / self._cond_max
        lam = torch.clamp(lam, min=floor)
        cov = (vec * lam) @ vec.T
        if not LITERAL_RECONDITIONING:
            # The matrix just rebuilt has the eigen-pairs (lam, vec) by construction, and cov_nearest
            # moves it by <= ~2e-15 relative (it lifts correlation eigenvalues from ~1e-16 to 1e-15, far
            # below the pinvh cut-off of n*eps).  The fast path therefore reuses this one factorisation
            # for the pseudo-inverse and the log-determinant.  tests/test_gpu_parity.py checks it against
            # the literal sequence below.
            self._eig = (lam, vec)
            return cov
        self._eig = None
        # cov_nearest(method='clipped', threshold=1e-15)
        std = torch.sqrt(torch.diagonal(cov))
        corr = cov / torch.outer(std, std)
        ev, evec = torch.linalg.eigh(corr)
        repaired = (evec * torch.clamp(ev, min=1e-15)) @ evec.T
        rstd = torch.sqrt(torch.diagonal(repaired))
        repaired = repaired / rstd / rstd[:, None]
        corr = torch.where((ev < 1e-15).any(), repaired, corr)
        return corr * torch.outer(std, std)

    def _variances_for_covariance(self, diff_regime: int):
        """The variances that enter the covariance model: the measured ones, or with `model=True` the fitted
        `a / n_o * dt**2` (kinisi/diffusion.py:406-420).  The reference fits `a` with `scipy.optimize.curve_fit` under
        the bound a >= 0; the model is linear in `a`, so the least-squares solution is closed form,
        a = max(0, sum(v f) / sum(f f)) with f = dt**2 / n_o, evaluated on the device."""
        v = self._stats_dev[1, diff_regime:].contiguous()
        if not self._model:
            self._model_v_dev = v
            return v
        f = self._dev(np.asarray(self._dt[diff_regime:], dtype=np.float64)**2 /
                      np.asarray(self._n_o[diff_regime:], dtype=np.float64))
        a = torch.clamp((v * f).sum() / (f * f).sum(), min=0.0)
        self._popt_dev = a
        self._model_v_dev = (a * f).contiguous()
        return self._model_v_dev

    def bootstrap_GLS(self,
                      start_dt: float,
                      cond_max: float = 1e16,
                      model: bool = False,
                      fit_intercept: bool = True,
                      n_samples: int = 1000,
                      n_walkers: int = 32,
                      n_burn: int = 500,
                      thin: int = 10,
                      progress: bool = True,
                      random_state: np.random.mtrand.RandomState = None):
        """
        Generalised-least-squares estimate of gradient and intercept under the model covariance, sampled
        with an affine-invariant ensemble sampler (kinisi/diffusion.py:305-395).  Same parameters and
        defaults as the reference.  Everything from the covariance build to the last MCMC step is
        enqueued on the device without a host synchronisation; the flat chain is copied back at the end.
        """
        if random_state is not None:
            np.random.seed(random_state.get_state()[1][1])
        if n_walkers % 2 or n_walkers < 2:
            raise ValueError('n_walkers must be an even number >= 2')

        self._start_dt = start_dt
        self._model = model
        self._cond_max = cond_max

        diff_regime = np.argwhere(self._dt >= self._start_dt)[0][0]

        xn = self._const(np.stack([np.asarray(self._dt[diff_regime:], dtype=np.float64),
                                   np.asarray(self._n_o[diff_regime:], dtype=np.float64)]))  # one upload: dt and n_o
        x, n_o_dev = xn[0], xn[1]
        y = self._stats_dev[0, diff_regime:].contiguous()
        n_dt = x.numel()
        self._cov_regime = diff_regime
        self._covariance_matrix_dev = None
        if not LITERAL_RECONDITIONING and n_dt <= engine.spectral_prepare_max_n():
            # One kernel: Jacobi spectrum of the model covariance -> clipped spectrum, pinvh cut-off, whitened
            # model, log-determinant and the GLS expansion (csrc/k4_eigen.cu).  The reconditioned matrix itself
            # is only materialised if `covariance_matrix` is read.
            v = self._variances_for_covariance(diff_regime)
            cov0 = engine.cov_build(v, n_o_dev)
            self._npd_covariance_matrix_dev = cov0
            self._evals, proj, theta_ml, logdet_term, self._sweeps = engine.spectral_prepare(
                cov0, x, y, cond_max, fit_intercept)
        else:
            cov = self.generate_covariance_matrix(diff_regime)
            self._covariance_matrix_dev = cov
            if self._eig is not None:
                lam, vec = self._eig
                logdet = torch.log(lam).sum()
            else:
                _, logdet = torch.linalg.slogdet(cov)
                lam, vec = torch.linalg.eigh(cov)  # pinvh is an eigendecomposition with a relative cut-off
            logdet_term = (logdet + float(np.log(2 * np.pi) * n_dt)).reshape(1)
            proj, theta_ml = engine.gls_prepare(lam, vec, x, y, fit_intercept)
        self._proj, self._logdet_term, self._theta_ml = proj, logdet_term, theta_ml

        ndim = 2 if fit_intercept else 1
        seed = int(np.random.randint(0, 2**31 - 1)) * 2654435761 + 12345
        chain, logp = engine.ensemble_sample(theta_ml, ndim, n_walkers, n_samples + n_burn, logdet_term, seed)
        self._chain_dev, self._logp_dev = chain, logp
        # emcee get_chain(flat=True, thin=thin, discard=n_burn)
        # The device->host copy of the flat chain goes to pinned memory without waiting for it: the host is free to
        # queue the next analysis (its transfer overlaps this regression); `flatchain`, `gradient`, `intercept`, `D`,
        # `D_J`, `sigma` wait for the copy when they are first read.
        flat = chain[n_burn + thin - 1::thin].reshape(-1, ndim)
        pinned = engine.pinned_take(flat.numel() * 8)
        pinned[:flat.numel() * 8].view(torch.float64).reshape(flat.shape).copy_(flat, non_blocking=True)
        event = torch.cuda.Event()
        event.record(engine.current_stream(self._device))
        self._flatchain = self._gradient = self._intercept_dist = None
        self._flat_pending = (pinned, event, tuple(flat.shape), bool(fit_intercept))

    def log_likelihood(self, theta) -> np.ndarray:
        """Batched log-likelihood (kinisi/diffusion.py:354-365) of parameter vectors `[n, ndim]`."""
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        return engine.mvn_loglike(self._dev(theta), self._proj, self._logdet_term).cpu().numpy()

    def diffusion(self, start_dt: float, **kwargs):
        """Diffusion coefficient in cm^2/s (kinisi/diffusion.py:427-436)."""
        self.bootstrap_GLS(start_dt, **kwargs)
        self._diffusion_coefficient = None
        self._coefficient = ('_diffusion_coefficient', 1.0 / (2e4 * self.dims))

    @property
    def D(self) -> Union[Distribution, None]:
        """:return: Diffusion coefficient, with units of cm^2 s^-1."""
        return self._converted('_diffusion_coefficient')

    def jump_diffusion(self, start_dt: float, **kwargs):
        """Jump diffusion coefficient in cm^2/s (kinisi/diffusion.py:447-457)."""
        self.bootstrap_GLS(start_dt, **kwargs)
        self._jump_diffusion_coefficient = None
        self._coefficient = ('_jump_diffusion_coefficient', 1.0 / (2e4 * self.dims * self._n_atoms()))

    @property
    def D_J(self) -> Union[Distribution, None]:
        """:return: Jump diffusion coefficient, with units of cm^2 s^-1."""
        return self._converted('_jump_diffusion_coefficient')

    def conductivity(self, start_dt: float, temperature: float, volume: float, **kwargs):
        """Ionic conductivity in mS/cm (kinisi/diffusion.py:468-482)."""
        self.bootstrap_GLS(start_dt, **kwargs)
        volume = volume * 1e-24  # cm^3
        conversion = 1000 / (volume * const.N_A) * (const.N_A * const.e)**2 / (const.R * temperature)
        self._sigma = None
        self._coefficient = ('_sigma', conversion / (2e4 * self.dims))  # D [cm^2 s^-1] = gradient / (2e4 dims)

    @property
    def sigma(self) -> Union[Distribution, None]:
        """:return: The estimated conductivity, with units mS cm^-1."""
        return self._converted('_sigma')

    def posterior_predictive(self, n_posterior_samples: int = None, n_predictive_samples: int = 256,
                             progress: bool = True) -> np.ndarray:
        """
        Sample the posterior predictive distribution (kinisi/diffusion.py:492-521); the result has the shape
        `(n_posterior_samples * n_predictive_samples, n_dt)`.  One eigen-factor of the covariance serves every
        posterior draw (the reference rebuilds a frozen multivariate normal per draw).
        """
        if n_posterior_samples is None:
            n_posterior_samples = self.gradient.size
        diff_regime = np.argwhere(self._dt >= self._start_dt)[0][0]
        x = self._dev(self._dt[diff_regime:])
        lam, vec = torch.linalg.eigh(self._covariance_device())
        factor = vec * torch.sqrt(torch.clamp(lam, min=0.0))  # cov = factor @ factor.T (allow_singular)
        draw = np.random.randint(0, self.gradient.size, size=n_posterior_samples)
        g = self._dev(self.gradient.samples[draw])
        c = self._dev(self.intercept.samples[draw])
        mu = g[:, None] * x[None, :] + c[:, None]
        gen = torch.Generator(device=self._device)
        gen.manual_seed(int(np.random.randint(0, 2**31 - 1)))
        z = torch.randn((n_posterior_samples, n_predictive_samples, x.numel()), dtype=torch.float64,
                        device=self._device, generator=gen)
        ppd = mu[:, None, :] + z @ factor.T
        return ppd.reshape(-1, x.numel()).cpu().numpy()

    def to_dict(self) -> dict:
        """:return: Dictionary description of the :py:class:`Bootstrap` (kinisi/diffusion.py:84-122): the statistics
        and the regression results as host arrays.  `displacements` is None: the displacement sequence stays on the
        device (the reference stores the whole O(N T n_dt) list)."""

        def dump(d):
            return None if d is None else d.to_dict()

        return {
            'displacements': None,
            'delta_t': self._delta_t,
            'n_o': np.asarray(self._n_o),
            'max_obs': self._max_obs,
            'dt': self._dt,
            'n': self._n,
            's': self._s,
            'v': self._v,
            'n_bootstrap': self._n_bootstrap,
            's_bootstrap': self._s_bootstrap,
            'v_bootstrap': self._v_bootstrap,
            'distributions': [d.to_dict() for d in self._distributions] if len(self._distributions) else None,
            'euclidian_displacements': self._euclidian_displacements,  # the lazy sequence itself, not its O(N T) values
            'sub_sample_dt': self._sub_sample_dt,
            'dimension': self._dimension,
            'ngp': self._ngp,
            'covariance_matrix': self.covariance_matrix,
            'diffusion_coefficient': dump(self.D),
            'jump_diffusion_coefficient': dump(self.D_J),
            'sigma': dump(self.sigma),
            'intercept': dump(self.intercept),
            'gradient': dump(self.gradient),
            'flatchain': self.flatchain,
            'start_dt': self._start_dt,
            'model': self._model,
            'cond_max': self._cond_max
        }

    @classmethod
    def from_dict(cls, my_dict: dict) -> 'Bootstrap':
        """
        Generate a :py:class:`Bootstrap` object from a dictionary (kinisi/diffusion.py:124-176): a host-side object
        holding the stored statistics and regression results; nothing is recomputed and no device is needed.

        :param my_dict: The input dictionary.

        :return: New :py:class:`Bootstrap` object.
        """

        def restore(d):
            return None if d is None else Distribution.from_dict(d)

        boot = object.__new__(Bootstrap)
        dimension = my_dict['dimension']
        if isinstance(dimension, bytes):
            dimension = dimension.decode()
        boot._displacements = my_dict.get('displacements')
        boot._delta_t = np.asarray(my_dict['delta_t'])
        boot._lazy = False
        boot._max_obs = my_dict['max_obs']
        boot._n_o = my_dict['n_o']
        boot._sub_sample_dt = my_dict['sub_sample_dt']
        boot._dimension = dimension
        boot._slice = DIMENSIONALITY[dimension.lower()]
        boot._mask = engine.dim_mask(dimension)
        boot.dims = len(dimension.lower())
        boot._dt = my_dict['dt']
        boot._stats_dev = None
        boot._stats_host = np.stack([np.asarray(my_dict[k], dtype=np.float64) for k in ('n', 'v', 's', 'ngp')])
        boot._n_bootstrap = my_dict['n_bootstrap']
        boot._s_bootstrap = my_dict['s_bootstrap']
        boot._v_bootstrap = my_dict['v_bootstrap']
        boot._distributions = ([Distribution.from_dict(d) for d in my_dict['distributions']]
                               if my_dict.get('distributions') is not None else [])
        ed = my_dict.get('euclidian_displacements', [])
        boot._euclidian_displacements = ed if isinstance(ed, _LazyEuclidian) else [Distribution.from_dict(d) for d in ed]
        boot._flat_pending = None
        boot._flatchain = my_dict['flatchain']
        boot._gradient = restore(my_dict['gradient'])
        boot._intercept_dist = restore(my_dict['intercept'])
        boot._coefficient = None
        boot._diffusion_coefficient = restore(my_dict['diffusion_coefficient'])
        boot._jump_diffusion_coefficient = restore(my_dict['jump_diffusion_coefficient'])
        boot._sigma = restore(my_dict['sigma'])
        boot._covariance_matrix_dev = None
        boot._npd_covariance_matrix_dev = None
        boot._covariance_matrix_host = my_dict['covariance_matrix']
        boot._cov_regime = None
        boot._eig = None
        boot._start_dt = my_dict['start_dt']
        boot._model = my_dict['model']
        boot._cond_max = my_dict['cond_max']
        boot._resample, boot._block, boot._atom_sum_cache = None, False, None
        boot._device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
        return boot


class MSDBootstrap(Bootstrap):
    """
    Mean and uncertainty of the mean squared displacements (kinisi/diffusion.py:524-602).

    :param delta_t: An array of the time interval values, units of ps.
    :param disp_3d: The displacements (see :py:class:`Bootstrap`).
    :param n_o: Number of statistically independent observations of the MSD at each time interval.
    :param sub_sample_dt, dimension, progress: as in the reference.
    :param bootstrap, n_resamples, max_resamples, alpha, random_state: `bootstrap=True` adds the resampled
        distributions of the mean (`_distributions`, `_n_bootstrap`, `_v_bootstrap`, `_s_bootstrap`;
        kinisi/diffusion.py:578-583), see :py:meth:`Bootstrap.sample_until_normal`.
    :param block: `block=True` replaces the mean and variance by the blocking-method estimates
        (kinisi/diffusion.py:584-595), see :py:meth:`Bootstrap._blocked`.
    """

    def __init__(self,
                 delta_t: np.ndarray,
                 disp_3d,
                 n_o: np.ndarray,
                 sub_sample_dt: int = 1,
                 bootstrap: bool = False,
                 block: bool = False,
                 n_resamples: int = 1000,
                 max_resamples: int = 10000,
                 dimension: str = 'xyz',
                 alpha: float = 1e-3,
                 random_state: np.random.mtrand.RandomState = None,
                 progress: bool = True):
        super().__init__(delta_t, disp_3d, n_o, sub_sample_dt, dimension)
        self._resampling(bootstrap, block, n_resamples, max_resamples, alpha, random_state)
        n_atoms = self._n_atoms()
        if self._lazy:
            sizes = n_atoms * self._displacements.n_obs_all()
        else:
            sizes = np.array([self._shape_of(i)[0] * self._shape_of(i)[1] for i in range(len(self._displacements))],
                             dtype=np.int64)
        used = np.flatnonzero(sizes > 1)  # diffusion.py:575
        cnt = sizes[used].astype(np.float64)
        div = np.asarray(n_o, dtype=np.float64)[used]  # diffusion.py:598: n_o is NOT strided by sub_sample_dt
        used = used.tolist()
        if used:
            sums = self._self_sums(np.array(used), self._mask)
            self._finish(used, sums, cnt, div, sums, cnt)
            if self._resample is not None or self._block:
                pops = (self._population(i, self._mask) for i in used)
                if self._resample is not None:
                    pops = [self._resample_entry(v, n_o[i]) for i, v in zip(used, pops)]  # diffusion.py:579
                if self._block:
                    self._apply_blocking(pops)
        else:
            self._finish(used, None, cnt, div, None, cnt)


class _CollectiveBootstrap(Bootstrap):
    """Shared body of MSTDBootstrap (kinisi/diffusion.py:636-688) and MSCDBootstrap (:723-780)."""

    def _run(self, n_o, weights, coll_mask):
        n_atoms = self._n_atoms()
        used, m, div, gcnt = [], [], [], []
        for i in range(int(len(self._displacements) / 2)):  # diffusion.py:650, 738: first half of the grid
            shp = self._shape_of(i)
            if shp[1] <= 1:  # coll_motion.size <= 1
                continue
            n_local = shp[0] if not self._lazy else n_atoms
            used.append(i)
            m.append(float(shp[1]))
            div.append(float(n_o[i]) / n_local)  # diffusion.py:684, 776
            gcnt.append(float(n_local * shp[1]))
        if used:
            u = np.array(used)
            # a streamed sequence rebuilds its displacements for every pass over the atoms: the per-atom moments and the
            # (charge-weighted) atom sum are taken from one pass
            self._coll_request = None
            if self._lazy and getattr(self._displacements, 'streamed', False):
                self._coll_request = {'weights': None if weights is None else self._dev(weights), 'S': None}
            self_sums = self._self_sums(u, self._mask)
            coll = self._collective_sums(u, weights, coll_mask)
            self._coll_request = None
            self._finish(used, coll, m, div, self_sums, gcnt)
            if self._resample is not None or self._block:
                pops = (self._population(i, coll_mask, weights, collective=True) for i in used)
                if self._resample is not None:  # diffusion.py:664, 756: n_samples = n_o[i] / d_squared.shape[0]
                    pops = [self._resample_entry(v, n_samples) for v, n_samples in zip(pops, div)]
                if self._block:
                    self._apply_blocking(pops)
        else:
            self._finish(used, None, m, div, None, gcnt)


class MSTDBootstrap(_CollectiveBootstrap):
    """
    Mean and uncertainty of the total (collective) mean squared displacements
    (kinisi/diffusion.py:605-688).  Only the first half of the time intervals is used (:650).
    """

    def __init__(self,
                 delta_t: np.ndarray,
                 disp_3d,
                 n_o: np.ndarray,
                 sub_sample_dt: int = 1,
                 bootstrap: bool = False,
                 block: bool = False,
                 n_resamples: int = 1000,
                 max_resamples: int = 10000,
                 dimension: str = 'xyz',
                 alpha: float = 1e-3,
                 random_state: np.random.mtrand.RandomState = None,
                 progress: bool = True):
        super().__init__(delta_t, disp_3d, n_o, sub_sample_dt, dimension)
        self._resampling(bootstrap, block, n_resamples, max_resamples, alpha, random_state)
        self._run(n_o, None, self._mask)


class MSCDBootstrap(_CollectiveBootstrap):
    """
    Mean and uncertainty of the mean squared charge displacements (kinisi/diffusion.py:691-780).

    :param ionic_charge: The charge on the mobile ions, either an array with a value for each ion or a
        scalar if all values are the same.
    The charged sum uses all three components whatever `dimension` says, as the reference does (:751).
    """

    def __init__(self,
                 delta_t: np.ndarray,
                 disp_3d,
                 ionic_charge: Union[np.ndarray, int],
                 n_o: np.ndarray,
                 sub_sample_dt: int = 1,
                 bootstrap: bool = False,
                 block: bool = False,
                 n_resamples: int = 1000,
                 max_resamples: int = 10000,
                 dimension: str = 'xyz',
                 alpha: float = 1e-3,
                 random_state: np.random.mtrand.RandomState = None,
                 progress: bool = True):
        super().__init__(delta_t, disp_3d, n_o, sub_sample_dt, dimension)
        self._resampling(bootstrap, block, n_resamples, max_resamples, alpha, random_state)
        try:
            _ = len(ionic_charge)
            ionic_charge = np.asarray(ionic_charge, dtype=np.float64)
        except TypeError:
            n_local = self._displacements.n_atoms if self._lazy else self._shape_of(0)[0]
            ionic_charge = np.ones(n_local) * ionic_charge
        self._run(n_o, ionic_charge, 7)
