"""Host-side logic that needs no GPU: the Distribution container, shard bounds, the lag-grid / n_o
bookkeeping of the parser mirror (checked against the oracle, which is pinned to the reference)."""
import numpy as np
import pytest

from kinisi_b200 import dist
from kinisi_b200.distribution import Distribution
from oracle import kinisi_oracle as ko


def test_distribution_surface():
    rng = np.random.RandomState(0)
    d = Distribution(rng.randn(1000) * 2 + 5)
    assert d.size == 1000 and len(d) == 1000
    assert d.n == pytest.approx(np.median(d.samples))
    lo, hi = d.con_int()
    assert lo < d.n < hi
    assert np.mean(d) == pytest.approx(d.samples.mean())
    d2 = Distribution.from_dict(d.to_dict())
    np.testing.assert_array_equal(d2.samples, d.samples)


def test_shard_bounds_partition_everything():
    for n in (0, 1, 7, 448, 100000):
        for w in (1, 2, 3, 8):
            cover = []
            for r in range(w):
                lo, hi = dist.shard_bounds(n, r, w)
                cover.extend(range(lo, hi))
            assert cover == list(range(n))


class _FakeTraj:
    """Lets Parser.get_time_intervals / get_disps run without a device."""

    def __init__(self, n_frames):
        self.n_frames = n_frames


@pytest.mark.parametrize('spacing', ['linear', 'logarithmic'])
@pytest.mark.parametrize('n_frames,n_atoms,n_steps', [(100, 100, 100), (5000, 192, 100), (37, 1, 20), (12, 3, 50)])
def test_lag_grid_and_n_o_match_oracle(spacing, n_frames, n_atoms, n_steps):
    import torch
    from kinisi_b200.parser import Parser
    p = Parser.__new__(Parser)
    p.time_step, p.step_skip, p.sampling = 0.5, 4, 'multi-origin'
    unit = p.time_step * p.step_skip
    p.min_dt, p.max_dt = unit, n_frames * unit
    p._traj = _FakeTraj(n_frames)
    p._n_atoms_global = n_atoms
    if n_steps > (n_frames - 1) + 1:
        n_steps = n_frames
    lags = p.get_time_intervals(n_steps, spacing)
    ref = ko.time_intervals(n_frames, 0.5, 4, None, None, n_steps, spacing)
    np.testing.assert_array_equal(lags, ref)
    dc = torch.zeros((n_atoms, 3, 16 * ((n_frames + 15) // 16)), dtype=torch.float64)
    delta_t, disp_3d, n_o = p.get_disps(lags, dc, False)
    parsed = ko.parse(np.zeros((n_atoms, n_frames, 3)), np.arange(n_atoms), [], 0.5, 4, n_steps=n_steps, spacing=spacing)
    np.testing.assert_allclose(delta_t, parsed['delta_t'])
    np.testing.assert_allclose(n_o, parsed['n_o'])
    np.testing.assert_array_equal(disp_3d.lags, parsed['lags'])
    assert disp_3d.shapes == [(n_atoms, n_frames - k + 1, 3) for k in parsed['lags']]


def test_lazy_displacements_materialise_like_the_reference():
    import torch
    from kinisi_b200.parser import LazyDisplacements
    rng = np.random.RandomState(1)
    dc = rng.randn(4, 30, 3).cumsum(axis=1)
    t = torch.zeros((4, 3, 32), dtype=torch.float64)
    t[:, :, :30] = torch.from_numpy(dc).permute(0, 2, 1)
    lazy = LazyDisplacements(t, 30, [1, 7, 30])
    for i, k in enumerate([1, 7, 30]):
        np.testing.assert_array_equal(lazy[i], ko.displacements_for_lag(dc, k))
    assert len(lazy[::2]) == 2 and lazy[::2].lags.tolist() == [1, 30]
