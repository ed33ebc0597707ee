"""Host-side logic that needs no GPU: the Distribution container, shard bounds, the lag-grid / n_o
bookkeeping of the parser mirror (checked against the oracle, which is pinned to the reference)."""
import os

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


def test_lag_grid_is_split_into_arithmetic_runs():
    """The host-side planner of kb2_self_moments_runs (no GPU needed): every lag lands in exactly one run, and
    kinisi's default linspace grids come out as a handful of long runs (with one-sample drift where the integer
    truncation of np.linspace shifts the progression)."""
    import numpy as np
    from kinisi_b200 import engine

    def check(lags, n_frames):
        lags = np.asarray(lags)
        runs, rest = engine.self_moments_plan(lags, n_frames)
        for k0, d, n, n_chg in runs:
            assert n >= 3 and d >= 1 and abs(n_chg) <= n // 8 + 2
            assert k0 in lags
            last = k0 + (n - 1) * d + n_chg
            assert last in lags, (k0, d, n, n_chg)
        assert sum(r[2] for r in runs) + rest == len(lags)
        return runs, rest

    runs, rest = check(np.linspace(1, 20000, 100, dtype=int), 20000)  # BASELINE configs[1]
    assert rest == 1 and runs == [(1, 202, 99, 0)]  # long progressions are not chained
    runs, rest = check(np.linspace(1, 5000, 100, dtype=int), 5000)  # configs[0]: alternating 50/51 spacing
    assert rest <= 2 and len(runs) <= 4
    runs, rest = check(np.linspace(1, 100000, 100, dtype=int), 100000)  # configs[2]: +1 every 11th lag
    assert rest == 0 and runs == [(1, 1010, 100, 9)]
    runs, rest = check(np.linspace(1, 50000, 100, dtype=int), 50000)  # configs[4]
    assert rest == 1 and len(runs) == 4 and all(r[1] == 505 for r in runs)
    runs, rest = check(np.linspace(1, 20000, 2000, dtype=int)[:192], 20000)  # 2000-point grid, one 192-lag launch
    assert rest == 0 and len(runs) <= 4
    runs, rest = check(np.unique(np.geomspace(1, 20000, 100, dtype=int)), 20000)
    assert runs[0][1] == 1  # the dense start 1, 2, 3, ...
    runs, rest = check([5], 100)
    assert runs == [] and rest == 1
    runs, rest = check([2, 3, 5, 7, 11, 13, 17, 19, 23], 100)
    assert rest + sum(r[2] for r in runs) == 9


def test_bootstrap_from_dict_needs_no_device():
    """Bootstrap.from_dict (kinisi/diffusion.py:124-176) restores the stored statistics and regression results on the host."""
    from kinisi_b200 import diffusion
    from kinisi_b200.distribution import Distribution
    rng = np.random.default_rng(0)
    grad = Distribution(rng.normal(3.0, 0.1, size=640))
    d = {
        'displacements': None, 'delta_t': np.arange(1.0, 6.0), 'n_o': np.arange(5.0, 0.0, -1.0), 'max_obs': 40,
        'dt': np.arange(1.0, 6.0), 'n': np.arange(5.0), 's': np.ones(5), 'v': np.ones(5), 'ngp': np.array([0.1, 0.5, 0.2, 0.0, -0.1]),
        'n_bootstrap': np.array([]), 's_bootstrap': np.array([]), 'v_bootstrap': np.array([]), 'distributions': None,
        'sub_sample_dt': 1, 'dimension': 'xy', 'covariance_matrix': np.eye(3), 'diffusion_coefficient': grad.to_dict(),
        'jump_diffusion_coefficient': None, 'sigma': None, 'intercept': None, 'gradient': grad.to_dict(),
        'flatchain': grad.samples[:, None], 'start_dt': 3.0, 'model': True, 'cond_max': 1e16, 'euclidian_displacements': []
    }
    b = diffusion.Bootstrap.from_dict(d)
    assert b.dims == 2 and b.n[3] == 3.0 and b.ngp.argmax() == 1 and b.intercept is None
    np.testing.assert_array_equal(b.D.samples, grad.samples)
    np.testing.assert_array_equal(b.covariance_matrix, np.eye(3))
    again = b.to_dict()
    np.testing.assert_array_equal(again['gradient']['samples'], grad.samples)
    assert again['start_dt'] == 3.0 and again['sigma'] is None


def _write_xdatcar(path, species, counts, frac, latt, repeat_header):
    """frac [F, N, 3], latt [F, 3, 3] -> XDATCAR text (one header, or one per configuration as NPT runs write)."""
    import gzip
    out = []
    for f in range(frac.shape[0]):
        if f == 0 or repeat_header:
            out += ['synthetic', '   1.000000'] + ['  '.join(f'{x:.9f}' for x in row) for row in latt[f]]
            out += ['  '.join(species), '  '.join(str(c) for c in counts)]
        out.append(f'Direct configuration= {f + 1:5d}')
        out += ['  '.join(f'{x:.9f}' for x in row) for row in frac[f]]
    text = '\n'.join(out) + '\n'
    opener = gzip.open if str(path).endswith('.gz') else open
    with opener(path, 'wt') as fh:
        fh.write(text)


def test_xdatcar_reader(tmp_path):
    """kinisi_b200/xdatcar.py: both XDATCAR layouts, gzip, and the surface PymatgenParser reads."""
    from kinisi_b200.xdatcar import Xdatcar
    rng = np.random.default_rng(0)
    frac = np.round(rng.random((6, 5, 3)), 9)
    latt = np.round(np.eye(3)[None] * 10.0 + 0.01 * rng.normal(size=(6, 3, 3)), 9)
    fixed = np.repeat(latt[:1], 6, axis=0)
    for name, cells, repeat in (('XDATCAR', fixed, False), ('XDATCAR_npt', latt, True), ('XDATCAR.gz', fixed, False)):
        path = tmp_path / name
        _write_xdatcar(path, ['Li', 'O'], [2, 3], frac, cells, repeat)
        x = Xdatcar(str(path))
        assert len(x.structures) == 6 and x.site_symbols == ['Li', 'Li', 'O', 'O', 'O']
        np.testing.assert_allclose(x.frac_coords, frac, atol=1e-12)
        np.testing.assert_allclose(x.lattices, cells, atol=1e-12)
        s0 = x.structures[3]
        assert [str(site.specie) for site in s0] == x.site_symbols and len(s0) == 5
        np.testing.assert_allclose(s0.volume, abs(np.linalg.det(cells[3])), rtol=1e-12)
        np.testing.assert_allclose(s0.lattice.matrix, cells[3], atol=1e-12)
    ref = '/root/reference/kinisi/tests/inputs/example_XDATCAR.gz'
    if os.path.exists(ref):  # the reference's own fixture (kinisi/tests/test_analyze.py:92-98), in the build container only
        x = Xdatcar(ref)
        assert x.frac_coords.shape == (140, 416, 3) and x.site_symbols.count('Li') == 192


def test_time_interval_grid_is_kept_by_value():
    """get_time_intervals keeps the grid of a set of arguments (a loop over trajectories of one shape asks for the same one
    every time): the caller gets a copy, and writing to it does not reach the next caller."""
    from kinisi_b200.parser import Parser
    p = Parser.__new__(Parser)
    p.time_step, p.step_skip = 0.5, 4
    p.min_dt, p.max_dt = 2.0, 5000 * 2.0
    first = p.get_time_intervals(100, 'linear')
    np.testing.assert_array_equal(first, np.linspace(1, 5000, 100, dtype=int))
    first[:] = -1
    again = p.get_time_intervals(100, 'linear')
    np.testing.assert_array_equal(again, np.linspace(1, 5000, 100, dtype=int))
    with pytest.raises(ValueError, match='Only linear or logarithmic'):
        p.get_time_intervals(100, 'cubic')


def test_peer_plumbing_is_inert_without_a_group():
    """Outside a process group nothing asks for peer-memory windows: no channel, no fused join."""
    import torch
    from kinisi_b200 import dist
    assert dist.frame_sums_channel(3, 20000, torch.device('cpu')) is None
    assert dist.moments_channel(200, torch.device('cpu')) is None
    assert not dist._PEER
    dist.check_peer_status()  # no status word has been raised


def test_direct_nccl_plumbing_is_inert_without_a_group():
    """dist.allreduce_sum_ is a no-op outside a process group, and the direct NCCL communicator is only considered for
    contiguous float64 CUDA tensors (everything else stays on torch.distributed)."""
    import ctypes
    import torch
    from kinisi_b200 import dist
    assert ctypes.sizeof(dist._NcclUniqueId) == 128  # ncclUniqueId is 128 bytes (nccl.h)
    t = torch.arange(4, dtype=torch.float64)
    assert not dist.active() and dist.allreduce_sum_(t, 'moments') is t and t.tolist() == [0.0, 1.0, 2.0, 3.0]
    assert dist._direct('moments', t) is None  # a CPU tensor
    assert dist.world_size() == 1 and dist.rank() == 0
    assert dist.shard_bounds(10, 1, 3) == (4, 7)


def test_xdatcar_reader_cartesian_and_volume_scale(tmp_path):
    """Cartesian configurations are converted with the cell; a negative scale line is the cell volume (VASP convention)."""
    from kinisi_b200.xdatcar import Xdatcar
    cell = np.array([[4.0, 0.0, 0.0], [0.5, 5.0, 0.0], [0.0, 0.3, 6.0]])
    frac = np.array([[0.1, 0.2, 0.3], [0.6, 0.7, 0.8]])
    cart = frac @ cell
    lines = ['cart', '  1.0'] + ['  '.join(f'{x:.10f}' for x in r) for r in cell] + ['Na Cl', '1 1', 'Cartesian configuration=     1']
    lines += ['  '.join(f'{x:.10f}' for x in r) for r in cart]
    (tmp_path / 'X1').write_text('\n'.join(lines) + '\n')
    x = Xdatcar(str(tmp_path / 'X1'))
    np.testing.assert_allclose(x.frac_coords[0], frac, atol=1e-9)
    assert x.site_symbols == ['Na', 'Cl']
    unit = np.eye(3)
    lines = ['vol', '  -27.0'] + ['  '.join(f'{x:.10f}' for x in r) for r in unit] + ['Li', '1', 'Direct configuration=     1', '0.5 0.5 0.5']
    (tmp_path / 'X2').write_text('\n'.join(lines) + '\n')
    x = Xdatcar(str(tmp_path / 'X2'))
    np.testing.assert_allclose(x.lattices[0], 3.0 * np.eye(3), rtol=1e-12)
    np.testing.assert_allclose(x.structures[0].volume, 27.0, rtol=1e-12)
    (tmp_path / 'X3').write_text('\n'.join(['empty', '1.0', '1 0 0', '0 1 0', '0 0 1', 'Li', '1']) + '\n')
    with pytest.raises(ValueError, match='no configurations'):
        Xdatcar(str(tmp_path / 'X3'))


def test_universe_walk_matches_the_restatement():
    """MDAnalysisParser.get_structure_coords_latt (host arithmetic of kinisi/parser.py:600-638) on a duck-typed universe
    against the oracle's restatement; sub-sampling of atoms and frames; index selection by atom type."""
    from duck_atoms import universe_from_arrays
    from kinisi_b200.parser import MDAnalysisParser
    from oracle import kinisi_oracle as ko
    rng = np.random.default_rng(4)
    frac = rng.random((6, 9, 3))
    latt = np.eye(3)[None] * 8.0 + 0.05 * rng.normal(size=(9, 3, 3))
    types = ['Li', 'O', 'Li', 'O', 'O', 'Li']
    u = universe_from_arrays(types, frac, latt, np.float64)
    structure, coords, cells, volume = MDAnalysisParser.get_structure_coords_latt(u, progress=False)
    cart = np.einsum('nfk,fkj->fnj', frac, latt)
    want, want_cells = ko.universe_coords_latt(cart, latt)
    np.testing.assert_allclose(np.concatenate(coords, axis=1), want, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(np.array(cells), want_cells, rtol=0, atol=0)
    np.testing.assert_allclose(volume, abs(np.linalg.det(latt[0])), rtol=1e-12)
    assert len(coords) == 10 and coords[0].shape == (6, 1, 3)
    _, sub, sub_cells, _ = MDAnalysisParser.get_structure_coords_latt(u, sub_sample_atoms=2, sub_sample_traj=3, progress=False)
    assert len(sub) == 4 and sub[0].shape == (3, 1, 3)
    np.testing.assert_allclose(sub[1][:, 0], want[::2, 1 + 0], rtol=1e-13)
    np.testing.assert_allclose(sub[2][:, 0], want[::2, 1 + 3], rtol=1e-13)
    assert MDAnalysisParser.get_indices(structure, 'Li', None) == ([0, 2, 5], [1, 3, 4])
    assert MDAnalysisParser.get_indices(structure, ['O'], [0])[1] == [0]
    with pytest.raises(ValueError, match='no species'):
        MDAnalysisParser.get_indices(structure, 'Xe', None)
