"""Pins the CPU oracle (oracle/kinisi_oracle.py) against vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known-answer tests."""
import os

import warnings

import numpy as np
import pytest

from oracle import kinisi_oracle as ko
from synth_inputs import (ASE_CASES, CHARGES12, REGRESS_CASES, ase_case_inputs, list_bootstrap_inputs,
                          parser_random_disp, regress_case_inputs)

RTOL = 1e-12


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + '.npz'))


def ase_parse(symbols, frac, latt, pp):
    """ASEParser front-end semantics (kinisi/parser.py:260-356) expressed on arrays."""
    sub = pp.get('sub_sample_traj', 1)
    frac, latt = frac[:, ::sub], latt[::sub]
    frac = np.concatenate([frac[:, :1], frac], axis=1)
    latt = np.concatenate([latt[:1], latt], axis=0)
    idx = [i for i, s in enumerate(symbols) if s == pp['specie']]
    fw = [i for i, s in enumerate(symbols) if s != pp['specie']]
    if isinstance(pp.get('framework_indices'), (list, tuple)):
        fw = pp['framework_indices']
    dc = ko.get_disp(frac, latt)
    return ko.parse(dc, idx, fw, pp['time_step'], pp['step_skip'], pp.get('min_dt'), pp.get('max_dt'),
                    pp.get('n_steps', 100), pp.get('spacing', 'linear'))


@pytest.mark.parametrize('name', list(ASE_CASES))
def test_parser_and_moments(golden_dir, name):
    g = load(golden_dir, name)
    symbols, frac, latt, pp = ase_case_inputs(name)
    parsed = ase_parse(symbols, frac, latt, pp)
    np.testing.assert_allclose(parsed['dc'], g['dc'], rtol=RTOL, atol=1e-13)
    np.testing.assert_array_equal(parsed['lags_all'], g['time_intervals'])
    np.testing.assert_allclose(parsed['delta_t'], g['delta_t'], rtol=1e-15)
    np.testing.assert_allclose(parsed['n_o'], g['n_o'], rtol=1e-15)
    shapes = np.array([(parsed['dc_sel'].shape[0], parsed['dc_sel'].shape[1] - k + 1, 3) for k in parsed['lags']])
    np.testing.assert_array_equal(shapes, g['disp_shapes'])
    np.testing.assert_allclose(ko.displacements_for_lag(parsed['dc_sel'], parsed['lags'][0]), g['disp_first'], rtol=RTOL,
                               atol=1e-13)
    mid = len(parsed['lags']) // 2
    np.testing.assert_allclose(ko.displacements_for_lag(parsed['dc_sel'], parsed['lags'][mid]), g['disp_mid'], rtol=RTOL,
                               atol=1e-13)
    for dim in ('xyz', 'xy', 'z', 'xz'):
        for kind, q in (('msd', None), ('mstd', None), ('mscd', CHARGES12)):
            res = ko.bootstrap_streaming(parsed, kind, q, dimension=dim)
            for key in ('dt', 'n', 'v', 's', 'ngp', 'n_o'):
                np.testing.assert_allclose(res[key], g[f'{kind}_{dim}_{key}'], rtol=1e-11, err_msg=f'{kind} {dim} {key}')
    res = ko.bootstrap_streaming(parsed, 'mscd', 2)
    for key in ('dt', 'n', 'v', 's', 'ngp'):
        np.testing.assert_allclose(res[key], g[f'mscd_scalar2_{key}'], rtol=1e-11)
    res = ko.bootstrap_streaming(parsed, 'msd', sub_sample_dt=3)
    for key in ('dt', 'n', 'v', 's', 'ngp'):
        np.testing.assert_allclose(res[key], g[f'msd_sub3_{key}'], rtol=1e-11)


def test_parser_random(golden_dir):
    g = load(golden_dir, 'parser_random')
    disp = parser_random_disp()
    p = ko.parse(disp, np.arange(100), [], 1, 1)
    assert p['delta_t'].size == 100 and g['lin_delta_t'].size == 100
    np.testing.assert_allclose(p['delta_t'], g['lin_delta_t'])
    np.testing.assert_allclose(p['n_o'], g['lin_n_o'])
    res = ko.bootstrap_streaming(p, 'msd')
    for key in ('dt', 'n', 'v', 's', 'ngp'):
        np.testing.assert_allclose(res[key], g[f'lin_msd_{key}'], rtol=1e-11)
    p = ko.parse(disp, np.arange(0, 100, 2), np.arange(1, 100, 2), 0.5, 4, 10., 150., 30, 'logarithmic')
    np.testing.assert_array_equal(p['lags_all'], g['log_intervals'])
    np.testing.assert_allclose(p['dc'], g['log_dc'], rtol=RTOL)
    np.testing.assert_allclose(p['n_o'], g['log_n_o'])
    res = ko.bootstrap_streaming(p, 'msd')
    for key in ('dt', 'n', 'v', 's', 'ngp'):
        np.testing.assert_allclose(res[key], g[f'log_msd_{key}'], rtol=1e-11)
    # N = 1: tail lags dropped by the parser
    p = ko.parse(disp[:1], np.arange(1), [], 1, 1, n_steps=20)
    np.testing.assert_allclose(p['delta_t'], g['one_delta_t'])
    np.testing.assert_allclose(p['n_o'], g['one_n_o'])
    assert len(p['lags']) == g['one_shapes'].shape[0]
    res = ko.bootstrap_streaming(p, 'msd')
    for key in ('dt', 'n', 'v', 's', 'ngp', 'n_o'):
        np.testing.assert_allclose(res[key], g[f'one_msd_{key}'], rtol=1e-11)


def test_list_bootstrap(golden_dir):
    g = load(golden_dir, 'list_bootstrap')
    dt, disp_3d, n_o, disp_skip = list_bootstrap_inputs()
    for kind, q in (('msd', None), ('mstd', None), ('mscd', 1)):
        res = ko.bootstrap_from_list(dt, disp_3d, n_o, kind, q)
        for key in ('dt', 'n', 'v', 's', 'ngp', 'n_o'):
            np.testing.assert_allclose(res[key], g[f'{kind}_{key}'], rtol=1e-12)
    assert g['mstd_n'].shape == (5, )  # half-grid rule, kinisi/tests/test_diffusion.py:402
    res = ko.bootstrap_from_list(np.linspace(100, 400, 4), disp_skip, np.ones(4), 'msd')
    for key in ('dt', 'n', 'v', 's', 'ngp', 'n_o'):
        np.testing.assert_allclose(res[key], g[f'skip_msd_{key}'], rtol=1e-12)


@pytest.mark.parametrize('name', list(REGRESS_CASES))
def test_regression_stage(golden_dir, name):
    g = load(golden_dir, name)
    symbols, frac, latt, pp, start_idx, kind = regress_case_inputs(name)
    parsed = ase_parse(symbols, frac, latt, pp)
    q = 1 if kind == 'mscd' else None
    stats = ko.bootstrap_streaming(parsed, kind, q)
    for key in ('dt', 'n', 'v', 's', 'ngp', 'n_o'):
        np.testing.assert_allclose(stats[key], g[key], rtol=1e-11)
    diff_regime = np.argwhere(stats['dt'] >= g['start_dt'])[0][0]
    npd = ko.populate_covariance_matrix(stats['v'][diff_regime:], stats['n_o'][diff_regime:])
    np.testing.assert_allclose(npd, g['npd_covariance_matrix'], rtol=1e-11)
    cov = ko.generate_covariance_matrix(stats['v'], stats['n_o'], diff_regime)
    scale = np.abs(g['covariance_matrix']).max()
    np.testing.assert_allclose(cov, g['covariance_matrix'], rtol=1e-6, atol=1e-9 * scale)
    ll_fn, _, _ = ko.make_log_likelihood(stats['dt'][diff_regime:], stats['n'][diff_regime:], g['covariance_matrix'])
    ll = np.array([ll_fn(t) for t in g['thetas']])
    assert ll[-1] == -np.inf and g['loglike'][-1] == -np.inf
    np.testing.assert_allclose(ll[:-1], g['loglike'][:-1], rtol=1e-9)


def test_reference_kats(golden_dir):
    g = load(golden_dir, 'reference_kats')
    assert ko.ngp_calculation(np.array([1, 2, 3])) == pytest.approx(-0.3)  # test_diffusion.py:105-107
    assert float(g['ngp_123']) == pytest.approx(-0.3)
    # drift KAT, test_parser.py:151-168
    frac, latt = g['drift_frac'], g['drift_latt']
    frac = np.concatenate([frac[:, :1], frac], axis=1)
    latt = np.concatenate([latt[:1], latt], axis=0)
    dc = ko.get_disp(frac, latt)
    corrected = ko.correct_drift(list(range(1, 9)), dc)
    np.testing.assert_allclose(corrected[0], [[0, 0, 0], [0.2, 0.2, 0.2], [0.4, 0.4, 0.4], [1, 1, 1]], atol=1e-12)
    np.testing.assert_allclose(ko.correct_drift([], dc)[0], g['nodrift_dc0'], atol=1e-12)
    # docs fixture
    r = ko.moments_from_disp(g['docs_disp'], 384.0)
    assert r[0] == pytest.approx(5.1108279, rel=1e-7)
    assert r[1] == pytest.approx(0.0911415, rel=1e-6)
    assert r[2] == pytest.approx(0.4038798, rel=1e-6)
    assert r[0] == pytest.approx(float(g['docs_msd']), rel=1e-14)
    # real XDATCAR (Li6PS5Cl-like, 416 atoms x 140 frames), test_analyze.py:92-98
    symbols = list(g['lips_symbols'])
    parsed = ase_parse(symbols, g['lips_frac'].astype(np.float64), np.tile(g['lips_latt'], (140, 1, 1)),
                       dict(specie='Li', time_step=2.0, step_skip=50))
    assert parsed['dc_sel'].shape == (192, 140, 3)
    np.testing.assert_allclose(parsed['delta_t'], g['lips_delta_t'])
    res = ko.bootstrap_streaming(parsed, 'msd')
    for key in ('dt', 'n', 'v', 's', 'ngp', 'n_o'):
        np.testing.assert_allclose(res[key], g[f'lips_msd_{key}'], rtol=1e-11)


def test_molecule_centres_reference_kat():
    """kinisi/tests/test_parser.py:127-149 (example_center.XDATCAR, first frame): centre of geometry and centre of mass
    (masses 1, 16, 1) of the three atoms."""
    frame0 = np.array([[0.2, 0.2, 0.2], [0.4, 0.2, 0.2], [0.22, 0.4, 0.2]])
    coords = frame0[:, None, :]
    cog = ko.molecule_centres(coords, [[0, 1, 2]])
    np.testing.assert_array_almost_equal(cog[0, 0], [0.2733333, 0.2666667, 0.2])
    com = ko.molecule_centres(coords, [[0, 1, 2]], masses=np.array([1.0, 16.0, 1.0]))
    np.testing.assert_array_almost_equal(com[0, 0], [0.3788889, 0.2111111, 0.2])
    # across the periodic boundary: members at 0.95 and 0.05 have their centre at 0.0, not 0.5
    wrap = np.array([[0.95, 0.5, 0.5], [0.05, 0.5, 0.5]])[:, None, :]
    c = ko.molecule_centres(wrap, [[0, 1]])[0, 0]
    assert min(c[0], 1 - c[0]) < 1e-12 and abs(c[1] - 0.5) < 1e-12


def test_resampling_restatement_is_pinned_to_the_reference(golden_dir):
    """`_bootstrap` / `sample_until_normal` (kinisi/diffusion.py:258-289, 783-801) restated in the oracle give, with the
    same seeded RandomState, the values the reference + scikit-learn gave (tests/golden/make_golden.py resample)."""
    from synth_inputs import list_bootstrap_inputs
    g = np.load(os.path.join(golden_dir, 'resample_reference.npz'))
    a = np.array(ko.bootstrap_means(np.arange(1, 100, 1), 200, 100, np.random.RandomState(0)))
    assert np.array_equal(a, g['bootstrap_arange'])
    v, capped = ko.sample_until_normal(np.arange(1, 10, 1), 5, 100, 100, random_state=np.random.RandomState(0))
    assert capped and np.array_equal(v, g['until_normal_small'])
    v, capped = ko.sample_until_normal(g['skewed_population'], 7, 200, 2000, random_state=np.random.RandomState(1))
    assert capped and np.array_equal(v, g['until_normal_skewed'])
    dt, disp_3d, n_o, _ = list_bootstrap_inputs()
    rs = np.random.RandomState(0)
    means = []
    for i, d in enumerate(disp_3d):
        v, _ = ko.sample_until_normal(np.sum(d**2, axis=-1), n_o[i], 1000, 10000, 1e-3, rs)
        means.append(np.mean(v))
        if i == 0:
            assert np.array_equal(v, g['msd_first_distribution'])
    assert np.array_equal(np.array(means), g['msd_n_bootstrap'])


def test_counter_generator_known_answers():
    """Philox4x32-10 against the known answers published with Random123 (kat_vectors: all-zero and all-one inputs)
    and the multiply-high range reduction against exact integer arithmetic."""
    x = ko.philox4x32_10([0], [0], [0], [0], 0, 0)
    assert [int(v[0]) for v in x] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    x = ko.philox4x32_10([f], [f], [f], [f], f, f)
    assert [int(v[0]) for v in x] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    u = np.array([2**64 - 1, 12345678901234567890, 7, 2**63], dtype=np.uint64)
    for m in (1, 10, 2**40 + 3, 2**63 + 5):
        assert [int(a) for a in ko._mulhi64(u, m)] == [(int(a) * m) >> 64 for a in u]
    idx = ko.counter_resample_indices(1000, 20001, 5, 1234567)
    assert idx.size == 20001 and idx.min() >= 0 and idx.max() < 1000
    counts = np.bincount(idx, minlength=1000)
    assert abs(counts.mean() - 20.001) < 1e-9 and counts.std() < 6.0  # ~ Poisson(20)
    m = ko.counter_resample_means(np.arange(10.0), 64, 400, seed=3)
    assert abs(m.mean() - 4.5) < 5 * np.sqrt(8.25 / 64 / 400)


def test_reblocking_restatement_on_correlated_data():
    """The restated blocking method (pyblock is absent: unpinned) recovers the standard error of an AR(1) series, which
    the naive estimate misses by the correlation factor, and has no optimal block for constant data."""
    rng = np.random.default_rng(0)
    phi, n = 0.9, 2**15
    e = rng.normal(size=n)
    x = np.empty(n)
    x[0] = e[0]
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    stats = ko.reblock(x)
    assert [s[1] for s in stats] == [n >> l for l in range(15)]
    opt = ko.find_optimal_block(n, stats)
    true_se = np.sqrt(1.0 / (1 - phi)**2 / n)
    assert abs(stats[int(opt)][4] / true_se - 1) < 0.15 and stats[0][4] < 0.3 * true_se
    mean, var = ko.blocked_mean_and_variance(x)
    assert var == stats[int(opt)][4]**2 and mean == stats[int(opt)][2]
    assert np.isnan(ko.find_optimal_block(5, ko.reblock(np.ones(5))))
    odd = ko.reblock(np.arange(7.0))
    assert [s[1] for s in odd] == [7, 3] and odd[1][2] == 2.5  # (0.5, 2.5, 4.5): the 7th point is dropped


def test_universe_front_end_restatement(golden_dir):
    """The MDAnalysis walk (kinisi/parser.py:620-638) restated in the oracle against the reference run on a duck-typed
    universe: exact for float64 input, single-precision noise for the float32 arrays MDAnalysis hands out (the
    reference then works in float32; the oracle's unwrapping is float64)."""
    from synth_inputs import ase_case_inputs
    g = np.load(os.path.join(golden_dir, 'universe_reference.npz'))
    for name in ('ase_nvt', 'ase_npt_triclinic'):
        symbols, frac, latt, pp = ase_case_inputs(name)
        cart = np.einsum('nfk,fkj->fnj', frac, latt)
        fw = [i for i, sym in enumerate(symbols) if sym != 'Li']
        for tag, dtype, atol in (('f64', np.float64, 1e-13), ('f32', np.float32, 2e-5)):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                coords, cells = ko.universe_coords_latt(cart.astype(dtype), latt.astype(dtype))
                assert coords.dtype == dtype
                dc = ko.correct_drift(fw, ko.get_disp(coords, cells))
            np.testing.assert_allclose(dc, g[f'{name}_{tag}_dc'], rtol=0, atol=atol)
