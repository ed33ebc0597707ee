"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs,
against the golden vectors the unmodified reference produced, and against the reference's own KATs.

Tolerances (BASELINE.json north_star): MSD and variance <= 1e-9 relative, covariance <= 1e-6 relative,
posterior within Monte-Carlo error."""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

from duck_atoms import trajectory_from_arrays
from oracle import kinisi_oracle as ko
from synth_inputs import (ASE_CASES, CHARGES12, REGRESS_CASES, ase_case_inputs, list_bootstrap_inputs,
                          parser_random_disp, regress_case_inputs)

RTOL_MOMENT = 1e-9


@pytest.fixture(scope='module')
def kb():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from kinisi_b200 import analyze, diffusion, engine, parser
    engine.require_cuda()

    class NS:
        pass

    ns = NS()
    ns.analyze, ns.diffusion, ns.engine, ns.parser = analyze, diffusion, engine, parser
    return ns


@pytest.fixture(autouse=True)
def _default_variant_is_restored():
    """Every test starts on (and leaves behind) the shipped default moment kernel, the one bench.py times."""
    from kinisi_b200 import engine
    shipped = engine.DEFAULT_VARIANT
    assert shipped == engine.VARIANT_AUTO
    yield
    engine.DEFAULT_VARIANT = shipped


ALL_VARIANTS = [0, 1, 2, 3, 4]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + '.npz'))


def check_boot(b, g, prefix, rtol=RTOL_MOMENT, ngp_atol=1e-10):
    np.testing.assert_allclose(b.dt, g[prefix + 'dt'], rtol=1e-14, err_msg=prefix + 'dt')
    np.testing.assert_allclose(b.n, g[prefix + 'n'], rtol=rtol, err_msg=prefix + 'n')
    np.testing.assert_allclose(b.v, g[prefix + 'v'], rtol=rtol, err_msg=prefix + 'v')
    np.testing.assert_allclose(b.s, g[prefix + 's'], rtol=rtol, err_msg=prefix + 's')
    np.testing.assert_allclose(b.ngp, g[prefix + 'ngp'], rtol=max(1e-8, rtol), atol=ngp_atol, err_msg=prefix + 'ngp')
    np.testing.assert_allclose(np.asarray(b._n_o), g[prefix + 'n_o'], rtol=1e-15, err_msg=prefix + 'n_o')


# ---------------------------------------------------------------------------------------------
# stage 1 + 2 against the reference's golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('variant', ALL_VARIANTS)
@pytest.mark.parametrize('name', list(ASE_CASES))
def test_ase_parser_and_bootstraps(kb, golden_dir, name, variant, monkeypatch):
    monkeypatch.setattr(kb.engine, 'DEFAULT_VARIANT', variant)
    g = load(golden_dir, name)
    symbols, frac, latt, pp = ase_case_inputs(name)
    traj = trajectory_from_arrays(symbols, frac, latt)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        p = kb.parser.ASEParser(traj, **pp)
    np.testing.assert_allclose(p.dc, g['dc'], rtol=1e-11, atol=1e-12)
    np.testing.assert_array_equal(p.time_intervals, g['time_intervals'])
    np.testing.assert_allclose(p.delta_t, g['delta_t'], rtol=1e-15)
    np.testing.assert_allclose(p._n_o, g['n_o'], rtol=1e-15)
    assert p.volume == pytest.approx(float(g['volume']), rel=1e-12)
    np.testing.assert_array_equal(np.array(p.disp_3d.shapes), g['disp_shapes'])
    np.testing.assert_allclose(p.disp_3d[0], g['disp_first'], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(p.disp_3d[len(p.disp_3d) // 2], g['disp_mid'], rtol=1e-10, atol=1e-12)
    for dim in ('xyz', 'xy', 'z', 'xz'):
        b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, dimension=dim, progress=False)
        check_boot(b, g, f'msd_{dim}_')
        b = kb.diffusion.MSTDBootstrap(p.delta_t, p.disp_3d, p._n_o, dimension=dim, progress=False)
        check_boot(b, g, f'mstd_{dim}_')
        b = kb.diffusion.MSCDBootstrap(p.delta_t, p.disp_3d, CHARGES12, p._n_o, dimension=dim, progress=False)
        check_boot(b, g, f'mscd_{dim}_')
    b = kb.diffusion.MSCDBootstrap(p.delta_t, p.disp_3d, 2, p._n_o, progress=False)
    check_boot(b, g, 'mscd_scalar2_')
    b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, sub_sample_dt=3, progress=False)
    np.testing.assert_allclose(b.n, g['msd_sub3_n'], rtol=RTOL_MOMENT)
    np.testing.assert_allclose(b.v, g['msd_sub3_v'], rtol=RTOL_MOMENT)


def test_parser_on_displacement_array(kb, golden_dir):
    """Mirrors kinisi/tests/test_parser.py:31-67 (Parser fed with a numpy displacement array)."""
    g = load(golden_dir, 'parser_random')
    disp = parser_random_disp()
    p = kb.parser.Parser(disp, np.arange(100), [], 1, 1, progress=False)
    assert p.delta_t.size == 100
    assert p.disp_3d.shape_of(0) == (100, 100, 3)
    np.testing.assert_allclose(p.delta_t, g['lin_delta_t'])
    np.testing.assert_allclose(p._n_o, g['lin_n_o'])
    check_boot(kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False), g, 'lin_msd_')
    p = kb.parser.Parser(disp, np.arange(0, 100, 2), np.arange(1, 100, 2), 0.5, 4, min_dt=10., max_dt=150., n_steps=30,
                         spacing='logarithmic', progress=False)
    np.testing.assert_array_equal(p.time_intervals, g['log_intervals'])
    np.testing.assert_allclose(p.dc, g['log_dc'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(p._n_o, g['log_n_o'])
    check_boot(kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False), g, 'log_msd_')
    # a single atom: the parser drops the tail of the grid (kinisi/parser.py:214-215)
    p = kb.parser.Parser(disp[:1], np.arange(1), [], 1, 1, n_steps=20, progress=False)
    np.testing.assert_allclose(p.delta_t, g['one_delta_t'])
    np.testing.assert_allclose(p._n_o, g['one_n_o'])
    assert len(p.disp_3d) == g['one_shapes'].shape[0]
    check_boot(kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False), g, 'one_msd_')
    np.testing.assert_allclose(kb.parser.Parser.correct_drift(np.arange(1, 100, 2), disp),
                               ko.correct_drift(np.arange(1, 100, 2), disp), rtol=1e-13, atol=1e-15)


def test_list_seam(kb, golden_dir):
    """The *Bootstrap constructors on materialised lists, as kinisi/tests/test_diffusion.py feeds them."""
    g = load(golden_dir, 'list_bootstrap')
    dt, disp_3d, n_o, disp_skip = list_bootstrap_inputs()
    b = kb.diffusion.MSDBootstrap(dt, disp_3d, n_o, progress=False)
    assert b.n.shape == (10, ) and b.dt.shape == (10, )
    check_boot(b, g, 'msd_')
    np.testing.assert_allclose(b.v, b.s**2, rtol=1e-14)  # test_diffusion.py:147
    b = kb.diffusion.MSTDBootstrap(dt, disp_3d, n_o, progress=False)
    assert b.n.shape == (5, )  # half-grid rule, test_diffusion.py:402
    check_boot(b, g, 'mstd_')
    b = kb.diffusion.MSCDBootstrap(dt, disp_3d, 1, n_o, progress=False)
    assert b.n.shape == (5, )
    check_boot(b, g, 'mscd_')
    b = kb.diffusion.MSDBootstrap(np.linspace(100, 400, 4), disp_skip, np.ones(4), progress=False)
    assert b.n.shape == g['skip_msd_n'].shape  # skip rule, test_diffusion.py:228-243
    check_boot(b, g, 'skip_msd_')
    assert len(b.euclidian_displacements) == b.n.size
    assert b.euclidian_displacements[0].size == disp_skip[0].shape[0] * disp_skip[0].shape[1]


def test_reference_kats(kb, golden_dir):
    g = load(golden_dir, 'reference_kats')
    assert kb.diffusion.Bootstrap.ngp_calculation(np.array([1, 2, 3])) == pytest.approx(-0.3, abs=1e-14)
    symbols = list(g['drift_symbols'])
    traj = trajectory_from_arrays(symbols, g['drift_frac'], g['drift_latt'])
    p = kb.parser.ASEParser(traj, specie='H', time_step=1, step_skip=1, progress=False)
    np.testing.assert_allclose(p.dc[0], [[0, 0, 0], [0.2, 0.2, 0.2], [0.4, 0.4, 0.4], [1, 1, 1]], atol=1e-12)
    p = kb.parser.ASEParser(traj, specie='H', time_step=1, step_skip=1, progress=False, framework_indices=[])
    np.testing.assert_allclose(p.dc[0], g['nodrift_dc0'], atol=1e-12)
    b = kb.diffusion.MSDBootstrap(np.array([1.0]), [g['docs_disp']], np.array([384.0]), progress=False)
    assert b.n[0] == pytest.approx(5.1108279, rel=1e-7)
    assert b.v[0] == pytest.approx(0.0911415, rel=1e-6)
    assert b.ngp[0] == pytest.approx(0.4038798, rel=1e-6)
    assert b.n[0] == pytest.approx(float(g['docs_msd']), rel=1e-12)
    # the real Li6PS5Cl-like XDATCAR of kinisi/tests/test_analyze.py:92-98, float32 coordinates
    symbols = list(g['lips_symbols'])
    latt = np.tile(g['lips_latt'], (140, 1, 1))
    traj = trajectory_from_arrays(symbols, g['lips_frac'].astype(np.float64), latt)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        an = kb.analyze.DiffusionAnalyzer.from_ase(traj, parser_params=dict(specie='Li', time_step=2.0, step_skip=50,
                                                                            progress=False))
    assert an._disp_3d.shape_of(0) == (192, 140, 3) and an._disp_3d.shape_of(len(an._disp_3d) - 1) == (192, 1, 3)
    assert an._delta_t.max() == 14000.0
    check_boot(an._diff, g, 'lips_msd_')


# ---------------------------------------------------------------------------------------------
# the moment kernels against the oracle's raw sums: tiles, partial tiles, both variants, f32 input
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('variant', [0, 0 | (16 << 8), 1, 2, 3, 4])
@pytest.mark.parametrize('dimension', ['xyz', 'xz', 'y'])
def test_self_moment_kernels_raw_sums(kb, variant, dimension):
    n_atoms, n_frames = 5, 9300  # > 2 tiles of 4096 and > 4 tiles of 2048, last tile partial
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=11.0, sigma=0.12, seed=21)
    dc = ko.get_disp(frac, latt)
    lags = np.unique(np.concatenate([[1, 2, 3, 17, 255, 256, 257, 511, 512, 513, 2047, 2048, 2049, 4095, 4096, 4097],
                                     np.linspace(1, n_frames, 57, dtype=int),
                                     [n_frames - 4096, n_frames - 2, n_frames - 1, n_frames]]))
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags, dimension)
    dev = torch.device('cuda')
    pitch = kb.engine.pitch_for(n_frames)
    dct = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dct[:, :, :n_frames] = torch.from_numpy(dc).permute(0, 2, 1).to(dev)
    sums = kb.engine.self_moments(dct, n_frames, torch.as_tensor(lags.astype(np.int32)).to(dev),
                                  kb.engine.dim_mask(dimension), variant).cpu().numpy()
    np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)


@pytest.mark.parametrize('variant', ALL_VARIANTS)
def test_long_lag_grid_is_chunked(kb, variant):
    """More lags than one launch holds in shared memory (every lag of a 1500-frame trajectory)."""
    n_atoms, n_frames = 3, 1500
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=9.0, sigma=0.2, seed=5)
    dc = ko.get_disp(frac, latt)
    lags = np.arange(1, n_frames + 1)
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags)
    dev = torch.device('cuda')
    pitch = kb.engine.pitch_for(n_frames)
    dct = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dct[:, :, :n_frames] = torch.from_numpy(dc).permute(0, 2, 1).to(dev)
    sums = kb.engine.self_moments(dct, n_frames, torch.as_tensor(lags.astype(np.int32)).to(dev), 7, variant).cpu().numpy()
    np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)



def _device_dc(kb, dc):
    dev = torch.device('cuda')
    n_atoms, n_frames = dc.shape[0], dc.shape[1]
    dct = torch.zeros((n_atoms, 3, kb.engine.pitch_for(n_frames)), dtype=torch.float64, device=dev)
    dct[:, :, :n_frames] = torch.from_numpy(dc).permute(0, 2, 1).to(dev)
    return dct


RUN_GRIDS = {
    # kinisi's default grids (parser.py:150-168) at several truncation patterns, plus adversarial ones
    'linspace_100': lambda T: np.linspace(1, T, 100, dtype=int),
    'linspace_drift_up': lambda T: np.linspace(1, T, int(T / 12.09), dtype=int),
    'linspace_drift_down': lambda T: np.linspace(1, T, int(T / 12.91), dtype=int),
    'linspace_drift_fast': lambda T: np.linspace(2, T - 3, int(T / 9.13), dtype=int),
    'linspace_alternating': lambda T: np.linspace(1, T, int(T / 50.4949), dtype=int),
    'linspace_min_dt': lambda T: np.linspace(37, T - 11, 64, dtype=int),
    'geomspace': lambda T: np.unique(np.geomspace(1, T, 100, dtype=int)),
    'dense_1_to_300': lambda T: np.arange(1, 301),
    'stride_2': lambda T: np.arange(2, T + 1, 2),
    'stride_33_short': lambda T: np.arange(5, 5 + 33 * 11, 33),
    'stride_half': lambda T: np.array([3, 3 + T // 2 - 2, 3 + 2 * (T // 2 - 2)]),
    'two_interleaved': lambda T: np.unique(np.concatenate([np.arange(1, T, 97), np.arange(13, T, 211)])),
    'primes': lambda T: np.array([2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71]),
    'single': lambda T: np.array([T // 3]),
    'last_two': lambda T: np.array([T - 1, T]),
    'all': lambda T: np.arange(1, T + 1),
}


@pytest.mark.parametrize('grid', sorted(RUN_GRIDS))
@pytest.mark.parametrize('n_frames', [700, 5431])
def test_tiled_lag_window_kernel_on_structured_and_irregular_grids(kb, grid, n_frames):
    """kb2_self_moments_tile (the default: tiles staged in shared memory, windows of <= 8 lags k = I*D + r with consecutive
    I, leftovers as windows of one) against the oracle's raw sums on every kind of lag grid, 7 atoms (strips that end
    inside a row, tiles of several strips, row blocks, shared and separate partner rows)."""
    n_atoms = 7
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=10.0, sigma=0.15, seed=78)
    dc = ko.get_disp(frac, latt)
    lags = RUN_GRIDS[grid](n_frames)
    lags = lags[(lags >= 1) & (lags <= n_frames)]
    if grid == 'all' and n_frames > 2000:
        lags = lags[:1700]  # seven chunks of the 256-lag launch limit
    dct = _device_dc(kb, dc)
    for mask, dim in ((7, 'xyz'), (6, 'yz'), (1, 'x')):
        S1, S2, _, _ = ko.raw_sums_streaming(dc, lags, dim)
        sums = kb.engine.self_moments(dct, n_frames, lags, mask, kb.engine.VARIANT_TILE).cpu().numpy()
        np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
        np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)
    wins, n_tiles = kb.engine.self_moments_tiles(lags[:256], n_frames)
    length = np.array([w[2] for w in wins])
    assert int(np.sum(length)) == min(len(lags), 256) and np.all(length <= 8) and np.all(length >= 1) and n_tiles >= 1
    assert all(w[1] >= 32 for w in wins)  # row lengths: a strip of 32 columns at least


@pytest.mark.parametrize('grid', sorted(RUN_GRIDS))
@pytest.mark.parametrize('n_frames', [700, 5431])
def test_lag_window_kernel_on_structured_and_irregular_grids(kb, grid, n_frames):
    """kb2_self_moments_win (the default: partner window + accumulators in registers, progressions cut into windows of
    <= 8 lags, leftovers as windows of one) against the oracle's raw sums on every kind of lag grid."""
    n_atoms = 3
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=10.0, sigma=0.15, seed=77)
    dc = ko.get_disp(frac, latt)
    lags = RUN_GRIDS[grid](n_frames)
    lags = lags[(lags >= 1) & (lags <= n_frames)]
    if grid == 'all' and n_frames > 2000:
        lags = lags[:1700]  # seven chunks of the 256-lag launch limit
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags)
    dct = _device_dc(kb, dc)
    sums = kb.engine.self_moments(dct, n_frames, lags, 7, kb.engine.VARIANT_WIN).cpu().numpy()
    np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags, 'xz')
    sums = kb.engine.self_moments(dct, n_frames, lags, 5, kb.engine.VARIANT_WIN).cpu().numpy()
    np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)
    length = np.array([w[2] for w in kb.engine.self_moments_windows(lags[:256], n_frames)])
    assert int(np.sum(length)) == min(len(lags), 256) and np.all(length <= 8) and np.all(length >= 1)


@pytest.mark.parametrize('grid', sorted(RUN_GRIDS))
@pytest.mark.parametrize('n_frames', [700, 5431])
def test_arithmetic_run_kernel_on_structured_and_irregular_grids(kb, grid, n_frames):
    """kb2_self_moments_runs (register-window kernel + direct kernel for the leftover lags) against the
    oracle's raw sums on every kind of lag grid: bit-for-bit the same terms, reduction order aside."""
    n_atoms = 3
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=10.0, sigma=0.15, seed=77)
    dc = ko.get_disp(frac, latt)
    lags = RUN_GRIDS[grid](n_frames)
    lags = lags[(lags >= 1) & (lags <= n_frames)]
    if grid == 'all' and n_frames > 2000:
        lags = lags[:1700]  # nine chunks of the 192-lag launch limit
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags)
    sums = kb.engine.self_moments(_device_dc(kb, dc), n_frames, lags, 7, kb.engine.VARIANT_RUNS).cpu().numpy()
    np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)
    runs, rest = kb.engine.self_moments_plan(lags[:192], n_frames)
    assert sum(r[2] for r in runs) + rest == min(len(lags), 192)


@pytest.mark.parametrize('variant', [2, 3, 4])
def test_structured_kernels_many_atoms_and_tiles(kb, variant):
    """Enough atoms and frames that every persistent CTA takes several work items, with a short trajectory
    tail (partial tiles, warps that straddle two origin tiles, inactive lanes)."""
    n_atoms, n_frames = 37, 12347
    rng = np.random.default_rng(5)
    dc = np.cumsum(rng.normal(0, 0.1, size=(n_atoms, n_frames, 3)), axis=1)
    dc[:, 0] = 0.0
    lags = np.linspace(1, n_frames, 100, dtype=int)
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags)
    dct = _device_dc(kb, dc)
    for mask, dim in ((7, 'xyz'), (5, 'xz'), (2, 'y')):
        if dim != 'xyz':
            S1, S2, _, _ = ko.raw_sums_streaming(dc, lags, dim)
        sums = kb.engine.self_moments(dct, n_frames, lags, mask, variant).cpu().numpy()
        np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
        np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_stage1_dtypes_and_collective(kb, dtype):
    """float32 / float64 wrapped coordinates, framework drift, collective and charge-weighted sums."""
    n_mobile, n_fw, n_frames = 40, 24, 3000
    frac, latt = ko.synthetic_walk(n_mobile, n_frames, box=14.0, sigma=0.1, seed=33, n_framework=n_fw, dtype=dtype)
    idx, fw = list(range(n_mobile)), list(range(n_mobile, n_mobile + n_fw))
    parsed = ko.parse(ko.get_disp(frac.astype(np.float64), latt), idx, fw, 0.002, 2, n_steps=64)
    traj = kb.parser.Parser.get_disp(frac, latt)
    p = kb.parser.Parser(traj, idx, fw, 0.002, 2, n_steps=64, progress=False)
    np.testing.assert_allclose(p.dc, parsed['dc'], rtol=1e-11, atol=1e-11)
    q = np.where(np.arange(n_mobile) % 3 == 0, -2.0, 1.0)
    for kind, cls, args in (('msd', kb.diffusion.MSDBootstrap, ()), ('mstd', kb.diffusion.MSTDBootstrap, ()),
                            ('mscd', kb.diffusion.MSCDBootstrap, (q, ))):
        b = cls(p.delta_t, p.disp_3d, *args, p._n_o, progress=False)
        ref = ko.bootstrap_streaming(parsed, kind, q if kind == 'mscd' else None)
        for key in ('dt', 'n', 'v', 's'):
            np.testing.assert_allclose(getattr(b, key), ref[key], rtol=RTOL_MOMENT, err_msg=f'{kind} {key}')
        np.testing.assert_allclose(b.ngp, ref['ngp'], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_frame_major_ingest(kb, dtype, monkeypatch):
    """kb2_frames_to_atoms: frame-major input, from the device and from the host in several blocks, gives
    the atom-major array with the duplicated first frame that np.concatenate builds in the reference (parser.py:120,
    323-324); the parser on top of it gives the same displacements."""
    n_atoms, n_frames = 37, 131
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=9.0, sigma=0.2, seed=8, dtype=dtype)
    frames = np.ascontiguousarray(np.transpose(frac[:, 1:], (1, 0, 2)))  # [frame, site, 3] without the duplicate
    dev = torch.device('cuda')
    t_dev = kb.parser.Trajectory.from_frames(torch.from_numpy(frames).to(dev), latt[:1], device=dev)
    np.testing.assert_array_equal(t_dev.coords.cpu().numpy(), frac)
    monkeypatch.setattr(kb.parser, 'STAGE_BYTES', 10 * n_atoms * 3 * frames.itemsize)  # 14 blocks of 10 frames
    t_host = kb.parser.Trajectory.from_frames(frames, latt[:1], device=dev)
    np.testing.assert_array_equal(t_host.coords.cpu().numpy(), frac)
    t_nodup = kb.parser.Trajectory.from_frames(frames, None, mode=kb.parser.MODE_DISPLACEMENT, device=dev,
                                               duplicate_first=False)
    np.testing.assert_array_equal(t_nodup.coords.cpu().numpy(), frac[:, 1:])
    idx = list(range(0, n_atoms, 2))
    fw = list(range(1, n_atoms, 2))
    a = kb.parser.Parser(kb.parser.Parser.get_disp_from_frames(frames, latt[1:]), idx, fw, 0.001, 1, n_steps=20, progress=False)
    b = kb.parser.Parser(kb.parser.Parser.get_disp(frac, latt), idx, fw, 0.001, 1, n_steps=20, progress=False)
    # (the framework sums of the two runs are accumulated in a different order: rounding-level differences)
    np.testing.assert_allclose(a.disp_3d.dc.cpu().numpy(), b.disp_3d.dc.cpu().numpy(), rtol=1e-12, atol=1e-12)


def test_molecule_centres(kb):
    """kb2_molecule_centres against the oracle's restatement of kinisi/parser.py:706-724, the reference's
    KAT (kinisi/tests/test_parser.py:127-149) through the pymatgen front-end, and its error behaviour."""
    rng = np.random.RandomState(4)
    n_atoms, n_frames = 23, 300
    coords = rng.uniform(0, 1, size=(n_atoms, 1, 3)) + np.cumsum(rng.normal(0, 0.02, size=(n_atoms, n_frames, 3)), axis=1)
    coords = coords % 1.0
    members = np.array([[0, 5, 7, 2], [11, 12, 13, 14], [22, 1, 3, 20]])
    masses = np.array([1.0, 16.0, 12.0, 1.0])
    dev = torch.device('cuda')
    for dtype in (np.float32, np.float64):
        c = coords.astype(dtype)
        for w in (None, masses):
            ref = ko.molecule_centres(c.astype(np.float64), members, w)
            got = kb.engine.molecule_centres(torch.from_numpy(c).to(dev), torch.from_numpy(members.astype(np.int32)).to(dev),
                                             None if w is None else torch.from_numpy(w).to(dev)).cpu().numpy()
            d = np.abs(got - ref)
            assert np.minimum(d, 1.0 - d).max() < 1e-12  # equal up to the periodic image
    # the reference's KAT through the front-end (1-based indices, duck-typed pymatgen structures)
    class Site:
        def __init__(self, s):
            self.specie = s

    class Lattice:
        matrix = np.eye(3) * 5.0

    class Structure:
        lattice, volume = Lattice(), 125.0

        def __init__(self, frac):
            self.frac_coords = np.asarray(frac, dtype=float)

        def __iter__(self):
            return iter([Site('H'), Site('He'), Site('H')])

    f0 = [[0.2, 0.2, 0.2], [0.4, 0.2, 0.2], [0.22, 0.4, 0.2]]
    f1 = [[0.4, 0.2, 0.2], [0.4, 0.2, 0.2], [0.2, 0.4, 0.2]]
    structs = [Structure(f0), Structure(f1), Structure(f1), Structure(f1)]
    pk = dict(specie=None, time_step=1.0, step_skip=1, specie_indices=[[1, 2, 3]], progress=False)
    with pytest.warns(UserWarning):
        data = kb.parser.PymatgenParser(structs, **pk)
    assert data.indices == [0] and data.drift_indices == []
    np.testing.assert_array_almost_equal(data.coords_check, [[[0.2733333, 0.2666667, 0.2]]])
    with pytest.warns(UserWarning):
        data = kb.parser.PymatgenParser(structs, masses=[1, 16, 1], **pk)
    np.testing.assert_array_almost_equal(data.coords_check, [[[0.3788889, 0.2111111, 0.2]]])
    with pytest.raises(ValueError), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        kb.parser.PymatgenParser(structs, masses=[1, 16], **pk)
    with pytest.raises(ValueError), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        kb.parser.PymatgenParser(structs, **dict(pk, specie_indices=[[1, 2], [3]]))


def test_staged_ingest_matches_resident_input(kb, monkeypatch):
    """Host input is transferred block by block behind the stage-1 / stage-2 kernels (Parser._build_displacement_staged);
    the result must not depend on the blocking, on the order of the indices, or on where the input lives."""
    n_all, n_frames = 37, 2500
    frac, latt = ko.synthetic_walk(n_all, n_frames, box=11.0, sigma=0.12, seed=5, dtype=np.float32)
    rng = np.random.RandomState(3)
    perm = rng.permutation(n_all)
    idx, fw = list(perm[:17]), list(perm[17:30])  # scattered, unsorted; 7 atoms are neither (never transferred)
    dev = torch.device('cuda')
    ref = kb.parser.Parser(kb.parser.Parser.get_disp(torch.from_numpy(frac).to(dev), latt), idx, fw, 0.001, 1, n_steps=50,
                           progress=False)
    ref_b = kb.diffusion.MSDBootstrap(ref.delta_t, ref.disp_3d, ref._n_o, progress=False)
    row_bytes = (n_frames + 1) * 3 * 4
    for block_rows, pinned in ((1, True), (4, True), (5, False), (1000, True)):
        monkeypatch.setattr(kb.parser, 'STAGE_BYTES', block_rows * row_bytes)
        host = torch.from_numpy(frac.copy())
        if pinned:
            host = host.pin_memory()
        traj = kb.parser.Parser.get_disp(host, latt)
        assert not traj.resident
        p = kb.parser.Parser(traj, idx, fw, 0.001, 1, n_steps=50, progress=False)
        assert len(p.disp_3d.row_blocks()) == -(-17 // block_rows)
        b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
        np.testing.assert_allclose(b.n, ref_b.n, rtol=1e-13)
        np.testing.assert_allclose(b.v, ref_b.v, rtol=1e-11)
        np.testing.assert_array_equal(p.disp_3d.dc.cpu().numpy(), ref.disp_3d.dc.cpu().numpy())
        assert traj._staged is not None and int(traj._staged.sum()) == 30
        np.testing.assert_array_equal(p.disp_3d[3], ref.disp_3d[3])
        # a later consumer of the whole array (here the all-atom view) stages the rest
        np.testing.assert_allclose(p.dc, ref.dc, rtol=0, atol=0)
        assert traj.resident


def test_analyses_in_flight_on_two_streams(kb, monkeypatch):
    """Several analyses queued back to back on alternating streams, host (staged) and device input mixed, results read
    only at the end: every one must equal its serial evaluation (moments and GLS optimum to rounding, posterior
    statistically)."""
    dev = torch.device('cuda')
    monkeypatch.setattr(kb.parser, 'STAGE_BYTES', 1 << 20)
    cases = []
    for seed, (n_mob, n_fw, n_frames) in enumerate([(30, 10, 4000), (12, 20, 2500), (40, 0, 3000), (25, 7, 5200)]):
        frac, latt = ko.synthetic_walk(n_mob, n_frames, box=12.0, sigma=0.1, seed=40 + seed, n_framework=n_fw,
                                       dtype=np.float32)
        cases.append((frac, latt, list(range(n_mob)), list(range(n_mob, n_mob + n_fw))))

    def analyse(case, where, rs):
        frac, latt, idx, fw = case
        coords = torch.from_numpy(frac).to(dev) if where == 'device' else torch.from_numpy(frac).pin_memory()
        p = kb.parser.Parser(kb.parser.Parser.get_disp(coords, latt), idx, fw, 0.001, 1, n_steps=60, progress=False)
        b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
        b.diffusion(float(b.dt[6]), progress=False, random_state=np.random.RandomState(rs), n_samples=300, n_burn=100)
        b._keep = (p, coords)
        return b

    order = [(i % len(cases), 'device' if (i // 2) % 2 else 'host', 100 + i) for i in range(12)]
    serial = []
    for ci, where, rs in order:
        b = analyse(cases[ci], where, rs)
        serial.append((b.n.copy(), b.v.copy(), b.flatchain.copy(), b._theta_ml.cpu().numpy()))
    lanes = [torch.cuda.Stream(dev) for _ in range(2)]
    queued = []
    for i, (ci, where, rs) in enumerate(order):
        with torch.cuda.stream(lanes[i % 2]):
            queued.append(analyse(cases[ci], where, rs))
    for b, (n, v, fc, ml) in zip(queued, serial):
        np.testing.assert_allclose(b.n, n, rtol=1e-13)
        np.testing.assert_allclose(b.v, v, rtol=1e-11)
        # the GLS optimum and the quadratic form (gls[0:5], gls[7]) agree to rounding ...
        mine = b._theta_ml.cpu().numpy()
        np.testing.assert_allclose(mine[[0, 1, 2, 3, 4, 7]], ml[[0, 1, 2, 3, 4, 7]], rtol=1e-8)
        # ... the chains only statistically: the moment sums are reduced with atomics (rounding-level differences from
        # run to run) and the stretch move amplifies a perturbation by exp(0.079) per accepted move on average, so two
        # chains with the same seed decorrelate within a few hundred steps
        assert b.flatchain.shape == fc.shape
        sd = fc[:, 0].std()
        assert abs(b.flatchain[:, 0].mean() - fc[:, 0].mean()) < 0.5 * sd
        assert b.flatchain[:, 0].std() == pytest.approx(sd, rel=0.35)

# ---------------------------------------------------------------------------------------------
# bootstrap=True: the device resampler
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n_values,n_draws', [(1, 1), (9, 5), (1000, 2), (1000, 4097), (123457, 50001), (77, 2048)])
def test_resample_means_match_the_counter_restatement(kb, n_values, n_draws):
    """kb2_resample_means against the numpy restatement of the same counter-based draws: the indices are exact, the
    means agree to the reordering of a float64 sum; batches continue each other and repeat bit for bit."""
    rng = np.random.default_rng(n_values + n_draws)
    values = rng.exponential(size=n_values)**2
    dev = torch.from_numpy(values).cuda()
    seed = 0x9E3779B97F4A7C15 ^ n_draws
    got = kb.engine.resample_means(dev, n_draws, 12, 0, seed, 3).cpu().numpy()
    want = ko.counter_resample_means(values, n_draws, 12, 0, seed, 3)
    np.testing.assert_allclose(got, want, rtol=1e-13)
    again = kb.engine.resample_means(dev, n_draws, 12, 0, seed, 3).cpu().numpy()
    assert np.array_equal(got, again)
    tail = kb.engine.resample_means(dev, n_draws, 5, 7, seed, 3).cpu().numpy()
    assert np.array_equal(tail, got[7:])
    other = kb.engine.resample_means(dev, n_draws, 12, 0, seed + 1, 3).cpu().numpy()
    if n_values > 1:
        assert not np.array_equal(other, got)
    idx = ko.counter_resample_indices(n_values, n_draws, 4, seed, 3)
    assert idx.min() >= 0 and idx.max() < n_values


def test_resample_populations(kb):
    """kb2_lag_squares / kb2_disp_squares against the reference's per-lag arrays (parser.py:205-207, diffusion.py:574, 659)."""
    rng = np.random.default_rng(5)
    dc = np.cumsum(rng.normal(size=(7, 300, 3)), axis=1)
    dct = _device_dc(kb, dc)
    w = rng.normal(size=7)
    for lag in (1, 2, 37, 299, 300):
        disp = ko.displacements_for_lag(dc, lag)
        for mask, dims in ((7, [0, 1, 2]), (5, [0, 2]), (2, [1])):
            want = np.sum(disp[:, :, dims]**2, axis=-1)
            got = kb.engine.lag_squares(dct, 300, lag, mask).cpu().numpy()
            np.testing.assert_allclose(got, want.flatten(), rtol=1e-13, atol=1e-300)
            arr = torch.from_numpy(np.ascontiguousarray(disp)).cuda()
            np.testing.assert_allclose(kb.engine.disp_squares(arr, mask).cpu().numpy(), want.flatten(), rtol=1e-14)
            coll = np.sum(np.sum(disp * w[:, None, None], axis=0)[:, dims]**2, axis=-1)
            wd = torch.from_numpy(w).cuda()
            np.testing.assert_allclose(kb.engine.disp_squares(arr, mask, wd, True).cpu().numpy(), coll, rtol=1e-12)
            S = kb.engine.atom_sum(dct, wd)
            np.testing.assert_allclose(kb.engine.lag_squares(S.unsqueeze(0), 300, lag, mask).cpu().numpy(), coll, rtol=1e-11,
                                       atol=1e-18)


def test_sample_until_normal(kb, golden_dir):
    """The reference's own tests of sample_until_normal (kinisi/tests/test_diffusion.py:83-103) and the distribution of
    the resampled means against the reference's (tests/golden/resample_reference.npz, made with numpy's generator:
    only the statistics can agree)."""
    B = kb.diffusion.Bootstrap
    rng = np.random.RandomState(42)
    d1 = B.sample_until_normal(rng.normal(0, 1, size=2000), 50, 1000, 100000, random_state=np.random.RandomState(1))
    assert d1.size == 1000  # test_diffusion.py:83-89: normal at the first test
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        a = B.sample_until_normal(np.arange(1, 10, 1), 5, 100, 100, random_state=np.random.RandomState(0))
        b = B.sample_until_normal(np.arange(1, 10, 1), 5, 100, 100, random_state=np.random.RandomState(0))
    assert any('maximum number of resamples' in str(x.message) for x in w)
    assert a.size == 100 and np.array_equal(a.samples, b.samples)  # test_diffusion.py:91-103
    g = load(golden_dir, 'resample_reference')
    ref = g['until_normal_small']
    assert abs(a.samples.mean() - 5.0) < 5 * ref.std() / 10 and 0.6 < a.samples.std() / ref.std() < 1.6
    # a skewed population with few draws per resample: the loop adds hundreds until the test passes or the cap is hit
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        d = B.sample_until_normal(g['skewed_population'], 7, 200, 2000, random_state=np.random.RandomState(1))
    ref = g['until_normal_skewed']
    assert d.size == ref.size == 2000
    pop = g['skewed_population']
    assert abs(d.samples.mean() - pop.mean()) < 5 * pop.std() / np.sqrt(7 * 2000)
    assert 0.8 < d.samples.std() / ref.std() < 1.25
    assert 0.5 < np.percentile(d.samples, 97.5) / np.percentile(ref, 97.5) < 2.0


def test_bootstrap_true(kb, golden_dir):
    """`bootstrap=True` through the constructors (kinisi/diffusion.py:578-583, 663-669, 755-761), on the inputs of
    kinisi/tests/test_diffusion.py:192-208, 451-466, against the reference's resampled moments."""
    g = load(golden_dir, 'resample_reference')
    dt, disp_3d, n_o, _ = list_bootstrap_inputs()
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        b1 = kb.diffusion.MSDBootstrap(dt, disp_3d, n_o, bootstrap=True, random_state=np.random.RandomState(0), progress=False)
        b2 = kb.diffusion.MSDBootstrap(dt, disp_3d, n_o, bootstrap=True, random_state=np.random.RandomState(0), progress=False)
        t1 = kb.diffusion.MSTDBootstrap(dt, disp_3d, n_o, bootstrap=True, random_state=np.random.RandomState(0), progress=False)
        c1 = kb.diffusion.MSCDBootstrap(dt, disp_3d, 1, n_o, bootstrap=True, random_state=np.random.RandomState(0),
                                        progress=False)
    check_boot(b1, g, 'msd_')  # n, v, s, ngp are the unresampled ones (diffusion.py:596-600)
    assert len(b1._distributions) == 10 and all(d.size >= 1000 for d in b1._distributions)
    assert b1._distributions[-1].size == b2._distributions[-1].size
    assert np.array_equal(b1._distributions[-1].samples, b2._distributions[-1].samples)  # test_diffusion.py:206-208
    sizes = np.array([d.size for d in b1._distributions])
    se = np.sqrt(g['msd_v_bootstrap'] / sizes)
    assert np.all(np.abs(b1._n_bootstrap - b1.n) < 5 * se)
    assert np.all(np.abs(b1._v_bootstrap / g['msd_v_bootstrap'] - 1) < 0.3)
    np.testing.assert_allclose(b1._s_bootstrap**2, b1._v_bootstrap, rtol=1e-12)
    d = b1.to_dict()
    assert len(d['distributions']) == 10 and d['n_bootstrap'].shape == (10, )
    # collective: 20..16 values with one draw per resample never look normal: the cap is reached, as in the reference
    check_boot(t1, g, 'mstd_')
    assert np.array_equal(np.array([d.size for d in t1._distributions]), g['mstd_sizes'])
    assert np.all(np.abs(t1._n_bootstrap - t1.n) < 5 * np.sqrt(g['mstd_v_bootstrap'] / 10000))
    assert np.all(np.abs(t1._v_bootstrap / g['mstd_v_bootstrap'] - 1) < 0.15)
    assert np.array_equal(np.array([d.size for d in c1._distributions]), g['mscd_sizes'])
    assert np.all(np.abs(c1._v_bootstrap / g['mscd_v_bootstrap'] - 1) < 0.15)


def test_bootstrap_true_fused_path_equals_list_seam(kb):
    """From the parser's lazy sequence the populations come straight from the cumulative displacements
    (kb2_lag_squares): same values in the same order as the materialised list, hence the same resamples."""
    rng = np.random.default_rng(2)
    dc = np.cumsum(rng.normal(scale=0.1, size=(24, 400, 3)), axis=1)
    dc[:, 0] = 0.0
    charges = np.where(np.arange(24) % 2, 1.0, -1.0)
    p = kb.parser.Parser(dc, np.arange(24), [], 1.0, 1, n_steps=12, progress=False)
    listed = [np.asarray(d) for d in p.disp_3d]
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        for cls, args in ((kb.diffusion.MSDBootstrap, ()), (kb.diffusion.MSTDBootstrap, ()),
                          (kb.diffusion.MSCDBootstrap, (charges, ))):
            kw = dict(bootstrap=True, n_resamples=300, max_resamples=1000, progress=False)
            a = cls(p.delta_t, p.disp_3d, *args, p._n_o, random_state=np.random.RandomState(7), **kw)
            b = cls(p.delta_t, listed, *args, p._n_o, random_state=np.random.RandomState(7), **kw)
            assert len(a._distributions) == len(b._distributions) == a.n.size
            for x, y in zip(a._distributions, b._distributions):
                assert x.size == y.size
                np.testing.assert_allclose(x.samples, y.samples, rtol=1e-9)
            np.testing.assert_allclose(a._n_bootstrap, a.n, rtol=0.2)


@pytest.mark.parametrize('n_values', [2, 3, 5, 64, 1000, 4097, 100003])
def test_reblock_levels_match_the_restatement(kb, n_values):
    """kb2_reblock against the numpy restatement of pyblock.blocking.reblock (oracle; parity unpinned: pyblock is absent)."""
    rng = np.random.default_rng(n_values)
    x = np.cumsum(rng.normal(size=n_values)) * 0.1 + rng.exponential(size=n_values)  # serially correlated
    want = ko.reblock(x)
    acc = kb.engine.reblock(torch.from_numpy(x).cuda()).cpu().numpy()
    assert acc.shape[0] == len(want)
    for lvl, (block, ndata, mean, cov, std_err, _) in enumerate(want):
        assert ndata == n_values >> lvl
        np.testing.assert_allclose(acc[lvl, 0] / ndata, mean, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(acc[lvl, 1] / (ndata - 1), cov, rtol=1e-10)
    mean, var = kb.diffusion.Bootstrap._blocked(torch.from_numpy(x).cuda())
    wm, wv = ko.blocked_mean_and_variance(x)
    np.testing.assert_allclose([mean, var], [wm, wv], rtol=1e-10)


def test_block_true(kb):
    """`block=True` through the constructors (kinisi/diffusion.py:584-595, 670-681, 762-773): n, v, s become the blocked
    estimates of the flattened population of every lag; the NGP and dt are untouched; the fused path agrees with the list."""
    rng = np.random.default_rng(3)
    dc = np.cumsum(rng.normal(scale=0.1, size=(16, 600, 3)), axis=1)
    dc[:, 0] = 0.0
    charges = np.where(np.arange(16) % 2, 1.0, -1.0)
    p = kb.parser.Parser(dc, np.arange(16), [], 1.0, 1, n_steps=10, progress=False)
    listed = [np.asarray(d) for d in p.disp_3d]
    plain = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    for dimension, sl in (('xyz', [0, 1, 2]), ('xz', [0, 2])):
        a = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, block=True, dimension=dimension, progress=False)
        b = kb.diffusion.MSDBootstrap(p.delta_t, listed, p._n_o, block=True, dimension=dimension, progress=False)
        want = np.array([ko.blocked_mean_and_variance(np.sum(d[:, :, sl]**2, axis=-1)) for d in listed])
        for x in (a, b):
            np.testing.assert_allclose(x.n, want[:, 0], rtol=1e-9)
            np.testing.assert_allclose(x.v, want[:, 1], rtol=1e-9)
            np.testing.assert_allclose(x.s, np.sqrt(want[:, 1]), rtol=1e-9)
            np.testing.assert_allclose(x.dt, plain.dt)
    np.testing.assert_allclose(a.ngp[0], kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, dimension='xz',
                                                                    progress=False).ngp[0], rtol=1e-12)
    # overlapping origins are serially correlated: the blocked variance of the mean must exceed the naive one
    naive = np.array([np.var(np.sum(d**2, axis=-1), ddof=1) / d[..., 0].size for d in listed])
    a = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, block=True, progress=False)
    assert np.all(a.v[1:5] > 2 * naive[1:5])
    half = len(listed) // 2
    for cls, args, w, sl in ((kb.diffusion.MSTDBootstrap, (), np.ones(16), [0, 1, 2]),
                             (kb.diffusion.MSCDBootstrap, (charges, ), charges, [0, 1, 2])):
        a = cls(p.delta_t, p.disp_3d, *args, p._n_o, block=True, progress=False)
        b = cls(p.delta_t, listed, *args, p._n_o, block=True, progress=False)
        want = np.array([ko.blocked_mean_and_variance(np.sum(np.sum(d * w[:, None, None], axis=0)[:, sl]**2, axis=-1))
                         for d in listed[:half]])
        for x in (a, b):
            np.testing.assert_allclose(x.n, want[:, 0], rtol=1e-8)
            np.testing.assert_allclose(x.v, want[:, 1], rtol=1e-8)
    # the regression runs on the blocked variances
    a = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, block=True, progress=False)
    a.diffusion(p.delta_t[2], progress=False, n_samples=200, n_burn=100)
    assert a.D.samples.size == 200 * 32 // 10 and np.isfinite(a.D.n)  # thin = 10


def test_dictionary_roundtrip(kb):
    """to_dict / from_dict of the analyzers and the bootstrap (kinisi/tests/test_analyze.py:59-90, 255-275;
    kinisi/diffusion.py:84-176): the restored object answers from the stored host arrays."""
    symbols, frac, latt, pp, start_idx, kind = regress_case_inputs('regress_msd')
    traj = trajectory_from_arrays(symbols, frac, latt)
    A = kb.analyze.DiffusionAnalyzer
    a = A.from_ase(traj, parser_params=pp, uncertainty_params={'progress': False, 'bootstrap': True, 'n_resamples': 200,
                                                              'max_resamples': 400,
                                                              'random_state': np.random.RandomState(0)})
    with warnings.catch_warnings(record=True):
        warnings.simplefilter('always')
        a.diffusion(a.dt[start_idx], {'progress': False, 'n_samples': 300, 'n_burn': 100})
    d = a.to_dict()
    assert set(['delta_t', 'disp_3d', 'n_o', 'volume', 'diff']) == set(d.keys())
    for key in ('displacements', 'delta_t', 'n_o', 'max_obs', 'dt', 'n', 's', 'v', 'n_bootstrap', 's_bootstrap', 'v_bootstrap',
                'sub_sample_dt', 'dimension', 'ngp', 'covariance_matrix', 'distributions', 'diffusion_coefficient',
                'jump_diffusion_coefficient', 'sigma', 'intercept', 'gradient', 'start_dt', 'model', 'cond_max',
                'euclidian_displacements', 'flatchain'):
        assert key in d['diff'], key  # the keys of kinisi/diffusion.py:88-122
    b = A.from_dict(d)
    np.testing.assert_array_equal(a.dt, b.dt)
    np.testing.assert_array_equal(a.msd, b.msd)
    np.testing.assert_array_equal(a.msd_std, b.msd_std)
    assert a.ngp_max == b.ngp_max and a.volume == b.volume
    np.testing.assert_array_equal(a.D.samples, b.D.samples)
    np.testing.assert_array_equal(a.intercept.samples, b.intercept.samples)
    np.testing.assert_array_equal(a.flatchain, b.flatchain)
    np.testing.assert_array_equal(a._diff.covariance_matrix, b._diff.covariance_matrix)
    np.testing.assert_array_equal(a.distribution, b.distribution)
    np.testing.assert_array_equal(a._diff._n_bootstrap, b._diff._n_bootstrap)
    assert len(b._diff._distributions) == len(a._diff._distributions) > 0
    np.testing.assert_array_equal(a.dr[0].samples, b.dr[0].samples)
    assert b._diff.D_J is None and b._diff.sigma is None
    ppd = b.posterior_predictive({'n_posterior_samples': 8, 'n_predictive_samples': 4, 'progress': False})
    assert ppd.shape == (32, a.dt[start_idx:].size)
    base = kb.analyze.DiffusionAnalyzer.__mro__[1]
    plain = base.from_dict(base(a._delta_t, a._disp_3d, a._n_o, a._volume).to_dict())
    np.testing.assert_array_equal(plain._delta_t, a._delta_t)
    assert plain._volume == a._volume


def test_universe_front_end(kb, golden_dir):
    """MDAnalysisParser / from_universe on a duck-typed universe (tests/duck_atoms.py) against the reference run on the
    same object (tests/golden/universe_reference.npz; kinisi/parser.py:511-672, analyzer.py:236-271).  The Cartesian ->
    fractional product runs on the device in float64: parity to rounding for float64 input; for float32 input (what
    MDAnalysis hands out) the reference itself works in single precision, so only its noise level can be asked."""
    from duck_atoms import universe_from_arrays
    g = load(golden_dir, 'universe_reference')
    for name in ('ase_nvt', 'ase_npt_triclinic'):
        symbols, frac, latt, pp = ase_case_inputs(name)
        for tag, dtype, atol, rtol in (('f64', np.float64, 1e-11, 1e-9), ('f32', np.float32, 2e-5, 2e-4)):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                p = kb.parser.MDAnalysisParser(universe_from_arrays(symbols, frac, latt, dtype), **pp)
            key = f'{name}_{tag}_'
            np.testing.assert_allclose(p.dc, g[key + 'dc'], rtol=0, atol=atol)
            np.testing.assert_allclose(p.delta_t, g[key + 'delta_t'], rtol=1e-14)
            np.testing.assert_allclose(p._n_o, g[key + 'n_o'], rtol=1e-14)
            np.testing.assert_allclose(p.volume, g[key + 'volume'], rtol=1e-12)
            np.testing.assert_allclose(p.coords_check, g[key + 'coords_check'], rtol=0, atol=atol)
            b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
            check_boot(b, g, key + 'msd_', rtol=rtol, ngp_atol=1e-10 if dtype == np.float64 else 1e-5)
    symbols, frac, latt, pp = ase_case_inputs('ase_npt_triclinic')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        p = kb.parser.MDAnalysisParser(universe_from_arrays(symbols, frac, latt, np.float64), sub_sample_atoms=2,
                                       sub_sample_traj=3, **pp)
        np.testing.assert_allclose(p.dc, g['sub_dc'], rtol=0, atol=1e-11)
        np.testing.assert_allclose(p.delta_t, g['sub_delta_t'], rtol=1e-14)
        check_boot(kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False), g, 'sub_msd_')
        symbols, frac, latt, pp = ase_case_inputs('ase_nvt')
        u = universe_from_arrays(symbols, frac, latt, np.float64)
        p = kb.parser.MDAnalysisParser(u, **dict(pp, specie=None, specie_indices=[0, 2, 3, 5]))
        np.testing.assert_allclose(p.dc, g['idx_dc'], rtol=0, atol=1e-11)
        assert list(p.indices) == list(g['idx_indices']) and list(p.drift_indices) == list(g['idx_drift_indices'])
        with pytest.raises(ValueError, match='sub_sample_atom'):
            kb.parser.MDAnalysisParser(u, **dict(pp, specie=None, specie_indices=[0, 2], sub_sample_atoms=2))
        with pytest.raises(ValueError, match='no species'):
            kb.parser.MDAnalysisParser(u, **dict(pp, specie='Xe'))
        a = kb.analyze.DiffusionAnalyzer.from_universe(u, parser_params=pp, uncertainty_params={'progress': False})
        np.testing.assert_allclose(a.msd, g['ase_nvt_f64_msd_n'], rtol=1e-9)
        two = kb.analyze.DiffusionAnalyzer.from_universe([u, u], parser_params=pp, dtype='identical',
                                                         uncertainty_params={'progress': False})
        np.testing.assert_allclose(two.msd, a.msd, rtol=1e-12)  # the same trajectory twice: same mean, doubled n_o
        np.testing.assert_allclose(two._n_o, 2 * a._n_o, rtol=1e-14)
        with pytest.raises(ValueError, match='dtype'):
            kb.analyze.DiffusionAnalyzer.from_universe([u], parser_params=pp, dtype='consecutive')


def test_from_file_reads_xdatcar(kb, tmp_path):
    """`from_file` (kinisi/analyzer.py:199-234, diffusion_analyzer.py:148-175) through the built-in XDATCAR reader against
    `from_pymatgen` on the same configurations; a list of files with dtype='identical' and 'consecutive'."""
    from kinisi_b200.xdatcar import Xdatcar
    from test_host_logic import _write_xdatcar
    symbols, frac, latt, pp = ase_case_inputs('ase_npt_triclinic')
    order = np.argsort([0 if sym == 'Li' else 1 for sym in symbols], kind='stable')  # XDATCAR groups the species
    frac_f = np.round(np.transpose(frac[order], (1, 0, 2)), 9)
    latt_r = np.round(latt, 9)
    names = [symbols[i] for i in order]
    species = sorted(set(names), key=names.index)
    counts = [names.count(sp) for sp in species]
    path = tmp_path / 'XDATCAR'
    _write_xdatcar(path, species, counts, frac_f, latt_r, True)
    pp = dict(pp, time_step=2.0)  # the pymatgen front-end takes femtoseconds
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        a = kb.analyze.DiffusionAnalyzer.from_file(str(path), parser_params=pp, uncertainty_params={'progress': False})
        b = kb.analyze.DiffusionAnalyzer.from_pymatgen(Xdatcar(str(path)).structures, parser_params=pp,
                                                       uncertainty_params={'progress': False})
        c = kb.analyze.DiffusionAnalyzer.from_Xdatcar(Xdatcar(str(path)), parser_params=pp,
                                                      uncertainty_params={'progress': False})
        two = kb.analyze.DiffusionAnalyzer.from_file([str(path), str(path)], parser_params=pp, dtype='identical',
                                                     uncertainty_params={'progress': False})
        cons = kb.analyze.DiffusionAnalyzer.from_file([str(path), str(path)], parser_params=pp, dtype='consecutive',
                                                      uncertainty_params={'progress': False})
    for x in (b, c):
        np.testing.assert_allclose(a.msd, x.msd, rtol=1e-12)  # the moment sums are accumulated with atomics
        np.testing.assert_array_equal(a.dt, x.dt)
    np.testing.assert_allclose(two.msd, a.msd, rtol=1e-12)
    assert cons.dt.size >= a.dt.size and cons._diff._max_obs > a._diff._max_obs
    # the same numbers as the oracle on the arrays that were written
    coords = np.concatenate([frac_f[:1], frac_f], axis=0).transpose(1, 0, 2)
    cells = np.concatenate([latt_r[:1], latt_r], axis=0)
    li = [i for i, sym in enumerate(names) if sym == 'Li']
    fw = [i for i, sym in enumerate(names) if sym != 'Li']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        dc = ko.correct_drift(fw, ko.get_disp(coords, cells))
    k = int(round(a.dt[0] / (pp['time_step'] * 1e-3 * pp['step_skip'])))
    d = ko.displacements_for_lag(dc[li], k)
    np.testing.assert_allclose(a.msd[0], np.mean(np.sum(d**2, axis=-1)), rtol=1e-9)


def test_identical_trajectories_stack(kb):
    """dtype='identical' (kinisi/analyzer.py:111-116, 274-290): more atoms, n_o summed."""
    symbols = ['Li'] * 6 + ['O'] * 3
    trajs, parsed = [], []
    for seed in (1, 2):
        frac, latt = ko.synthetic_walk(6, 400, box=8.0, sigma=0.15, seed=seed, n_framework=3)
        trajs.append(trajectory_from_arrays(symbols, frac[:, 1:], latt[1:]))
        parsed.append(ko.parse(ko.get_disp(frac, latt), list(range(6)), [6, 7, 8], 0.001, 1, n_steps=30))
    pp = dict(specie='Li', time_step=0.001, step_skip=1, n_steps=30, progress=False)
    an = kb.analyze.DiffusionAnalyzer.from_ase(trajs, parser_params=pp, dtype='identical')
    joint = dict(parsed[0])
    joint['dc_sel'] = np.concatenate([parsed[0]['dc_sel'], parsed[1]['dc_sel']], axis=0)
    joint['n_o'] = parsed[0]['n_o'] + parsed[1]['n_o']
    ref = ko.bootstrap_streaming(joint, 'msd')
    assert an._disp_3d.shape_of(0)[0] == 12
    np.testing.assert_allclose(an.msd, ref['n'], rtol=RTOL_MOMENT)
    np.testing.assert_allclose(an._diff.v, ref['v'], rtol=RTOL_MOMENT)
    an2 = kb.analyze.DiffusionAnalyzer.from_ase(trajs, parser_params=pp, dtype='consecutive')
    assert an2._disp_3d.n_frames == 800
    with pytest.raises(ValueError):
        kb.analyze.DiffusionAnalyzer.from_ase(trajs, parser_params=pp, dtype='nonsense')


# ---------------------------------------------------------------------------------------------
# stage 3
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', list(REGRESS_CASES))
def test_covariance_likelihood_and_posterior(kb, golden_dir, name):
    g = load(golden_dir, name)
    symbols, frac, latt, pp, start_idx, kind = regress_case_inputs(name)
    traj = trajectory_from_arrays(symbols, frac, latt)
    cls = {'msd': kb.analyze.DiffusionAnalyzer, 'mstd': kb.analyze.JumpDiffusionAnalyzer,
           'mscd': kb.analyze.ConductivityAnalyzer}[kind]
    an = cls.from_ase(traj, parser_params=pp, uncertainty_params={'progress': False})
    check_boot(an._diff, g, '')
    start_dt = float(g['start_dt'])
    kw = {'progress': False, 'random_state': np.random.RandomState(3)}
    if kind == 'msd':
        an.diffusion(start_dt, kw)
        coeff = an.D
    elif kind == 'mstd':
        an.jump_diffusion(start_dt, kw)
        coeff = an.D_J
    else:
        an.conductivity(start_dt, 500.0, kw)
        coeff = an.sigma
    b = an._diff
    # covariance: <= 1e-6 relative (entries are all of the order of the variances)
    scale = np.abs(g['covariance_matrix']).max()
    np.testing.assert_allclose(b._npd_covariance_matrix, g['npd_covariance_matrix'], rtol=1e-9)
    np.testing.assert_allclose(b.covariance_matrix, g['covariance_matrix'], rtol=1e-6, atol=1e-9 * scale)
    assert b.covariance_matrix.shape == g['covariance_matrix'].shape
    # log-likelihood: the reference's logdet of a matrix pinned at condition number 1e16 is not
    # reproducible across LAPACK builds, so values are compared up to that additive constant
    ll = b.log_likelihood(g['thetas'])
    assert ll[-1] == -np.inf and g['loglike'][-1] == -np.inf
    d_mine = ll[1:-1] - ll[0]
    d_ref = g['loglike'][1:-1] - g['loglike'][0]
    np.testing.assert_allclose(d_mine, d_ref, rtol=1e-5, atol=1e-5)
    # the sampler's O(1) expansion of the quadratic form agrees with the O(n) whitened form
    chain = b._chain_dev[::97].reshape(-1, 2).cpu().numpy()
    logp = b._logp_dev[::97].reshape(-1).cpu().numpy()
    np.testing.assert_allclose(b.log_likelihood(chain), logp, rtol=1e-10, atol=1e-8)
    # flat chain layout (kinisi/tests/test_diffusion.py:280) and the posterior against the closed form
    assert an.flatchain.shape == (3200, 2)
    dr = int(np.argwhere(b.dt >= start_dt)[0][0])
    _, _, inv = ko.make_log_likelihood(b.dt[dr:], b.n[dr:], g['covariance_matrix'])
    mean, cov_theta = ko.gls_posterior(b.dt[dr:], b.n[dr:], inv)
    sd = np.sqrt(np.diag(cov_theta))
    grad, icpt = b.gradient.samples, b.intercept.samples
    # Monte-Carlo error of the mean of ~3200 correlated draws: allow 0.25 sigma; widths within 15 %
    assert abs(grad.mean() - mean[0]) < 0.25 * sd[0]
    assert abs(icpt.mean() - mean[1]) < 0.25 * sd[1]
    assert grad.std() == pytest.approx(sd[0], rel=0.15)
    assert icpt.std() == pytest.approx(sd[1], rel=0.15)
    # ... and against the reference's own draws (emcee stand-in, 640 samples)
    ref_grad = g['gradient_samples']
    assert abs(np.median(grad) - np.median(ref_grad)) < 0.35 * sd[0]
    lo, hi = np.percentile(grad, [2.5, 97.5])
    rlo, rhi = np.percentile(ref_grad, [2.5, 97.5])
    assert abs(lo - rlo) < 0.6 * sd[0] and abs(hi - rhi) < 0.6 * sd[0]
    # unit conversions (kinisi/diffusion.py:436, 456-457, 479-482)
    if kind == 'msd':
        np.testing.assert_allclose(coeff.samples, ko.diffusion_coefficient(grad, 3), rtol=1e-14)
    elif kind == 'mstd':
        np.testing.assert_allclose(coeff.samples, ko.jump_diffusion_coefficient(grad, 3, 32), rtol=1e-14)
    else:
        np.testing.assert_allclose(coeff.samples, ko.conductivity(grad, 3, an.volume, 500.0), rtol=1e-14)
    assert an.distribution.shape == (b.dt.size, 3200)
    ppd = an.posterior_predictive({'n_posterior_samples': 16, 'n_predictive_samples': 8, 'progress': False})
    assert ppd.shape == (128, b.dt.size - dr)


def test_long_grid_regression_takes_the_library_eigensolver(kb):
    """A 300-point dt grid (BASELINE configs[3] in small): the moment kernel runs in two 192-lag launches and the
    regression has n > 128, so the spectrum comes from the library eigensolver + kb2_gls_prepare instead of the Jacobi
    kernel.  Moments, covariance, the quadratic form and the posterior against the oracle."""
    n_atoms, n_frames, n_steps = 20, 4000, 300
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=12.0, sigma=0.1, seed=21)
    dc = ko.get_disp(frac, latt)
    parsed = ko.parse(dc, list(range(n_atoms)), [], 0.002, 1, n_steps=n_steps)
    ref = ko.bootstrap_streaming(parsed, 'msd')
    p = kb.parser.Parser(kb.parser.Parser.get_disp(frac, latt), list(range(n_atoms)), [], 0.002, 1, n_steps=n_steps,
                         progress=False)
    b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    for key in ('dt', 'n', 'v'):
        np.testing.assert_allclose(getattr(b, key), ref[key], rtol=RTOL_MOMENT, err_msg=key)
    start = 40
    b.diffusion(float(b.dt[start]), progress=False, random_state=np.random.RandomState(5))
    n_dt = len(b.dt) - start
    assert n_dt > kb.engine.spectral_prepare_max_n()
    cov_ref = ko.generate_covariance_matrix(ref['v'], ref['n_o'], start)
    scale = np.abs(cov_ref).max()
    np.testing.assert_allclose(b.covariance_matrix, cov_ref, rtol=1e-6, atol=1e-9 * scale)
    _, _, inv = ko.make_log_likelihood(ref['dt'][start:], ref['n'][start:], cov_ref)
    thetas = np.array([[0.18, 0.0], [0.17, 0.05], [0.19, -0.04], [0.2, 0.1]])
    x, y = ref['dt'][start:], ref['n'][start:]
    q_ref = np.array([(t[0] * x + t[1] - y) @ inv @ (t[0] * x + t[1] - y) for t in thetas])
    ll = b.log_likelihood(thetas)
    np.testing.assert_allclose(-2.0 * (ll - ll[0]), q_ref - q_ref[0], rtol=1e-5, atol=1e-5 * np.abs(q_ref).max())
    mean, cov_theta = ko.gls_posterior(x, y, inv)
    sd = np.sqrt(np.diag(cov_theta))
    assert b.flatchain.shape == (3200, 2)
    assert abs(b.gradient.samples.mean() - mean[0]) < 0.25 * sd[0]
    assert b.gradient.samples.std() == pytest.approx(sd[0], rel=0.15)


def test_model_variance(kb):
    """model=True (kinisi/diffusion.py:406-420): the covariance is built from the fitted variances a / n_o * dt**2;
    the closed-form fit against the oracle's scipy curve_fit, on both spectral routes."""
    n_atoms, n_frames = 24, 3000
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=12.0, sigma=0.1, seed=31)
    parsed = ko.parse(ko.get_disp(frac, latt), list(range(n_atoms)), [], 0.002, 1, n_steps=60)
    ref = ko.bootstrap_streaming(parsed, 'msd')
    p = kb.parser.Parser(kb.parser.Parser.get_disp(frac, latt), list(range(n_atoms)), [], 0.002, 1, n_steps=60,
                         progress=False)
    start = 8
    cov_ref = ko.generate_covariance_matrix(ref['v'], ref['n_o'], start, dt=ref['dt'], model=True)
    scale = np.abs(cov_ref).max()
    for literal in (False, True):
        kb.diffusion.LITERAL_RECONDITIONING = literal
        try:
            b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
            b.diffusion(float(b.dt[start]), model=True, progress=False, n_samples=200, n_burn=100)
            np.testing.assert_allclose(b.covariance_matrix, cov_ref, rtol=1e-6, atol=1e-9 * scale)
        finally:
            kb.diffusion.LITERAL_RECONDITIONING = False
        assert b.D.n > 0 and b.flatchain.shape == (32 * 20, 2)


def test_fast_reconditioning_matches_the_literal_sequence(kb, golden_dir):
    """The default stage-3 path factorises the covariance once; the literal path repeats the reference's
    eigh / cov_nearest / slogdet / pinvh sequence.  Both must give the same matrix and likelihood."""
    g = load(golden_dir, 'regress_msd')
    symbols, frac, latt, pp, start_idx, kind = regress_case_inputs('regress_msd')
    traj = trajectory_from_arrays(symbols, frac, latt)
    out = {}
    for literal in (False, True):
        kb.diffusion.LITERAL_RECONDITIONING = literal
        try:
            an = kb.analyze.DiffusionAnalyzer.from_ase(traj, parser_params=pp, uncertainty_params={'progress': False})
            an.diffusion(float(g['start_dt']), {'progress': False, 'n_samples': 100, 'n_burn': 50})
        finally:
            kb.diffusion.LITERAL_RECONDITIONING = False
        ll = an._diff.log_likelihood(g['thetas'][:-1])
        out[literal] = (an._diff.covariance_matrix, ll - ll[0])
    scale = np.abs(out[True][0]).max()
    np.testing.assert_allclose(out[False][0], out[True][0], rtol=1e-9, atol=1e-12 * scale)
    np.testing.assert_allclose(out[False][1], out[True][1], rtol=1e-6, atol=1e-6)
    d_ref = g['loglike'][:-1] - g['loglike'][0]
    np.testing.assert_allclose(out[True][1], d_ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('n', [1, 2, 3, 4, 5, 6, 7, 8, 17, 54, 90, 91, 92, 93, 94, 127, 128])
def test_jacobi_spectral_kernel(kb, n):
    """kb2_spectral_prepare against numpy: spectrum, log-determinant, and the quadratic form of pinvh."""
    from scipy.linalg import pinvh
    rng = np.random.RandomState(n)
    # a model-covariance-like matrix (kinisi/diffusion.py:838-854): not positive definite in general
    v = rng.uniform(0.5, 2.0, n) * np.linspace(1, 30, n)**2
    n_o = 1000.0 / np.arange(1, n + 1)
    cov = ko.populate_covariance_matrix(v, n_o)
    x = np.linspace(0.1, 9.0, n)
    y = 3.0 * x + 0.2 + rng.randn(n) * np.sqrt(v)
    dev = torch.device('cuda')
    evals, proj, gls, logdet_term, info = kb.engine.spectral_prepare(torch.from_numpy(cov).to(dev),
                                                                     torch.from_numpy(x).to(dev),
                                                                     torch.from_numpy(y).to(dev), 1e16, True)
    w = np.linalg.eigvalsh(cov)
    np.testing.assert_allclose(np.sort(evals.cpu().numpy()), w, rtol=0, atol=1e-13 * np.abs(w).max())
    assert 0 <= int(info.item()) <= 20
    # reference route: recondition, then pinvh (the oracle restates diffusion.py:397-425, 350-365)
    cov2 = ko.generate_covariance_matrix(v, n_o, 0)
    inv = pinvh(cov2)
    thetas = np.array([[3.0, 0.2], [2.5, 0.5], [3.4, -0.3], [0.1, 4.0]])
    ll = kb.engine.mvn_loglike(torch.from_numpy(thetas).to(dev), proj, logdet_term).cpu().numpy()
    q_ref = np.array([(t[0] * x + t[1] - y) @ inv @ (t[0] * x + t[1] - y) for t in thetas])
    q_mine = -2.0 * ll - float(logdet_term.item())
    np.testing.assert_allclose(q_mine - q_mine[0], q_ref - q_ref[0], rtol=1e-6, atol=1e-6 * np.abs(q_ref).max())
    # the GLS optimum and the O(1) expansion about it
    g = gls.cpu().numpy()
    q_exp = g[7] + 2 * ((thetas[:, 0] - g[0]) * g[5] + (thetas[:, 1] - g[1]) * g[6]) + (thetas[:, 0] - g[0])**2 * g[2] \
        + 2 * (thetas[:, 0] - g[0]) * (thetas[:, 1] - g[1]) * g[3] + (thetas[:, 1] - g[1])**2 * g[4]
    np.testing.assert_allclose(q_exp, q_mine, rtol=1e-9, atol=1e-9 * np.abs(q_mine).max())
    if n > 2:
        mean, _ = ko.gls_posterior(x, y, inv)
        if mean[0] > 0:
            np.testing.assert_allclose(g[:2], mean, rtol=1e-5, atol=1e-7)


def test_sampler_sizes_and_options(kb):
    """Output sizes as pinned by kinisi/tests/test_diffusion.py:280, 318, 331, 357."""
    dt, disp_3d, n_o, _ = list_bootstrap_inputs()
    b = kb.diffusion.MSDBootstrap(dt, disp_3d, n_o, progress=False)
    b.diffusion(0, progress=False)
    assert b.covariance_matrix.shape == (10, 10) and b.gradient.size == 3200 and b.intercept.size == 3200
    assert b.D.size == 3200
    b.diffusion(0, thin=1, progress=False)
    assert b.gradient.size == 32000
    b.diffusion(0, n_samples=100, fit_intercept=False, progress=False)
    assert b.gradient.size == 320 and b.intercept is None
    b.diffusion(0, n_walkers=16, progress=False)
    assert b.gradient.size == 1600
    b.diffusion(400, progress=False)
    assert b.covariance_matrix.shape == (7, 7)  # start_dt, test_diffusion.py:290
    with pytest.raises(IndexError):
        b.diffusion(1e9, progress=False)
    assert np.all(b.gradient.samples >= 0)


# ---------------------------------------------------------------------------------------------
# error behaviour of the boundary
# ---------------------------------------------------------------------------------------------
def test_errors(kb):
    disp = parser_random_disp()
    with pytest.raises(ValueError):
        kb.parser.Parser(disp, np.arange(100), [], 1, 1, min_dt=100, progress=False)
    with pytest.raises(ValueError):
        kb.parser.Parser(disp, np.arange(100), [], 1, 1, spacing='cubic', progress=False)
    with pytest.raises(ValueError):
        kb.parser.Parser(disp, np.arange(100), [], 1, 1, sampling='all-origins', progress=False)
    with pytest.raises(MemoryError):
        kb.parser.Parser(disp, np.arange(100), [], 1, 1, memory_limit=1e-9, progress=False)
    symbols, frac, latt, pp = ase_case_inputs('ase_nvt')
    traj = trajectory_from_arrays(symbols, frac, latt)
    with pytest.raises(ValueError):
        kb.parser.ASEParser(traj, specie='Na', time_step=1, step_skip=1, progress=False)
    with pytest.raises(TypeError):
        kb.parser.ASEParser(traj, specie=None, time_step=1, step_skip=1, progress=False)
    with pytest.warns(UserWarning):
        s2, f2, l2, pp2 = ase_case_inputs('ase_npt_triclinic')
        kb.parser.ASEParser(trajectory_from_arrays(s2, f2, l2), **pp2)
    p = kb.parser.Parser(disp, np.arange(100), [], 1, 1, sampling='single-origin', progress=False)
    b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    ref = np.array([np.mean(np.sum(disp[:, i]**2, axis=-1)) for i in range(b.n.size)])
    np.testing.assert_allclose(b.n, ref, rtol=1e-12)  # single-origin takes frame i (kinisi/parser.py:203)


# ---------------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]): no oracle at this size, so size-independent checks
# ---------------------------------------------------------------------------------------------
def test_full_size_stage1_against_torch_fp64(kb):
    """BASELINE configs[1] shape (1536 atoms, 448 mobile, 20 000 frames, f32 wrapped coordinates): stage 1 against a
    plain PyTorch float64 restatement of kinisi/parser.py:101-148 on the device.  Exercises 16 tiles per atom of the
    persistent unwrap kernel (totals exchanged between CTAs) and the staged ingest from pinned host memory."""
    n_mob, n_fw, n_frames, box = 448, 1088, 20000, 26.0
    dev = torch.device('cuda')
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    steps = torch.randn((n_mob + n_fw, n_frames + 1, 3), dtype=torch.float64, device=dev, generator=gen) * (0.1 / box)
    steps[:, 0] = torch.rand((n_mob + n_fw, 3), dtype=torch.float64, device=dev, generator=gen)
    steps[:, 1] = 0.0  # the front-ends duplicate the first frame
    frac32 = (torch.cumsum(steps, dim=1) % 1.0).to(torch.float32)
    latt = np.eye(3)[None] * box
    idx, fw = np.arange(n_mob), np.arange(n_mob, n_mob + n_fw)
    # reference: float64 torch ops on the same float32 input
    f64 = frac32.to(torch.float64)
    d = f64[:, 1:] - f64[:, :-1]
    u = (d - torch.floor(d + 0.5)) * box
    dc_all = torch.cumsum(u, dim=1)
    ref = (dc_all[:n_mob] - dc_all[n_mob:].mean(dim=0, keepdim=True)).permute(0, 2, 1)  # [atom, axis, frame]
    scale = float(ref.abs().max())
    for source in ('device', 'pinned host'):
        coords = frac32 if source == 'device' else frac32.cpu().pin_memory()
        p = kb.parser.Parser(kb.parser.Parser.get_disp(coords, latt), idx, fw, 0.001, 1, n_steps=100, memory_limit=64.,
                             progress=False)
        dc = p.disp_3d.dc[:, :, :n_frames]
        err = float((dc - ref).abs().max())
        assert err <= 1e-11 * scale, (source, err, scale)
        pad = p.disp_3d.dc[:, :, n_frames:]
        assert pad.numel() == 0 or float(pad.abs().max()) == 0.0  # the row padding stays zero


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('drift', [True, False])
@pytest.mark.parametrize('shape', [(5, 3, 40), (37, 9, 223), (64, 8, 224), (70, 11, 225), (33, 7, 1571), (1000, 30, 301),
                                   (2100, 12, 449), (311, 5, 1280), (40, 6, 1281), (9, 4, 6431)])
def test_unwrap_kernels_agree(kb, monkeypatch, dtype, drift, shape):
    """The three unwrap kernels -- the bulk-copy tile kernel (producer warp, TMA loads and stores; the default: 1280-frame
    tiles for float32, 640-frame tiles for float64), the one-warp-per-atom kernel (bulk-copy ring, scan carried in registers) and the cp.async tile kernel --
    against each other and against float64 torch: rows at every misalignment (odd frame counts), tiles and sub-tiles that
    end inside the trajectory, one and several tiles per atom (look-back), more atoms than warps (a warp walks several
    atoms without draining its ring), a scattered selection that includes the last row of the array (the bytes a 16-byte
    copy would overrun), with and without the framework drift, and a zero row pad."""
    from kinisi_b200._lib import MODE_FRACTIONAL
    n_sel, n_fw, n_frames = shape
    dev = torch.device('cuda')
    gen = torch.Generator(device=dev)
    gen.manual_seed(11 + n_frames)
    n_all = n_sel + n_fw + 3
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    steps = torch.randn((n_all, n_frames + 1, 3), dtype=torch.float64, device=dev, generator=gen) * 0.05
    steps[:, 0] = torch.rand((n_all, 3), dtype=torch.float64, device=dev, generator=gen)
    coords = (torch.cumsum(steps, dim=1) % 1.0).to(tdt).contiguous()
    cell = np.array([[9.0, 0.3, 0.0], [0.0, 8.0, 0.2], [0.4, 0.0, 10.0]])
    latt = torch.tensor(cell, dtype=torch.float64, device=dev).reshape(1, 3, 3)
    inv = kb.engine.invert_lattice(latt)
    perm = torch.randperm(n_all, device=dev, generator=gen)
    sel = perm[:n_sel].to(torch.int32)
    sel[0] = n_all - 1  # the last row of the coordinate array
    fw = perm[n_sel:n_sel + n_fw].to(torch.int32).contiguous()
    pitch = kb.engine.pitch_for(n_frames) + (32 if n_frames % 2 else 0)  # sometimes a pad longer than a 16-frame round-up
    sums = kb.engine.frame_sums(coords, MODE_FRACTIONAL, latt, inv, 0, fw, None, pitch) if drift else None
    scale = 1.0 / n_fw if drift else 0.0
    got = {}
    modes = ('2', '1', '0')
    for mode in modes:
        monkeypatch.setenv('KB2_UW_KERNEL', mode)
        out = torch.full((n_sel, 3, pitch), float('nan'), dtype=torch.float64, device=dev)
        got[mode] = kb.engine.unwrap_cumsum(coords, MODE_FRACTIONAL, latt, inv, 0, sel, sums, scale, pitch, out=out).clone()
    monkeypatch.delenv('KB2_UW_KERNEL')
    f64 = coords.to(torch.float64)[sel.long()]
    d = f64[:, 1:] - f64[:, :-1]
    u = (d - torch.floor(d + 0.5)) @ latt[0]
    if drift:
        u = u - sums[:, :n_frames].T[None] * scale
    ref = torch.cumsum(u, dim=1).permute(0, 2, 1)
    tol = 1e-12 * float(ref.abs().max()) * max(1.0, n_frames / 200)
    for mode in modes:
        assert not torch.isnan(got[mode]).any(), mode
        assert float((got[mode][:, :, :n_frames] - ref).abs().max()) <= tol, mode
        assert pitch == n_frames or float(got[mode][:, :, n_frames:].abs().max()) == 0.0, mode
        assert float((got[mode] - got['0']).abs().max()) <= tol, mode


def test_full_size_properties(kb):
    n_atoms, n_frames = 448, 20000
    dev = torch.device('cuda')
    gen = torch.Generator(device=dev)
    gen.manual_seed(1002)
    steps = torch.randn((n_atoms, 3, n_frames), dtype=torch.float64, device=dev, generator=gen) * 0.1
    steps[:, :, 0] = 0
    pitch = kb.engine.pitch_for(n_frames)
    dc = torch.zeros((n_atoms, 3, pitch), dtype=torch.float64, device=dev)
    dc[:, :, :n_frames] = torch.cumsum(steps, dim=2)
    lags = np.linspace(1, n_frames, 100, dtype=int)
    lt = torch.as_tensor(lags.astype(np.int32)).to(dev)
    a = kb.engine.self_moments(dc, n_frames, lt, 7, 0).cpu().numpy()
    b = kb.engine.self_moments(dc, n_frames, lt, 7, 1).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-12)  # the kernel variants agree
    r = kb.engine.self_moments(dc, n_frames, lags, 7, kb.engine.VARIANT_RUNS).cpu().numpy()
    np.testing.assert_allclose(r, a, rtol=1e-12)
    # scaling: positions x 2 -> sums x 4 and x 16 (exact in binary floating point)
    c = kb.engine.self_moments(dc * 2.0, n_frames, lt, 7, 1).cpu().numpy()
    # (exact per term; the reduction order differs from launch to launch, hence a tolerance)
    np.testing.assert_allclose(c[:, 0], 4.0 * b[:, 0], rtol=1e-13)
    np.testing.assert_allclose(c[:, 1], 16.0 * b[:, 1], rtol=1e-13)
    # components add up: sum d^2 over xyz = x + y + z
    parts = sum(kb.engine.self_moments(dc, n_frames, lt, m, 1).cpu().numpy()[:, 0] for m in (1, 2, 4))
    np.testing.assert_allclose(parts, b[:, 0], rtol=1e-12)
    # atoms are independent: halves add up
    h1 = kb.engine.self_moments(dc[:224].contiguous(), n_frames, lt, 7, 1).cpu().numpy()
    h2 = kb.engine.self_moments(dc[224:].contiguous(), n_frames, lt, 7, 1).cpu().numpy()
    np.testing.assert_allclose(h1 + h2, b, rtol=1e-12)
    # a Gaussian walk: MSD slope 3 sigma^2 per frame, NGP -> 0 at short lags
    cnt = n_atoms * (n_frames - lags + 1.0)
    msd = b[:, 0] / cnt
    short = lags < 2500
    slope = np.polyfit(lags[short], msd[short], 1)[0]
    assert slope == pytest.approx(3 * 0.1**2, rel=0.05)
    ngp = 3 * (b[:, 1] / cnt) / (5 * msd**2) - 1
    assert np.all(np.abs(ngp[short][1:]) < 0.05)


# ---------------------------------------------------------------------------------------------
# the CUDA path against the oracle at BASELINE sizes ("What's weak" 2)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('variant', ALL_VARIANTS)
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_baseline_length_slice_against_oracle(kb, variant, dtype, monkeypatch):
    """BASELINE configs[1] in length and lag grid (20 000 frames, the real 100-lag grid, framework drift), on an atom
    slice the streaming oracle does in seconds: MSD, MSTD and MSCD through the classes, every moment-kernel variant,
    f32 and f64 wrapped coordinates.  The same terms as the reference (kinisi/diffusion.py:571-602, 655-688, 739-780)."""
    monkeypatch.setattr(kb.engine, 'DEFAULT_VARIANT', variant)
    n_mobile, n_fw, n_frames = 32, 64, 20000
    frac, latt = ko.synthetic_walk(n_mobile, n_frames, box=26.0, sigma=0.1, seed=1001, n_framework=n_fw, dtype=dtype)
    idx, fw = list(range(n_mobile)), list(range(n_mobile, n_mobile + n_fw))
    parsed = ko.parse(ko.get_disp(frac.astype(np.float64), latt), idx, fw, 0.001, 1, n_steps=100)
    p = kb.parser.Parser(kb.parser.Parser.get_disp(frac, latt[:1]), idx, fw, 0.001, 1, n_steps=100, memory_limit=64.,
                         progress=False)
    np.testing.assert_array_equal(p.time_intervals, parsed['lags'])
    np.testing.assert_allclose(p._n_o, parsed['n_o'], rtol=1e-15)
    q = np.where(np.arange(n_mobile) % 2 == 0, 1.0, -1.0)
    for kind, cls, args in (('msd', kb.diffusion.MSDBootstrap, ()), ('mstd', kb.diffusion.MSTDBootstrap, ()),
                            ('mscd', kb.diffusion.MSCDBootstrap, (q, ))):
        b = cls(p.delta_t, p.disp_3d, *args, p._n_o, progress=False)
        ref = ko.bootstrap_streaming(parsed, kind, q if kind == 'mscd' else None)
        for key in ('dt', 'n', 'v', 's'):
            np.testing.assert_allclose(getattr(b, key), ref[key], rtol=RTOL_MOMENT, err_msg=f'{kind} {key}')
        np.testing.assert_allclose(b.ngp, ref['ngp'], rtol=1e-8, atol=1e-10, err_msg=kind)


@pytest.mark.parametrize('variant', ALL_VARIANTS)
@pytest.mark.parametrize('n_frames', [50000, 100000])
def test_production_lag_grids_against_oracle(kb, variant, n_frames):
    """`np.linspace(1, T, 100, dtype=int)` at T = 50 000 (four progressions of 25 lags) and T = 100 000 (one progression
    with a one-sample carry every 11th lag): BASELINE configs[4] / configs[2] in length.  Raw per-lag sums of every
    kernel variant against `raw_sums_streaming` (kinisi/parser.py:208-213 + diffusion.py:572-574), two dimensions."""
    n_atoms = 3
    rng = np.random.default_rng(n_frames)
    dc = np.cumsum(rng.normal(0, 0.1, size=(n_atoms, n_frames, 3)), axis=1)
    dc[:, 0] = 0.0
    lags = np.linspace(1, n_frames, 100, dtype=int)
    dct = _device_dc(kb, dc)
    for mask, dim in ((7, 'xyz'), (3, 'xy')):
        S1, S2, _, _ = ko.raw_sums_streaming(dc, lags, dim)
        sums = kb.engine.self_moments(dct, n_frames, lags, mask, variant).cpu().numpy()
        np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
        np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)


AUTO_GRIDS = {  # name -> (n_frames, lags, the kernel kb2_self_moments_choose takes)
    'default grid, 20 000 frames': (20000, np.linspace(1, 20000, 100, dtype=int), 4),
    'default grid, 5000 frames': (5000, np.linspace(1, 5000, 100, dtype=int), 3),
    'dense grid (configs[3])': (20000, np.linspace(1, 20000, 2000, dtype=int)[:1000], 2),
    '400 points, spacing drifting every 8th lag': (20000, np.linspace(1, 20000, 400, dtype=int), 0),
    'logarithmic': (20000, np.unique(np.geomspace(1, 20000, 100).astype(int)), 0),
}


@pytest.mark.parametrize('grid', sorted(AUTO_GRIDS))
def test_kernel_chosen_by_the_grid_against_oracle(kb, grid):
    """variant=None (VARIANT_AUTO, what the classes and bench.py run): the kernel kb2_self_moments_choose takes for the
    grid gives the oracle's raw sums (kinisi/parser.py:208-213 + diffusion.py:572-574), and so does every other kernel
    on the same grid -- the choice changes the speed, never the terms."""
    n_frames, lags, want = AUTO_GRIDS[grid]
    assert kb.engine.DEFAULT_VARIANT == kb.engine.VARIANT_AUTO
    assert kb.engine.choose_variant(lags, n_frames) == want
    n_atoms = 2
    rng = np.random.default_rng(len(lags) + n_frames)
    dc = np.cumsum(rng.normal(0, 0.1, size=(n_atoms, n_frames, 3)), axis=1)
    dc[:, 0] = 0.0
    S1, S2, _, _ = ko.raw_sums_streaming(dc, lags)
    dct = _device_dc(kb, dc)
    before = kb.engine.LAUNCHES['n']
    sums = kb.engine.self_moments(dct, n_frames, lags, 7).cpu().numpy()
    assert kb.engine.LAUNCHES['n'] > before
    np.testing.assert_allclose(sums[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(sums[:, 1], S2, rtol=1e-12)
    for variant in (0, 2, 3, 4):
        if variant != want:
            other = kb.engine.self_moments(dct, n_frames, lags, 7, variant).cpu().numpy()
            np.testing.assert_allclose(other, sums, rtol=1e-12)


@pytest.mark.parametrize('n_frames,n_steps', [(20000, 100), (3333, 100), (8000, 250), (707, 20)])
def test_tile_kernel_packs_ragged_strips(kb, n_frames, n_steps):
    """Row lengths with a short last strip (D mod 32 <= 12: 202 at BASELINE configs[1], 101 / 202 / 808, 257, 74 here) put
    that strip of 2-4 consecutive atoms into one tile (csrc/k2_moments_tile.cu, tile_pack_atoms).  Atom counts around the
    pair and the L2-batch boundaries (one atom, an odd count, 50 + 1 at 20 000 frames), three dimension masks: the sums are
    the oracle's (kinisi/parser.py:208-213 + diffusion.py:572-574) for the small counts and the direct kernel's for all."""
    lags = np.unique(np.linspace(1, n_frames, n_steps, dtype=int))
    wins, _ = kb.engine.self_moments_tiles(lags, n_frames)
    assert any(w[2] > 1 and 0 < w[1] % 32 <= 12 for w in wins), 'no ragged row length in this grid'
    rng = np.random.default_rng(n_frames + n_steps)
    for n_atoms in (1, 2, 3, 51):
        dc = np.cumsum(rng.normal(0, 0.1, size=(n_atoms, n_frames, 3)), axis=1)
        dc[:, 0] = 0.0
        dct = _device_dc(kb, dc)
        for mask, dim in ((7, 'xyz'), (5, 'xz'), (2, 'y')):
            tiled = kb.engine.self_moments(dct, n_frames, lags, mask, kb.engine.VARIANT_TILE).cpu().numpy()
            direct = kb.engine.self_moments(dct, n_frames, lags, mask, kb.engine.VARIANT_DIRECT).cpu().numpy()
            np.testing.assert_allclose(tiled, direct, rtol=1e-12, err_msg=f'{n_atoms} atoms, {dim}')
            if n_atoms <= 3 and mask == 7:
                S1, S2, _, _ = ko.raw_sums_streaming(dc, lags, dim)
                np.testing.assert_allclose(tiled[:, 0], S1, rtol=1e-12)
                np.testing.assert_allclose(tiled[:, 1], S2, rtol=1e-12)


def test_full_size_default_kernel_against_the_direct_kernel_and_an_oracle_slice(kb):
    """BASELINE configs[1] at full size (448 atoms x 20 000 frames x 100 lags): the shipped default kernel against the
    direct kernel on all atoms, and both against the oracle's sums of the first 16 atoms (additivity over atoms)."""
    n_atoms, n_frames = 448, 20000
    rng = np.random.default_rng(1002)
    dc = np.cumsum(rng.normal(0, 0.1, size=(n_atoms, n_frames, 3)), axis=1)
    dc[:, 0] = 0.0
    lags = np.linspace(1, n_frames, 100, dtype=int)
    dct = _device_dc(kb, dc)
    default = kb.engine.self_moments(dct, n_frames, lags, 7).cpu().numpy()
    direct = kb.engine.self_moments(dct, n_frames, lags, 7, kb.engine.VARIANT_DIRECT).cpu().numpy()
    np.testing.assert_allclose(default, direct, rtol=1e-12)
    S1, S2, _, _ = ko.raw_sums_streaming(dc[:16], lags)
    head = kb.engine.self_moments(dct[:16].contiguous(), n_frames, lags, 7).cpu().numpy()
    tail = kb.engine.self_moments(dct[16:].contiguous(), n_frames, lags, 7).cpu().numpy()
    np.testing.assert_allclose(head[:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(head[:, 1], S2, rtol=1e-12)
    np.testing.assert_allclose(head + tail, default, rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# log-determinant (a10) and posterior predictive (f1)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [2, 5, 31, 64, 90, 128, 200])
def test_log_determinant_and_absolute_log_likelihood(kb, n):
    """`np.linalg.slogdet` of the reconditioned covariance (kinisi/diffusion.py:350-351).  Well-conditioned positive
    definite input: the log-determinant to 1e-10 and ABSOLUTE log-likelihood values against the oracle's
    slogdet + pinvh closure (:354-365).  Model covariance (indefinite): the clipped-spectrum sum
    sum_r log max(lambda_r, lambda_max / cond_max) (:883-887).  n <= 128 takes the Jacobi kernel, above it the
    library eigensolver + kb2_gls_prepare through the Bootstrap."""
    rng = np.random.RandomState(100 + n)
    dev = torch.device('cuda')
    x = np.linspace(0.1, 9.0, n)
    a = rng.randn(n, n)
    cov = a @ a.T / n + np.diag(rng.uniform(0.5, 1.5, n))  # condition number ~ 10
    y = 3.0 * x + 0.2 + rng.multivariate_normal(np.zeros(n), cov)
    thetas = np.array([[3.0, 0.2], [2.5, 0.5], [3.4, -0.3], [0.1, 4.0]])
    sign, logdet = np.linalg.slogdet(cov)
    assert sign > 0
    ll_fn, _, _ = ko.make_log_likelihood(x, y, cov)
    ll_ref = np.array([ll_fn(t) for t in thetas])
    if n <= kb.engine.spectral_prepare_max_n():
        evals, proj, gls, logdet_term, info = kb.engine.spectral_prepare(torch.from_numpy(cov).to(dev),
                                                                         torch.from_numpy(x).to(dev),
                                                                         torch.from_numpy(y).to(dev), 1e16, True)
    else:
        lam, vec = torch.linalg.eigh(torch.from_numpy(cov).to(dev))
        logdet_term = (torch.log(lam).sum() + float(np.log(2 * np.pi) * n)).reshape(1)
        proj, gls = kb.engine.gls_prepare(lam, vec, torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), True)
    np.testing.assert_allclose(float(logdet_term.item()) - n * np.log(2 * np.pi), logdet, rtol=1e-10, atol=1e-10)
    ll = kb.engine.mvn_loglike(torch.from_numpy(thetas).to(dev), proj, logdet_term).cpu().numpy()
    np.testing.assert_allclose(ll, ll_ref, rtol=1e-9, atol=1e-9)
    if n <= kb.engine.spectral_prepare_max_n():
        # an indefinite model covariance: the spectrum is clipped at lambda_max / cond_max before the logarithm
        v = rng.uniform(0.5, 2.0, n) * np.linspace(1, 30, n)**2
        n_o = 1000.0 / np.arange(1, n + 1)
        cov = ko.populate_covariance_matrix(v, n_o)
        w = np.linalg.eigvalsh(cov)
        for cond_max in (1e16, 1e8):
            _, _, _, logdet_term, _ = kb.engine.spectral_prepare(torch.from_numpy(cov).to(dev), torch.from_numpy(x).to(dev),
                                                                 torch.from_numpy(y).to(dev), cond_max, True)
            want = np.sum(np.log(np.maximum(w, w.max() / cond_max)))
            # an eigenvalue within rounding of the clip level may fall on either side of it: |d log| <= eps * cond
            np.testing.assert_allclose(float(logdet_term.item()) - n * np.log(2 * np.pi), want, rtol=1e-9, atol=1e-6)


def test_log_likelihood_values_of_a_well_conditioned_analysis(kb):
    """`Bootstrap.log_likelihood` is an API output: on a trajectory whose model covariance is positive definite and well
    conditioned, its VALUES (not differences) against the oracle's closure over the oracle's own covariance."""
    n_atoms, n_frames = 64, 4000
    frac, latt = ko.synthetic_walk(n_atoms, n_frames, box=12.0, sigma=0.1, seed=77)
    parsed = ko.parse(ko.get_disp(frac, latt), list(range(n_atoms)), [], 0.002, 1, n_steps=12)
    ref = ko.bootstrap_streaming(parsed, 'msd')
    p = kb.parser.Parser(kb.parser.Parser.get_disp(frac, latt), list(range(n_atoms)), [], 0.002, 1, n_steps=12,
                         progress=False)
    b = kb.diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    start = 2
    b.diffusion(float(b.dt[start]), progress=False, n_samples=200, n_burn=100)
    cov_ref = ko.generate_covariance_matrix(ref['v'], ref['n_o'], start)
    w = np.linalg.eigvalsh(ko.populate_covariance_matrix(ref['v'][start:], ref['n_o'][start:]))
    if w.min() <= 0 or w.max() / w.min() > 1e10:
        pytest.skip('the model covariance of this walk is not well conditioned')
    ll_fn, _, _ = ko.make_log_likelihood(ref['dt'][start:], ref['n'][start:], cov_ref)
    thetas = np.array([[0.18, 0.0], [0.17, 0.05], [0.19, -0.04], [0.2, 0.1]])
    np.testing.assert_allclose(b.log_likelihood(thetas), [ll_fn(t) for t in thetas], rtol=1e-7, atol=1e-7)


def test_posterior_predictive_distribution(kb, golden_dir):
    """`posterior_predictive` (kinisi/diffusion.py:492-521) against the oracle's restatement of the
    reference loop (scipy multivariate_normal per posterior draw): the drawn posterior samples are the same (same
    global-RNG call), per-draw means sit on the straight line, the pooled residual covariance is the reconditioned
    covariance, and mean / variance per dt agree with the reference loop within Monte-Carlo error."""
    g = load(golden_dir, 'regress_msd')
    symbols, frac, latt, pp, start_idx, kind = regress_case_inputs('regress_msd')
    an = kb.analyze.DiffusionAnalyzer.from_ase(trajectory_from_arrays(symbols, frac, latt), parser_params=pp,
                                               uncertainty_params={'progress': False})
    an.diffusion(float(g['start_dt']), {'progress': False, 'random_state': np.random.RandomState(11)})
    b = an._diff
    dr = int(np.argwhere(b.dt >= float(g['start_dt']))[0][0])
    x = b.dt[dr:]
    P, M = 48, 400
    np.random.seed(5)
    ppd = an.posterior_predictive({'n_posterior_samples': P, 'n_predictive_samples': M, 'progress': False})
    assert ppd.shape == (P * M, x.size)
    cov = b.covariance_matrix
    ref, drawn = ko.posterior_predictive(b.gradient.samples, b.intercept.samples, x, cov, P, M, seed=5)
    mu = b.gradient.samples[drawn][:, None] * x[None, :] + b.intercept.samples[drawn][:, None]
    sd = np.sqrt(np.diag(cov))
    got = ppd.reshape(P, M, -1)
    # per-draw means: standard error sd / sqrt(M); 5 sigma over P * n values
    assert np.all(np.abs(got.mean(axis=1) - mu) < 5.0 * sd / np.sqrt(M))
    # pooled residual covariance against the reconditioned covariance: relative error ~ sqrt(2 / (P M)) = 1 %
    res = (got - mu[:, None, :]).reshape(-1, x.size)
    emp = res.T @ res / res.shape[0]
    np.testing.assert_allclose(np.diag(emp), np.diag(cov), rtol=0.06)
    np.testing.assert_allclose(emp, cov, rtol=0, atol=0.06 * np.outer(sd, sd).max())
    # against the reference loop's samples: mean and variance per dt
    ref = ref.reshape(P, M, -1)
    se = np.sqrt(ppd.var(axis=0) / (P * M) + ref.reshape(-1, x.size).var(axis=0) / (P * M))
    res_ref = (ref - mu[:, None, :]).reshape(-1, x.size)
    assert np.all(np.abs(res.mean(axis=0) - res_ref.mean(axis=0)) < 5.0 * np.sqrt(2.0) * sd / np.sqrt(P * M))
    np.testing.assert_allclose(res.var(axis=0), res_ref.var(axis=0), rtol=0.08)
    assert np.all(se > 0)


def test_identical_single_origin_trajectories_stack(kb):
    """dtype='identical' with sampling='single-origin' (kinisi/analyzer.py:274-290 over parser.py:202-206): entry i of
    the stacked sequence is frame i of every trajectory, not the lag-1 multi-origin displacement."""
    symbols = ['Li'] * 5
    trajs, dcs = [], []
    for seed in (3, 4):
        frac, latt = ko.synthetic_walk(5, 60, box=8.0, sigma=0.2, seed=seed)
        trajs.append(trajectory_from_arrays(symbols, frac[:, 1:], latt[1:]))
        dcs.append(ko.get_disp(frac, latt))
    pp = dict(specie='Li', time_step=0.001, step_skip=1, n_steps=20, sampling='single-origin', progress=False)
    an = kb.analyze.DiffusionAnalyzer.from_ase(trajs, parser_params=pp, dtype='identical')
    one = kb.parser.ASEParser(trajs[0], **pp)
    frames = one.disp_3d.frames
    assert an._disp_3d.shape_of(0) == (10, 1, 3) and type(an._disp_3d).__name__ == 'SingleOriginDisplacements'
    joint = np.concatenate(dcs, axis=0)
    want = np.array([np.mean(np.sum(joint[:, f]**2, axis=-1)) for f in frames[:an.msd.size]])
    np.testing.assert_allclose(an.msd, want, rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(an._disp_3d[3], joint[:, frames[3]:frames[3] + 1], rtol=1e-12, atol=1e-13)


def test_streamed_displacements(kb, monkeypatch):
    """A displacement array above parser.STREAM_BYTES is never built (BASELINE configs[4] at one GPU: 120 GB of float64):
    stage 1 is re-run block by block in front of the moment kernels and the atom sum.  Same statistics as the resident
    path and the oracle (MSD, MSTD, MSCD with drift), several blocks with a partial last one; indexing the sequence
    still gives the reference's array (kinisi/parser.py:208-213)."""
    n_mobile, n_fw, n_frames = 45, 8, 2500
    frac, latt = ko.synthetic_walk(n_mobile, n_frames, box=14.0, sigma=0.1, seed=91, n_framework=n_fw, dtype=np.float32)
    idx, fw = list(range(n_mobile)), list(range(n_mobile, n_mobile + n_fw))
    parsed = ko.parse(ko.get_disp(frac.astype(np.float64), latt), idx, fw, 0.002, 1, n_steps=50)
    dev = torch.device('cuda')
    coords = torch.from_numpy(frac).to(dev)
    q = np.where(np.arange(n_mobile) % 3 == 0, -2.0, 1.0)
    pitch = kb.engine.pitch_for(n_frames)
    monkeypatch.setattr(kb.parser, 'STREAM_BYTES', 10 * 3 * pitch * 8)       # anything above 10 atoms streams
    monkeypatch.setattr(kb.parser, 'STREAM_BLOCK_BYTES', 7 * 3 * pitch * 8)  # 7 atoms per block: 6 blocks + 3 atoms
    p = kb.parser.Parser(kb.parser.Parser.get_disp(coords, latt[:1]), idx, fw, 0.002, 1, n_steps=50, progress=False)
    assert p.disp_3d.streamed and p.disp_3d.n_atoms == n_mobile
    for kind, cls, args in (('msd', kb.diffusion.MSDBootstrap, ()), ('mstd', kb.diffusion.MSTDBootstrap, ()),
                            ('mscd', kb.diffusion.MSCDBootstrap, (q, ))):
        b = cls(p.delta_t, p.disp_3d, *args, p._n_o, progress=False)
        ref = ko.bootstrap_streaming(parsed, kind, q if kind == 'mscd' else None)
        for key in ('dt', 'n', 'v', 's'):
            np.testing.assert_allclose(getattr(b, key), ref[key], rtol=RTOL_MOMENT, err_msg=f'{kind} {key}')
        np.testing.assert_allclose(b.ngp, ref['ngp'], rtol=1e-8, atol=1e-10)
        assert p.disp_3d.streamed  # the reductions did not build the array
    b.conductivity(float(b.dt[4]), 500.0, p.volume if p.volume else 1000.0, progress=False, n_samples=200, n_burn=100)
    assert np.isfinite(b.sigma.n)
    np.testing.assert_allclose(p.disp_3d[3], ko.displacements_for_lag(parsed['dc_sel'], int(parsed['lags'][3])), rtol=1e-10,
                               atol=1e-11)
    assert not p.disp_3d.streamed  # indexing built (and kept) it
    with pytest.raises(MemoryError):
        kb.parser.Parser(kb.parser.Parser.get_disp(coords, latt[:1]), idx, fw, 0.002, 1, n_steps=50, progress=False,
                         memory_limit=1e-6)
