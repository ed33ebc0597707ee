"""
Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(bjmorgan/kinisi 1.1.1, imported from /root/reference) in the build container.

    python tests/golden/make_golden.py

The reference imports three packages that are not installed here (uravu, emcee, statsmodels);
oracle/shims/ provides stand-ins (see oracle/shims/README.md).  Everything dumped below except the
sampler output is produced by the reference's own numpy/scipy arithmetic.  The reference cannot
travel to the GPU box, so the vectors are committed; this script is committed with them.
"""
import gzip
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shims'))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import emcee  # the shim; records the last sampler so the likelihood closure can be evaluated
from kinisi import diffusion, parser  # the unmodified reference
from kinisi.analyze import ConductivityAnalyzer, DiffusionAnalyzer, JumpDiffusionAnalyzer

from duck_atoms import trajectory_from_arrays
from synth_inputs import (CHARGES12, ase_case_inputs, list_bootstrap_inputs, parser_random_disp,
                          regress_case_inputs, ASE_CASES, REGRESS_CASES)

warnings.simplefilter('ignore')


def boot_arrays(b):
    return {'dt': b.dt, 'n': b.n, 'v': b.v, 's': b.s, 'ngp': b.ngp, 'n_o': np.asarray(b._n_o)}


def case_ase(name, charges):
    symbols, frac, latt, pp = ase_case_inputs(name)
    traj = trajectory_from_arrays(symbols, frac, latt)
    p = parser.ASEParser(traj, **pp)
    out = {
        'dc': p.dc,
        'delta_t': p.delta_t,
        'n_o': p._n_o,
        'time_intervals': p.time_intervals,
        'volume': p.volume,
        'disp_shapes': np.array([d.shape for d in p.disp_3d]),
        'disp_first': p.disp_3d[0],
        'disp_mid': p.disp_3d[len(p.disp_3d) // 2],
        'charges': charges,
    }
    for dim in ('xyz', 'xy', 'z', 'xz'):
        b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, dimension=dim, progress=False)
        for k, v in boot_arrays(b).items():
            out[f'msd_{dim}_{k}'] = v
        b = diffusion.MSTDBootstrap(p.delta_t, p.disp_3d, p._n_o, dimension=dim, progress=False)
        for k, v in boot_arrays(b).items():
            out[f'mstd_{dim}_{k}'] = v
        b = diffusion.MSCDBootstrap(p.delta_t, p.disp_3d, charges, p._n_o, dimension=dim, progress=False)
        for k, v in boot_arrays(b).items():
            out[f'mscd_{dim}_{k}'] = v
    b = diffusion.MSCDBootstrap(p.delta_t, p.disp_3d, 2, p._n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'mscd_scalar2_{k}'] = v
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, sub_sample_dt=3, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'msd_sub3_{k}'] = v
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'lags', p.time_intervals.size, 'kept', len(p.disp_3d))


def case_parser_random():
    """Mirrors kinisi/tests/test_parser.py:31-67 (numpy-only TestParser cases)."""
    disp = parser_random_disp()
    out = {}
    p = parser.Parser(disp, np.arange(100), [], 1, 1, progress=False)
    out['lin_delta_t'] = p.delta_t
    out['lin_n_o'] = p._n_o
    out['lin_shapes'] = np.array([d.shape for d in p.disp_3d])
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'lin_msd_{k}'] = v
    p = parser.Parser(disp, np.arange(0, 100, 2), np.arange(1, 100, 2), 0.5, 4, min_dt=10., max_dt=150., n_steps=30,
                      spacing='logarithmic', progress=False)
    out['log_delta_t'] = p.delta_t
    out['log_n_o'] = p._n_o
    out['log_intervals'] = p.time_intervals
    out['log_dc'] = p.dc
    out['log_shapes'] = np.array([d.shape for d in p.disp_3d])
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'log_msd_{k}'] = v
    # N = 1: the tail lags are dropped by the parser (parser.py:214-215)
    p = parser.Parser(disp[:1], np.arange(1), [], 1, 1, n_steps=20, progress=False)
    out['one_delta_t'] = p.delta_t
    out['one_n_o'] = p._n_o
    out['one_shapes'] = np.array([d.shape for d in p.disp_3d])
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'one_msd_{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'parser_random.npz'), **out)
    print('parser_random ok')


def case_list_bootstrap():
    """Mirrors the inputs of kinisi/tests/test_diffusion.py (RandomState(43).randn(100, i, 3) lists)."""
    out = {}
    dt, disp_3d, n_o, disp_skip = list_bootstrap_inputs()
    b = diffusion.MSDBootstrap(dt, disp_3d, n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'msd_{k}'] = v
    b = diffusion.MSTDBootstrap(dt, disp_3d, n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'mstd_{k}'] = v
    b = diffusion.MSCDBootstrap(dt, disp_3d, 1, n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out[f'mscd_{k}'] = v
    # the skip rule (test_diffusion.py:228-243): arrays with a single observation are skipped
    b = diffusion.MSDBootstrap(np.linspace(100, 400, 4), disp_skip, np.ones(4), progress=False)
    for k, v in boot_arrays(b).items():
        out[f'skip_msd_{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'list_bootstrap.npz'), **out)
    print('list_bootstrap ok')


def case_regression(name):
    """Stage 3 on a random walk: covariance, reconditioning, slogdet/pinvh, likelihood values."""
    symbols, frac, latt, pp, start_idx, kind = regress_case_inputs(name)
    traj = trajectory_from_arrays(symbols, frac, latt)
    cls = {'msd': DiffusionAnalyzer, 'mstd': JumpDiffusionAnalyzer, 'mscd': ConductivityAnalyzer}[kind]
    an = cls.from_ase(traj, parser_params=pp, uncertainty_params={'progress': False})
    start_dt = an.dt[start_idx]
    kw = {'progress': False, 'n_samples': 200, 'n_burn': 100, 'random_state': np.random.RandomState(3)}
    if kind == 'msd':
        an.diffusion(start_dt, kw)
        coeff = an.D.samples
    elif kind == 'mstd':
        an.jump_diffusion(start_dt, kw)
        coeff = an.D_J.samples
    else:
        an.conductivity(start_dt, 500.0, kw)
        coeff = an.sigma.samples
    b = an._diff
    loglike = emcee.LAST_SAMPLER.log_prob_fn
    g0, c0 = np.median(b.gradient.samples), np.median(b.intercept.samples)
    rng = np.random.RandomState(11)
    thetas = np.array([[g0, c0]] + [[g0 * (1 + 0.05 * rng.randn()), c0 + 0.05 * abs(c0 + 1e-3) * rng.randn()]
                                    for _ in range(10)] + [[-1.0, 0.0]])
    ll = np.array([loglike(t) for t in thetas])
    out = {
        'start_dt': start_dt,
        'npd_covariance_matrix': b._npd_covariance_matrix,
        'covariance_matrix': b.covariance_matrix,
        'thetas': thetas,
        'loglike': ll,
        'flatchain': b.flatchain,
        'coeff_samples': coeff,
        'gradient_samples': b.gradient.samples,
        'intercept_samples': b.intercept.samples,
        'volume': an.volume,
    }
    out.update(boot_arrays(b))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'n_dt', b.dt.size, 'cov', b.covariance_matrix.shape, 'grad', g0)


def read_xdatcar(path):
    opener = gzip.open if path.endswith('.gz') else open
    with opener(path, 'rt') as f:
        lines = f.read().split('\n')
    scale = float(lines[1])
    cell = np.array([[float(x) for x in lines[i].split()] for i in (2, 3, 4)]) * scale
    species = lines[5].split()
    counts = [int(x) for x in lines[6].split()]
    natoms = sum(counts)
    symbols = [s for s, c in zip(species, counts) for _ in range(c)]
    frames = []
    i = 7
    while i < len(lines):
        if lines[i].strip().startswith('Direct'):
            frames.append([[float(x) for x in lines[i + 1 + a].split()[:3]] for a in range(natoms)])
            i += natoms + 1
        else:
            i += 1
    frac = np.transpose(np.array(frames), (1, 0, 2))
    return symbols, frac, np.tile(cell, (frac.shape[1], 1, 1))


def case_reference_kats():
    """Known answers held by the reference's own tests and documentation."""
    out = {}
    # kinisi/tests/test_diffusion.py:105-107
    out['ngp_123'] = diffusion.Bootstrap.ngp_calculation(np.array([1, 2, 3]))
    # kinisi/tests/test_parser.py:151-168 (example_drift.XDATCAR: 1 H + 8 He, 4 frames, 10 A cube)
    symbols, frac, latt = read_xdatcar(os.path.join(REF, 'kinisi/tests/inputs/example_drift.XDATCAR'))
    traj = trajectory_from_arrays(symbols, frac, latt)
    p = parser.ASEParser(traj, specie='H', time_step=1, step_skip=1, progress=False)
    out['drift_symbols'], out['drift_frac'], out['drift_latt'] = np.array(symbols), frac, latt
    out['drift_dc0'] = p.dc[0]
    assert np.allclose(p.dc[0], [[0, 0, 0], [0.2, 0.2, 0.2], [0.4, 0.4, 0.4], [1, 1, 1]])
    p = parser.ASEParser(traj, specie='H', time_step=1, step_skip=1, progress=False, framework_indices=[])
    out['nodrift_dc0'] = p.dc[0]
    # docs/source/_static/displacements.npz: one materialised displacement array [192, 91, 3]
    disp = np.load(os.path.join(REF, 'docs/source/_static/displacements.npz'))['disp']
    d2 = np.sum(disp**2, axis=-1)
    out['docs_disp'] = disp
    out['docs_msd'] = d2.mean()
    out['docs_var_over_384'] = np.var(d2, ddof=1) / 384
    out['docs_ngp'] = diffusion.Bootstrap.ngp_calculation(d2)
    # kinisi/tests/test_analyze.py:92-98 on the real Li6PS5Cl-like XDATCAR (416 atoms x 140 frames)
    symbols, frac, latt = read_xdatcar(os.path.join(REF, 'kinisi/tests/inputs/example_XDATCAR.gz'))
    traj = trajectory_from_arrays(symbols, frac, latt)
    an = DiffusionAnalyzer.from_ase(traj, parser_params=dict(specie='Li', time_step=2.0, step_skip=50, progress=False),
                                    uncertainty_params={'progress': False})
    out['lips_symbols'] = np.array(symbols)
    out['lips_frac'] = frac.astype(np.float32)  # the file holds 8 decimals; stored as f32 to keep the fixture small
    out['lips_latt'] = latt[0]
    # rerun on the f32-rounded coordinates so the stored input reproduces the stored output exactly
    traj = trajectory_from_arrays(symbols, out['lips_frac'].astype(np.float64), latt)
    an = DiffusionAnalyzer.from_ase(traj, parser_params=dict(specie='Li', time_step=2.0, step_skip=50, progress=False),
                                    uncertainty_params={'progress': False})
    assert an._disp_3d[0].shape == (192, 140, 3) and an._disp_3d[-1].shape == (192, 1, 3)
    assert an._delta_t.max() == 14000.0  # ASEParser keeps time_step units (no fs->ps division)
    out['lips_delta_t'] = an._delta_t
    out['lips_n_o'] = an._n_o
    for k, v in boot_arrays(an._diff).items():
        out[f'lips_msd_{k}'] = v
    np.savez_compressed(os.path.join(HERE, 'reference_kats.npz'), **out)
    print('reference_kats ok', out['docs_msd'], out['docs_var_over_384'], out['docs_ngp'])


def case_resample():
    """`bootstrap=True` (kinisi/diffusion.py:258-289, 578-583, 783-801) run by the reference + scikit-learn with seeded
    RandomStates, on the inputs of kinisi/tests/test_diffusion.py:86-103, 192-208, 451-466, 839-842."""
    out = {}
    out['bootstrap_arange'] = np.array(diffusion._bootstrap(np.arange(1, 100, 1), 200, 100, np.random.RandomState(0)))
    d = diffusion.Bootstrap.sample_until_normal(np.arange(1, 10, 1), 5, 100, 100, random_state=np.random.RandomState(0))
    out['until_normal_small'] = d.samples
    rng = np.random.RandomState(5)
    skewed = rng.exponential(size=(40, 25))**2
    out['skewed_population'] = skewed
    d = diffusion.Bootstrap.sample_until_normal(skewed, 7, 200, 2000, random_state=np.random.RandomState(1))
    out['until_normal_skewed'] = d.samples
    dt, disp_3d, n_o, _ = list_bootstrap_inputs()
    for kind, cls, args in (('msd', diffusion.MSDBootstrap, ()), ('mstd', diffusion.MSTDBootstrap, ())):
        b = cls(dt, disp_3d, *args, n_o, bootstrap=True, random_state=np.random.RandomState(0), progress=False)
        out[f'{kind}_n_bootstrap'] = b._n_bootstrap
        out[f'{kind}_v_bootstrap'] = b._v_bootstrap
        out[f'{kind}_s_bootstrap'] = b._s_bootstrap
        out[f'{kind}_sizes'] = np.array([x.size for x in b._distributions])
        out[f'{kind}_first_distribution'] = b._distributions[0].samples
        for k, v in boot_arrays(b).items():
            out[f'{kind}_{k}'] = v
    b = diffusion.MSCDBootstrap(dt, disp_3d, 1, n_o, bootstrap=True, random_state=np.random.RandomState(0), progress=False)
    out['mscd_n_bootstrap'] = b._n_bootstrap
    out['mscd_v_bootstrap'] = b._v_bootstrap
    out['mscd_sizes'] = np.array([x.size for x in b._distributions])
    np.savez_compressed(os.path.join(HERE, 'resample_reference.npz'), **out)
    print('resample_reference ok', out['msd_sizes'], out['mstd_sizes'], out['until_normal_skewed'].size)


def case_universe():
    """The MDAnalysis front-end (kinisi/parser.py:511-672) run by the reference on a duck-typed universe
    (tests/duck_atoms.py): float64 positions (tight parity) and float32 ones (what MDAnalysis hands out: the reference
    then inverts the cell and forms the fractional coordinates in single precision)."""
    from duck_atoms import universe_from_arrays
    out = {}
    for name in ('ase_nvt', 'ase_npt_triclinic'):
        symbols, frac, latt, pp = ase_case_inputs(name)
        for tag, dtype in (('f64', np.float64), ('f32', np.float32)):
            p = parser.MDAnalysisParser(universe_from_arrays(symbols, frac, latt, dtype), **pp)
            key = f'{name}_{tag}_'
            out[key + 'dc'] = p.dc
            out[key + 'delta_t'] = p.delta_t
            out[key + 'n_o'] = p._n_o
            out[key + 'volume'] = p.volume
            out[key + 'coords_check'] = p.coords_check
            b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
            for k, v in boot_arrays(b).items():
                out[key + 'msd_' + k] = v
    symbols, frac, latt, pp = ase_case_inputs('ase_npt_triclinic')
    pp = dict(pp, sub_sample_atoms=2, sub_sample_traj=3)
    p = parser.MDAnalysisParser(universe_from_arrays(symbols, frac, latt, np.float64), **pp)
    out['sub_dc'], out['sub_delta_t'], out['sub_n_o'] = p.dc, p.delta_t, p._n_o
    b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
    for k, v in boot_arrays(b).items():
        out['sub_msd_' + k] = v
    pp = dict(ase_case_inputs('ase_nvt')[3], specie=None, specie_indices=[0, 2, 3, 5])
    symbols, frac, latt, _ = ase_case_inputs('ase_nvt')
    p = parser.MDAnalysisParser(universe_from_arrays(symbols, frac, latt, np.float64), **pp)
    out['idx_dc'], out['idx_delta_t'], out['idx_n_o'] = p.dc, p.delta_t, p._n_o
    out['idx_indices'], out['idx_drift_indices'] = np.array(p.indices), np.array(p.drift_indices)
    np.savez_compressed(os.path.join(HERE, 'universe_reference.npz'), **out)
    print('universe_reference ok', out['sub_dc'].shape, out['idx_dc'].shape)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'universe':  # only the fixture added for the MDAnalysis front-end
        case_universe()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'resample':  # only the fixture added for SURVEY 8f-4
        case_resample()
        sys.exit(0)
    for name in ASE_CASES:
        case_ase(name, CHARGES12)
    case_parser_random()
    case_list_bootstrap()
    for name in REGRESS_CASES:
        case_regression(name)
    case_reference_kats()
    case_resample()
    case_universe()
