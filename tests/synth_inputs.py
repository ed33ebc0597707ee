"""Seeded synthetic inputs shared by tests/golden/make_golden.py and the tests (numpy legacy
RandomState streams are frozen, so regenerating from the seed reproduces the fixture inputs)."""
import numpy as np


def random_walk_frac(n_mobile, n_fw, n_frames, box, sigma, seed, triclinic=False, npt=False):
    """Wrapped fractional coordinates [N_all, n_frames, 3] and lattices [n_frames, 3, 3] of a 3-D
    Gaussian random walk (mobile atoms first) plus jittering, slowly drifting framework atoms."""
    rng = np.random.RandomState(seed)
    n_all = n_mobile + n_fw
    start = rng.uniform(0, 1, size=(n_all, 1, 3))
    latt = np.tile(np.eye(3) * box, (n_frames, 1, 1))
    if triclinic:
        latt[:, 0, 1] = 0.07 * box
        latt[:, 2, 0] = -0.04 * box
        latt[:, 1, 2] = 0.02 * box
    if npt:
        breathe = 1.0 + 0.01 * np.sin(np.arange(n_frames) / 7.0) + 0.002 * rng.randn(n_frames)
        latt = latt * breathe[:, None, None]
    steps = rng.normal(0, sigma / box, size=(n_mobile, n_frames, 3))
    steps[:, 0] = 0
    frac = np.empty((n_all, n_frames, 3))
    frac[:n_mobile] = start[:n_mobile] + np.cumsum(steps, axis=1)
    if n_fw:
        frac[n_mobile:] = (start[n_mobile:] + rng.normal(0, 0.05 * sigma / box, size=(n_fw, n_frames, 3)) +
                           (0.02 * sigma / box) * np.arange(n_frames)[None, :, None])
    return frac % 1.0, latt


# (seed, n_mobile, n_fw, n_frames, box, triclinic, npt, parser kwargs)
ASE_CASES = {
    'ase_nvt': (101, 12, 6, 90, 8.0, False, False, {}),
    'ase_npt_triclinic': (102, 12, 5, 70, 7.5, True, True, {'n_steps': 40}),
    'ase_subsample': (103, 12, 0, 120, 9.0, False, False, {
        'sub_sample_traj': 2,
        'min_dt': 0.05,
        'max_dt': 0.4,
        'n_steps': 25,
        'framework_indices': [0, 3]
    }),
}
ASE_BASE_PARAMS = dict(specie='Li', time_step=0.002, step_skip=5, progress=False)
CHARGES12 = np.array([1., 1., 1., -1., -1., 2., 1., -2., 1., 1., -1., 1.])

# name -> (seed, n_mobile, n_frames, n_steps, start_idx, kind); 4 framework atoms, box 12, sigma 0.2
REGRESS_CASES = {
    'regress_msd': (201, 48, 1500, 60, 6, 'msd'),
    'regress_mstd': (202, 32, 2000, 80, 4, 'mstd'),
    'regress_mscd': (203, 32, 2000, 80, 4, 'mscd'),
}
REGRESS_BASE_PARAMS = dict(specie='Li', time_step=0.001, step_skip=10, progress=False)


def ase_case_inputs(name):
    seed, n_mobile, n_fw, n_frames, box, tri, npt, kw = ASE_CASES[name]
    frac, latt = random_walk_frac(n_mobile, n_fw, n_frames, box, 0.25, seed, tri, npt)
    symbols = ['Li'] * n_mobile + ['O'] * n_fw
    pp = dict(ASE_BASE_PARAMS)
    pp.update(kw)
    return symbols, frac, latt, pp


def regress_case_inputs(name):
    seed, n_mobile, n_frames, n_steps, start_idx, kind = REGRESS_CASES[name]
    frac, latt = random_walk_frac(n_mobile, 4, n_frames, 12.0, 0.2, seed)
    symbols = ['Li'] * n_mobile + ['O'] * 4
    pp = dict(REGRESS_BASE_PARAMS)
    pp['n_steps'] = n_steps
    return symbols, frac, latt, pp, start_idx, kind


def parser_random_disp():
    return np.random.RandomState(7).random_sample((100, 100, 3))


def list_bootstrap_inputs():
    """Same construction as kinisi/tests/test_diffusion.py: RandomState(43).randn(100, i, 3)."""
    rng = np.random.RandomState(43)
    disp_3d = [rng.randn(100, i, 3) for i in range(20, 10, -1)]
    n_o = np.ones(len(disp_3d)) * 100
    dt = np.linspace(100, 1000, 10)
    disp_skip = [rng.randn(1, i, 3) for i in range(4, 0, -1)]
    return dt, disp_3d, n_o, disp_skip
