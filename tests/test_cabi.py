"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports every symbol that
include/kinisi_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    from kinisi_b200 import _build
    return _build.build(force=False)


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'kinisi_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(kb2_\w+)\s*\(', text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 22
    for name in names:
        assert hasattr(lib, name), f'{name} is declared in include/kinisi_b200.h but not exported'


def test_bindings_cover_the_header(lib_path):
    from kinisi_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.kb2_abi_version() == _lib.ABI_VERSION
    assert lib.kb2_self_moments_workspace_bytes(100) > 0
    assert lib.kb2_sampler_workspace_bytes(1500, 32) > 1500 * 32 * 28


def test_argument_errors_are_reported_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call, with a message through kb2_last_error()."""
    from kinisi_b200 import _lib
    lib = _lib.load()
    rc = lib.kb2_cov_build(None, None, 0, None, None)
    assert rc != 0 and b'kb2_cov_build' in lib.kb2_last_error()
    rc = lib.kb2_self_moments(None, 1, 1, 2, None, 1, 7, 0, None, None, None)
    assert rc != 0 and b'null pointer' in lib.kb2_last_error()
    import numpy as np
    lags = np.array([5, 3, 9], dtype=np.int32)  # not ascending
    buf = np.zeros(lib.kb2_self_moments_runs_workspace_bytes(3, 100), dtype=np.uint8)
    rc = lib.kb2_self_moments_runs_plan_pack(lags.ctypes.data, 3, 100, buf.ctypes.data)
    assert rc != 0 and b'ascend' in lib.kb2_last_error()
    lags = np.array([1, 50, 101], dtype=np.int32)  # beyond the trajectory
    rc = lib.kb2_self_moments_runs_plan_pack(lags.ctypes.data, 3, 100, buf.ctypes.data)
    assert rc != 0 and b'out of range' in lib.kb2_last_error()
    rc = lib.kb2_spectral_prepare(1, 1, 1, 129, 1e16, 1, 1, 1, 1, 1, None, None)
    assert rc != 0 and b'[1, 128]' in lib.kb2_last_error()
    rc = lib.kb2_ensemble_sample(1, 2, 33, 10, 1, 0, 1, 1, 1, None)
    assert rc != 0 and b'even' in lib.kb2_last_error()
    rc = lib.kb2_unwrap_cumsum(1, 0, 0, 4, 10, 1, 1, 0, 1, 2, None, 0.0, 1, 9, None, None)  # odd pitch, no workspace
    assert rc != 0
    rc = lib.kb2_frames_to_atoms(1, 0, 4, 10, 1, 10, 0, 1, None)  # no room for the duplicated first frame
    assert rc != 0 and b'do not fit' in lib.kb2_last_error()
    rc = lib.kb2_frames_to_atoms_fractional(1, 0, 4, 10, 1, 3, 1, 11, 1, 1, None)
    assert rc != 0 and b'cell_stride' in lib.kb2_last_error()
    rc = lib.kb2_molecule_centres(1, 7, 4, 10, 1, 1, 3, None, 1, None)
    assert rc != 0 and b'dtype' in lib.kb2_last_error()
    rc = lib.kb2_lag_squares(1, 2, 10, 10, 11, 7, 1, None)
    assert rc != 0 and b'out of range' in lib.kb2_last_error()
    rc = lib.kb2_disp_squares(1, 2, 10, 0, None, 0, 1, None)
    assert rc != 0 and b'dim_mask' in lib.kb2_last_error()
    rc = lib.kb2_resample_means(1, 0, 5, 10, 0, 1, 0, 1, None)
    assert rc != 0 and b'empty population' in lib.kb2_last_error()
    rc = lib.kb2_resample_means(1, 10, 0, 10, 0, 1, 0, 1, None)
    assert rc != 0 and b'n_draws' in lib.kb2_last_error()
    rc = lib.kb2_reblock(16, 1, 16, 16, None)
    assert rc != 0 and b'two values' in lib.kb2_last_error()
    rc = lib.kb2_reblock(24, 8, 16, 16, None)
    assert rc != 0 and b'aligned' in lib.kb2_last_error()
    assert lib.kb2_reblock_levels(1) == 0 and lib.kb2_reblock_levels(2) == 1 and lib.kb2_reblock_levels(7) == 2
    assert lib.kb2_reblock_workspace_bytes(7) >= 8 * (4 + 2)


def test_peer_entry_points_reject_bad_arguments_without_a_device(lib_path):
    """The joins over ranks (csrc/k7_peer.cu, peer.cuh): the sizing helpers are host arithmetic, and the argument checks come
    before any CUDA call."""
    import ctypes
    from kinisi_b200 import _lib
    lib = _lib.load()
    assert lib.kb2_peer_slots() == 8
    assert lib.kb2_peer_rows_slices(1024) == 1 and lib.kb2_peer_rows_slices(1026) == 2 and lib.kb2_peer_rows_slices(131072) == 128
    assert lib.kb2_peer_rows_slices(131074) == 0 and lib.kb2_peer_rows_slices(1023) == 0  # too long for one join; odd
    assert lib.kb2_peer_window_bytes(0, 1024) == 0 and lib.kb2_peer_window_bytes(17, 1024) == 0
    two, eight = lib.kb2_peer_window_bytes(2, 1 << 15), lib.kb2_peer_window_bytes(8, 1 << 15)
    assert two >= 8 * 2 * (1 << 15) * 8 and eight >= 4 * two - (1 << 20)  # slots x ranks x capacity doubles, plus the flags
    windows = (ctypes.c_void_p * 2)(16, 32)
    rc = lib.kb2_peer_allreduce_f64(16, 100, windows, 2, 0, 64, 0, None)
    assert rc != 0 and b'capacity' in lib.kb2_last_error()
    rc = lib.kb2_peer_allreduce_f64(24, 10, windows, 2, 0, 64, 0, None)
    assert rc != 0 and b'aligned' in lib.kb2_last_error()
    rc = lib.kb2_peer_allreduce_rows_f64(16, 3, 1023, windows, 2, 0, 1 << 20, 0, None)
    assert rc != 0 and b'even' in lib.kb2_last_error()
    rc = lib.kb2_moment_stats_joined(16, 16, 16, 16, 16, 600, 16, 16, 16, 16, windows, 2, 0, 1 << 15, 0, None)
    assert rc != 0 and b'512' in lib.kb2_last_error()
    rc = lib.kb2_frame_sums_joined(16, 0, 0, 4, 10, 16, 16, 0, 16, None, 2, 16, 16, windows, 2, 0, 32, 0, 16, None)
    assert rc != 0 and b'capacity' in lib.kb2_last_error()
    assert lib.kb2_peer_status() == 0


def test_run_plan_is_host_only(lib_path):
    """kb2_self_moments_runs_plan_pack needs no device: the packed plan of BASELINE configs[1]'s grid has one block with
    one 99-lag progression and one single lag."""
    import numpy as np
    from kinisi_b200 import _lib
    lib = _lib.load()
    lags = np.linspace(1, 20000, 100, dtype=int).astype(np.int32)
    n = lib.kb2_self_moments_runs_workspace_bytes(100, 20000)
    buf = np.zeros(n, dtype=np.uint8)
    assert lib.kb2_self_moments_runs_plan_pack(lags.ctypes.data, 100, 20000, buf.ctypes.data) == 0
    header = buf[:32].view(np.int32)
    assert header[0] == 100 and header[1] == 2 and header[3] == 0  # K, runs, no drift
    assert header[2] > 80  # (run, strip) pairs


def test_kernel_choice_is_host_only(lib_path):
    """kb2_self_moments_choose needs no device.  BASELINE configs[1] / [2] / [4] (kinisi's default 100-point grid,
    kinisi/parser.py:150-168, on >= 20 000 frames) keep the tiled kernel; configs[3]'s dense grid (2000 points on 20 000
    frames) is one long run of spacing 10 -> the run kernel; the default grid on configs[0]'s 5000 frames (spacing
    alternating 50 / 51) spreads over too many residues for the tiles -> the window kernel; a logarithmic grid
    (kinisi/parser.py:165) has no structure -> the direct kernel.  tile_plan_cost agrees with the terms of the grid."""
    import numpy as np
    from kinisi_b200 import _lib, engine
    lib = _lib.load()

    def grid(T, n, used=None):
        return np.unique(np.linspace(1, T, n, dtype=int))[:used].astype(np.int32)

    assert engine.choose_variant(grid(20000, 100), 20000) == engine.VARIANT_TILE
    assert engine.choose_variant(grid(100000, 100, 50), 100000) == engine.VARIANT_TILE
    assert engine.choose_variant(grid(50000, 100), 50000) == engine.VARIANT_TILE
    assert engine.choose_variant(grid(20000, 2000, 1000), 20000) == engine.VARIANT_RUNS
    assert engine.choose_variant(grid(5000, 100), 5000) == engine.VARIANT_WIN
    log = np.unique(np.geomspace(1, 20000, 100).astype(int)).astype(np.int32)
    v, feats = engine.choose_variant(log, 20000, with_features=True)
    assert v == engine.VARIANT_DIRECT and feats['short_window_share'] > 0.5 and feats['tile_bytes_per_term'] > 2.0
    assert engine.choose_variant(np.array([3], dtype=np.int32), 50) in (0, 2, 3, 4)
    # errors: unsorted lags, a lag beyond the trajectory
    bad = np.array([5, 3], dtype=np.int32)
    assert lib.kb2_self_moments_choose(bad.ctypes.data, 2, 100, None) == -1 and b'ascend' in lib.kb2_last_error()
    with pytest.raises(ValueError, match='out of range'):
        engine.choose_variant(np.array([1, 200], dtype=np.int32), 100)
    # the plan walks every term of the grid once (ragged strips count whole: a few per cent above)
    lags = grid(20000, 100)
    out = np.zeros(3)
    assert lib.kb2_self_moments_tile_plan_cost(lags.ctypes.data, len(lags), 20000, out.ctypes.data) == 0
    terms = float(np.sum(20000 - lags))
    assert terms <= out[2] <= 1.05 * terms
    assert out[0] < 2.0 * terms and out[1] > 0


def test_window_cut_follows_the_instruction_cache(lib_path):
    """The tile planner cuts a progression into windows so that the column walks running side by side fit the instruction
    cache (csrc/k2_moments_tile.cu, plan_tiles): groups of 11 lags between carries (first half of kinisi's default grid,
    kinisi/parser.py:150-168, on 100 000 frames: what MSTD / MSCD keep) become 6 + 5, not 8 + 3 next to a 6; one long
    progression (20 000 frames) and four of 25 (50 000 frames) stay windows of 8 and a remainder.  Host only."""
    import numpy as np
    from collections import Counter
    from kinisi_b200 import engine

    def cut(T, used):
        lags = np.linspace(1, T, 100, dtype=int)[:used].astype(np.int32)
        wins, n_tiles = engine.self_moments_tiles(lags, T)
        assert sum(w[2] for w in wins) == used and n_tiles >= 1
        return Counter(w[2] for w in wins)

    assert cut(100000, 50) == Counter({6: 5, 5: 4})
    assert cut(20000, 100) == Counter({8: 12, 3: 1, 1: 1})
    assert cut(50000, 100) == Counter({8: 12, 1: 4})


def test_sass_has_tma_and_fp64(lib_path):
    """The TMA variant really uses the bulk-copy engine (UBLKCP) and the moment kernels are FP64."""
    cuobjdump = '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    assert 'UBLKCP' in sass and 'SYNCS' in sass
    assert 'UBLKPF' in sass  # cp.async.bulk.prefetch.L2 (the unwrap kernel's look-ahead)
    assert 'DFMA' in sass


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'kinisi_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'oracle' not in src.replace('kinisi_oracle', 'oracle') or 'import' not in [
                line for line in src.splitlines() if 'oracle' in line and 'import' in line
            ], fn
            for line in src.splitlines():
                assert not (('import' in line) and ('oracle' in line)), f'{fn}: {line}'


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('only meaningful on a box without a GPU')
    import numpy as np
    from kinisi_b200 import diffusion, parser
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        parser.Parser(np.zeros((4, 10, 3)), [0, 1], [], 1, 1, progress=False)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        diffusion.MSDBootstrap([1.0], [np.zeros((2, 2, 3))], [1.0])
