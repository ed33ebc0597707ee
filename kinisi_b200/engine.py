"""
Thin wrappers that call the C ABI (include/kinisi_b200.h) on PyTorch CUDA tensors.  PyTorch is only the
carrier of device memory and streams here; every function below enqueues hand-written sm_100a kernels
on torch's current stream and returns without synchronising.
"""
import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, MODE_DISPLACEMENT, MODE_FRACTIONAL, check

DIM_MASK = {'x': 1, 'y': 2, 'z': 4, 'xy': 3, 'xz': 5, 'yz': 6, 'xyz': 7}

# self-moment kernel variants (kb2_self_moments `variant`)
VARIANT_DIRECT = 0
VARIANT_TMA = 1
VARIANT_RUNS = 2  # arithmetic-run kernel (origin window in registers); plans on the host, needs the lags there
VARIANT_WIN = 3   # lag-window kernel (partner window + accumulators in registers); plans on the host too
VARIANT_TILE = 4  # tiled lag-window kernel: tiles staged in shared memory by TMA bulk copies; plans on the host too
VARIANT_AUTO = -1  # chosen per lag grid by kb2_self_moments_choose (tiles for kinisi's default grid on long trajectories)
DEFAULT_VARIANT = VARIANT_AUTO

# count of kernel launches issued through this module (bench.py reports it as gpu_launches)
LAUNCHES = {'n': 0}


def _count(n=1):
    LAUNCHES['n'] += n


_LIB = None


def require_cuda():
    global _LIB
    if _LIB is None:
        if not torch.cuda.is_available():
            raise RuntimeError('kinisi_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
        _LIB = _lib.load()
    return _LIB


def _stream():
    # the raw handle of torch's current stream (the python Stream object costs ~15 us to build, per launch)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


_STREAM_OBJECTS = {}


def current_stream(device):
    """torch.cuda.current_stream(device) without building a Stream object per call (~10 us each, a dozen per analysis):
    the objects are kept by raw handle (torch's streams come from a pool and live as long as the process)."""
    idx = device.index if device.index is not None else torch._C._cuda_getDevice()
    raw = torch._C._cuda_getCurrentRawStream(idx)
    s = _STREAM_OBJECTS.get((idx, raw))
    if s is None:
        s = _STREAM_OBJECTS[(idx, raw)] = torch.cuda.current_stream(device)
    return s


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def dim_mask(dimension):
    if isinstance(dimension, bytes):
        dimension = dimension.decode()
    return DIM_MASK[dimension.lower()]


def pitch_for(n_frames):
    """Row pitch (in float64 elements) of the component-major displacement array: a multiple of 16."""
    return (int(n_frames) + 15) // 16 * 16


def coords_dtype_code(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float64:
        return F64
    raise TypeError(f'coordinates must be float32 or float64, got {t.dtype}')


def invert_lattice(latt):
    lib = require_cuda()
    latt = latt.contiguous()
    inv = torch.empty_like(latt)
    check(lib.kb2_invert_lattice(latt.data_ptr(), inv.data_ptr(), latt.shape[0], _stream()), 'kb2_invert_lattice')
    _count()
    return inv


def frames_to_atoms(frames, out, frame_offset, duplicate_first=False, cell_inv=None):
    """Transposes frame-major frames [n_frames, n_atoms, 3] into the atom-major array out [n_atoms, F_out, 3] at
    frame_offset; see kb2_frames_to_atoms.  With cell_inv ([n_frames, 3, 3] or [1, 3, 3] float64) the frames are Cartesian
    and out (float64) receives the fractional coordinates; see kb2_frames_to_atoms_fractional."""
    lib = require_cuda()
    if cell_inv is not None:
        assert frames.is_contiguous() and out.is_contiguous() and out.dtype == torch.float64 and cell_inv.is_contiguous()
        stride = 9 if cell_inv.shape[0] > 1 else 0
        assert stride == 0 or cell_inv.shape[0] == frames.shape[0]
        check(
            lib.kb2_frames_to_atoms_fractional(frames.data_ptr(), coords_dtype_code(frames), frames.shape[1], frames.shape[0],
                                               cell_inv.data_ptr(), stride, out.data_ptr(), out.shape[1], int(frame_offset),
                                               int(bool(duplicate_first)), _stream()), 'kb2_frames_to_atoms_fractional')
        _count()
        return out
    assert frames.is_contiguous() and out.is_contiguous() and frames.dtype == out.dtype
    check(
        lib.kb2_frames_to_atoms(frames.data_ptr(), coords_dtype_code(frames), frames.shape[1], frames.shape[0],
                                out.data_ptr(), out.shape[1], int(frame_offset), int(bool(duplicate_first)), _stream()),
        'kb2_frames_to_atoms')
    _count()
    return out


def molecule_centres(coords, members, weights=None):
    """Pseudo-centres of mass [n_molecules, F, 3] (float64, fractional) of the atom groups members [n_mol, m] (int32
    device tensor) of coords [n_atoms, F, 3]; see kb2_molecule_centres."""
    lib = require_cuda()
    n_mol, m = members.shape
    out = torch.empty((n_mol, coords.shape[1], 3), dtype=torch.float64, device=coords.device)
    check(
        lib.kb2_molecule_centres(coords.data_ptr(), coords_dtype_code(coords), coords.shape[0], coords.shape[1],
                                 members.data_ptr(), n_mol, m, _ptr(weights), out.data_ptr(), _stream()),
        'kb2_molecule_centres')
    _count()
    return out


def frame_sums(coords, mode, latt, latt_inv, latt_stride, idx, weights, pitch, join=None):
    """[3, pitch] per-frame (weighted) sums over the atoms in idx; see kb2_frame_sums.  `join`: a peer-memory channel
    (dist.frame_sums_channel) -- the kernel then ends with the sum over ranks (kb2_frame_sums_joined)."""
    lib = require_cuda()
    out = torch.empty((3, pitch), dtype=torch.float64, device=coords.device)
    if join is not None:
        windows, world, rank, capacity, seq, counters = join.frame_sums_join()
        check(
            lib.kb2_frame_sums_joined(coords.data_ptr(), coords_dtype_code(coords), mode, coords.shape[0], coords.shape[1],
                                      _ptr(latt), _ptr(latt_inv), latt_stride, idx.data_ptr(), _ptr(weights), idx.numel(),
                                      out.data_ptr(), pitch, windows, world, rank, capacity, seq, counters, _stream()),
            'kb2_frame_sums_joined')
        _count(2)
        return out
    check(
        lib.kb2_frame_sums(coords.data_ptr(), coords_dtype_code(coords), mode, coords.shape[0], coords.shape[1],
                           _ptr(latt), _ptr(latt_inv), latt_stride, idx.data_ptr(), _ptr(weights), idx.numel(),
                           out.data_ptr(), pitch, _stream()), 'kb2_frame_sums')
    _count(2)
    return out


def scan_frames(buf, n_frames, cumulative, divide, sub=None, sub_scale=0.0):
    lib = require_cuda()
    check(
        lib.kb2_scan_frames(buf.data_ptr(), n_frames, buf.shape[1], int(bool(cumulative)), float(divide), _ptr(sub),
                            float(sub_scale), _stream()), 'kb2_scan_frames')
    _count()
    return buf


def unwrap_cumsum(coords, mode, latt, latt_inv, latt_stride, idx, step_sums, step_scale, pitch, out=None):
    """dc [n_idx, 3, pitch] float64 component-major, drift-corrected with the per-frame framework sums;
    see kb2_unwrap_cumsum.  `out`: a contiguous [n_idx, 3, pitch] block to fill instead of a new tensor."""
    lib = require_cuda()
    if out is None:
        out = torch.empty((idx.numel(), 3, pitch), dtype=torch.float64, device=coords.device)
    assert out.is_contiguous() and tuple(out.shape) == (idx.numel(), 3, pitch)
    ws = torch.empty(lib.kb2_unwrap_cumsum_workspace_bytes(idx.numel(), pitch), dtype=torch.uint8, device=coords.device)
    check(
        lib.kb2_unwrap_cumsum(coords.data_ptr(), coords_dtype_code(coords), mode, coords.shape[0], coords.shape[1],
                              _ptr(latt), _ptr(latt_inv), latt_stride, idx.data_ptr(), idx.numel(), _ptr(step_sums),
                              float(step_scale), out.data_ptr(), pitch, ws.data_ptr(), _stream()), 'kb2_unwrap_cumsum')
    _count()
    return out


def atom_sum(dc, weights=None):
    """[3, pitch] (weighted) sum over atoms of a component-major displacement array; see kb2_atom_sum."""
    lib = require_cuda()
    out = torch.empty((3, dc.shape[2]), dtype=torch.float64, device=dc.device)
    check(lib.kb2_atom_sum(dc.data_ptr(), dc.shape[0], dc.shape[2], _ptr(weights), out.data_ptr(), _stream()),
          'kb2_atom_sum')
    _count(2)
    return out


def self_moments(dc, n_frames, lags, mask=7, variant=None):
    """sums [K, 2] = (sum d^2, sum d^4) per lag over atoms x observations; see kb2_self_moments_tile (the default),
    kb2_self_moments_win, kb2_self_moments_runs, kb2_self_moments.

    dc: [n_atoms, 3, pitch] float64; lags: ascending, a host array (numpy / list) or an int32 device tensor.
    The tiled, window and run variants plan on the host: with device-resident lags that costs a device->host copy.
    variant=None: DEFAULT_VARIANT; VARIANT_AUTO (the shipped default) chooses by the lag grid (choose_variant)."""
    lib = require_cuda()
    K = len(lags) if not isinstance(lags, torch.Tensor) else lags.numel()
    if variant is None:
        variant = DEFAULT_VARIANT
    if variant == VARIANT_AUTO:
        lags_host = np.ascontiguousarray(lags.cpu().numpy() if isinstance(lags, torch.Tensor) else lags, dtype=np.int32)
        variant = choose_variant(lags_host, n_frames)
    sums = torch.empty((K, 2), dtype=torch.float64, device=dc.device)
    if (variant & 0xff) == VARIANT_TILE:
        lags_host = np.ascontiguousarray(lags.cpu().numpy() if isinstance(lags, torch.Tensor) else lags, dtype=np.int32)
        plan_host, plan_dev = _host_plan(lib, 'tile', lags_host, n_frames, dc.device)
        blocks = lib.kb2_self_moments_tile_blocks(K)
        counters = torch.empty(256 * blocks, dtype=torch.uint8, device=dc.device)
        check(
            lib.kb2_self_moments_tile_planned(dc.data_ptr(), dc.shape[0], n_frames, dc.shape[2], K, mask,
                                              plan_host.data_ptr(), plan_dev.data_ptr(), counters.data_ptr(),
                                              sums.data_ptr(), _stream()), 'kb2_self_moments_tile_planned')
        _count(blocks)
        return sums
    if (variant & 0xff) == VARIANT_WIN:
        lags_host = np.ascontiguousarray(lags.cpu().numpy() if isinstance(lags, torch.Tensor) else lags, dtype=np.int32)
        plan_host, plan_dev = _host_plan(lib, 'win', lags_host, n_frames, dc.device)
        blocks = lib.kb2_self_moments_win_blocks(K)
        counters = torch.empty(256 * blocks, dtype=torch.uint8, device=dc.device)
        check(
            lib.kb2_self_moments_win_planned(dc.data_ptr(), dc.shape[0], n_frames, dc.shape[2], K, mask,
                                             plan_host.data_ptr(), plan_dev.data_ptr(), counters.data_ptr(),
                                             sums.data_ptr(), _stream()), 'kb2_self_moments_win_planned')
        _count(blocks)
        return sums
    if (variant & 0xff) == VARIANT_RUNS:
        lags_host = np.ascontiguousarray(lags.cpu().numpy() if isinstance(lags, torch.Tensor) else lags, dtype=np.int32)
        plan_host, plan_dev = _host_plan(lib, 'runs', lags_host, n_frames, dc.device)
        blocks = lib.kb2_self_moments_runs_blocks(K)
        counters = torch.empty(256 * blocks, dtype=torch.uint8, device=dc.device)
        check(
            lib.kb2_self_moments_runs_planned(dc.data_ptr(), dc.shape[0], n_frames, dc.shape[2], K, mask,
                                              plan_host.data_ptr(), plan_dev.data_ptr(), counters.data_ptr(),
                                              sums.data_ptr(), _stream()), 'kb2_self_moments_runs_planned')
        _count(blocks)
        return sums
    if not isinstance(lags, torch.Tensor):
        lags = small_to_device(lags, dc.device, np.int32)
    ws = torch.empty(lib.kb2_self_moments_workspace_bytes(K), dtype=torch.uint8, device=dc.device)
    check(
        lib.kb2_self_moments(dc.data_ptr(), dc.shape[0], n_frames, dc.shape[2], lags.data_ptr(), K, mask, variant,
                             ws.data_ptr(), sums.data_ptr(), _stream()), 'kb2_self_moments')
    _count(3 * ((K + 767) // 768))
    return sums


_CHOSEN = {}  # (lags, n_frames) -> (variant, features)


def choose_variant(lags_host, n_frames, with_features=False):
    """The moment kernel for a lag grid (kb2_self_moments_choose): the tiled kernel while its plan copies <= 2 bytes per
    term (kinisi's default grid, kinisi/parser.py:150-168, on long trajectories), else the run kernel (long runs of one
    spacing), the window kernel (long windows) or the direct kernel (no structure: spacing='logarithmic',
    kinisi/parser.py:165).  Host only; cached per (grid, n_frames)."""
    lags_host = np.ascontiguousarray(lags_host, dtype=np.int32)
    key = (lags_host.tobytes(), int(n_frames))
    hit = _CHOSEN.get(key)
    if hit is None:
        feats = np.zeros(4)
        v = _lib.load().kb2_self_moments_choose(lags_host.ctypes.data, len(lags_host), int(n_frames), feats.ctypes.data)
        if v < 0:
            raise ValueError('kb2_self_moments_choose: ' + _lib.load().kb2_last_error().decode('utf-8', 'replace'))
        hit = (v, dict(zip(('tile_bytes_per_term', 'runs_model', 'short_window_share', 'mean_window'), feats.tolist())))
        if len(_CHOSEN) >= 64:
            _CHOSEN.pop(next(iter(_CHOSEN)))
        _CHOSEN[key] = hit
    return hit if with_features else hit[0]


_RUNS_PLANS = {}  # (kind, lags, n_frames, device) -> (pinned host plan, device plan, upload event)


def _host_plan(lib, kind, lags_host, n_frames, device):
    """The packed plan of a host-planned moment kernel ('win' or 'runs') for a lag grid, planned once and kept on the
    device: repeated analyses of trajectories of one shape (and the blocks of a staged ingest) only launch.  The plan
    depends on (lags, n_frames) only."""
    key = (kind, lags_host.tobytes(), int(n_frames), device.type, device.index)
    hit = _RUNS_PLANS.get(key)
    stream = current_stream(device)
    if hit is not None:
        plan_host, plan_dev, uploaded = hit
        if uploaded is not None:
            if uploaded.query():
                _RUNS_PLANS[key] = (plan_host, plan_dev, None)  # landed: later hits need not ask again
            else:  # uploaded on another stream and still in flight
                stream.wait_event(uploaded)
        plan_dev.record_stream(stream)
        return plan_host, plan_dev
    if kind == 'tile':
        nbytes = lib.kb2_self_moments_tile_workspace_bytes(lags_host.ctypes.data, len(lags_host), int(n_frames))
        if nbytes == 0:
            raise ValueError('kb2_self_moments_tile_workspace_bytes: invalid lag grid')
    elif kind == 'win':
        nbytes = lib.kb2_self_moments_win_workspace_bytes(len(lags_host))
    else:
        nbytes = lib.kb2_self_moments_runs_workspace_bytes(len(lags_host), int(n_frames))
    plan_host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    pack = {'tile': lib.kb2_self_moments_tile_plan_pack, 'win': lib.kb2_self_moments_win_plan_pack,
            'runs': lib.kb2_self_moments_runs_plan_pack}[kind]
    check(pack(lags_host.ctypes.data, len(lags_host), int(n_frames), plan_host.data_ptr()), f'kb2_self_moments_{kind}_plan_pack')
    plan_dev = plan_host.to(device, non_blocking=True)
    uploaded = torch.cuda.Event()
    uploaded.record(stream)
    if len(_RUNS_PLANS) >= 32:
        _RUNS_PLANS.pop(next(iter(_RUNS_PLANS)))
    _RUNS_PLANS[key] = (plan_host, plan_dev, uploaded)
    return plan_host, plan_dev


def self_moments_windows(lags, n_frames):
    """The windows kb2_self_moments_win would use: [(first lag, spacing, number of lags), ...]."""
    lib = _lib.load()
    lags_host = np.ascontiguousarray(lags, dtype=np.int32)
    cap = max(1, len(lags_host))
    arrs = [np.zeros(cap, dtype=np.int32) for _ in range(3)]
    n = lib.kb2_self_moments_win_plan(lags_host.ctypes.data, len(lags_host), int(n_frames), arrs[0].ctypes.data,
                                      arrs[1].ctypes.data, arrs[2].ctypes.data, cap)
    return [tuple(int(a[i]) for a in arrs) for i in range(min(n, cap))]


def self_moments_tiles(lags, n_frames):
    """The windows kb2_self_moments_tile would use, [(first lag, row length = lag spacing, number of lags), ...], and
    the number of tile templates."""
    lib = _lib.load()
    import ctypes
    lags_host = np.ascontiguousarray(lags, dtype=np.int32)
    cap = max(1, len(lags_host))
    arrs = [np.zeros(cap, dtype=np.int32) for _ in range(3)]
    n_tiles = ctypes.c_int(0)
    n = lib.kb2_self_moments_tile_plan(lags_host.ctypes.data, len(lags_host), int(n_frames), arrs[0].ctypes.data,
                                       arrs[1].ctypes.data, arrs[2].ctypes.data, cap, ctypes.byref(n_tiles))
    return [tuple(int(a[i]) for a in arrs) for i in range(min(n, cap))], n_tiles.value


def self_moments_plan(lags, n_frames):
    """The runs kb2_self_moments_runs would use: ([(k0, delta, len, n_chg), ...], n_lags_in_runs_of_one); a run
    covers the lags k0 + i*delta + e_i with e_i stepping by sign(n_chg) at |n_chg| indices."""
    lib = _lib.load()
    import ctypes
    lags_host = np.ascontiguousarray(lags, dtype=np.int32)
    cap = max(1, len(lags_host))
    arrs = [np.zeros(cap, dtype=np.int32) for _ in range(4)]
    rest = ctypes.c_int(0)
    n = lib.kb2_self_moments_plan(lags_host.ctypes.data, len(lags_host), int(n_frames), arrs[0].ctypes.data,
                                  arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data, cap, ctypes.byref(rest))
    return [tuple(int(a[i]) for a in arrs) for i in range(n)], rest.value


def disp_moments(disp, mask=7, weights=None, coll_mask=0):
    """[4] = sum d^2, sum d^4, sum c, sum c^2 of one materialised displacement array [N, M, 3]."""
    lib = require_cuda()
    out = torch.empty(4, dtype=torch.float64, device=disp.device)
    check(
        lib.kb2_disp_moments(disp.data_ptr(), disp.shape[0], disp.shape[1], mask, _ptr(weights), coll_mask,
                             out.data_ptr(), _stream()), 'kb2_disp_moments')
    _count(3 if coll_mask else 2)
    return out


def lag_squares(dc, n_frames, lag, mask=7):
    """The squared displacements of one lag, [n_atoms * (n_frames - lag + 1)], from a component-major displacement
    array [n_atoms, 3, pitch] (first-observation term included); see kb2_lag_squares."""
    lib = require_cuda()
    out = torch.empty(dc.shape[0] * (n_frames - lag + 1), dtype=torch.float64, device=dc.device)
    check(lib.kb2_lag_squares(dc.data_ptr(), dc.shape[0], n_frames, dc.shape[2], int(lag), mask, out.data_ptr(), _stream()),
          'kb2_lag_squares')
    _count()
    return out


def disp_squares(disp, mask=7, weights=None, collective=False):
    """|disp[a][j]|^2 ([N * M]) or |sum_a w_a disp[a][j]|^2 ([M]) of one materialised array [N, M, 3]."""
    lib = require_cuda()
    n, m = disp.shape[0], disp.shape[1]
    out = torch.empty(m if collective else n * m, dtype=torch.float64, device=disp.device)
    check(lib.kb2_disp_squares(disp.data_ptr(), n, m, mask, _ptr(weights), int(bool(collective)), out.data_ptr(), _stream()),
          'kb2_disp_squares')
    _count()
    return out


def resample_means(values, n_draws, n_resamples, first_resample=0, seed=0, stream_id=0):
    """[n_resamples] means of n_draws values drawn with replacement from `values`; see kb2_resample_means."""
    lib = require_cuda()
    out = torch.empty(n_resamples, dtype=torch.float64, device=values.device)
    check(lib.kb2_resample_means(values.data_ptr(), values.numel(), int(n_draws), int(n_resamples), int(first_resample),
                                 int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id), out.data_ptr(), _stream()),
          'kb2_resample_means')
    _count()
    return out


def reblock(values):
    """[levels, 2] = (sum, sum of squared deviations) of every reblocking level of `values`; see kb2_reblock."""
    lib = require_cuda()
    n = values.numel()
    levels = lib.kb2_reblock_levels(n)
    ws = torch.empty(lib.kb2_reblock_workspace_bytes(n) // 8, dtype=torch.float64, device=values.device)
    stats = torch.empty((levels + 1, 2), dtype=torch.float64, device=values.device)
    check(lib.kb2_reblock(values.data_ptr(), n, ws.data_ptr(), stats.data_ptr(), _stream()), 'kb2_reblock')
    _count(levels + 1)
    return stats[:levels]


def moment_stats(main_sums, cnt, divisor, ngp_sums, gcnt, join=None):
    """[4, n] = n, v, s, ngp from the per-lag sums; see kb2_moment_stats.  `join`: a peer-memory channel
    (dist.moments_channel) -- ngp_sums is then still local to this rank and the kernel sums it over the ranks first, in
    place (kb2_moment_stats_joined)."""
    lib = require_cuda()
    n = cnt.numel()
    out = torch.empty((4, n), dtype=torch.float64, device=cnt.device)
    if join is not None:
        windows, world, rank, capacity, seq = join.next_join()
        check(
            lib.kb2_moment_stats_joined(main_sums.data_ptr(), cnt.data_ptr(), divisor.data_ptr(), ngp_sums.data_ptr(),
                                        gcnt.data_ptr(), n, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                        out[3].data_ptr(), windows, world, rank, capacity, seq, _stream()),
            'kb2_moment_stats_joined')
        _count()
        return out
    check(
        lib.kb2_moment_stats(main_sums.data_ptr(), cnt.data_ptr(), divisor.data_ptr(), ngp_sums.data_ptr(),
                             gcnt.data_ptr(), n, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                             out[3].data_ptr(), _stream()), 'kb2_moment_stats')
    _count()
    return out


def cov_build(v, n_o):
    lib = require_cuda()
    n = v.numel()
    cov = torch.empty((n, n), dtype=torch.float64, device=v.device)
    check(lib.kb2_cov_build(v.data_ptr(), n_o.data_ptr(), n, cov.data_ptr(), _stream()), 'kb2_cov_build')
    _count()
    return cov


def gls_prepare(evals, evecs, x, y, fit_intercept=True):
    lib = require_cuda()
    n = evals.numel()
    evecs = evecs.contiguous()
    proj = torch.empty((4, n), dtype=torch.float64, device=evals.device)
    theta = torch.empty(8, dtype=torch.float64, device=evals.device)  # g0, c0, AA, AB, BB, RA, RB, Q0
    check(
        lib.kb2_gls_prepare(evals.data_ptr(), evecs.data_ptr(), x.data_ptr(), y.data_ptr(), n, int(bool(fit_intercept)),
                            proj.data_ptr(), theta.data_ptr(), _stream()), 'kb2_gls_prepare')
    _count(2)
    return proj, theta


def spectral_prepare_max_n():
    return require_cuda().kb2_spectral_prepare_max_n()


def spectral_prepare(cov, x, y, cond_max, fit_intercept=True):
    """One-kernel spectral step (Jacobi) -> evals [n], proj [4, n], gls [8], logdet_term [1], sweeps [1]."""
    lib = require_cuda()
    n = cov.shape[0]
    dev = cov.device
    evals = torch.empty(n, dtype=torch.float64, device=dev)
    proj = torch.empty((4, n), dtype=torch.float64, device=dev)
    gls = torch.empty(8, dtype=torch.float64, device=dev)
    logdet_term = torch.empty(1, dtype=torch.float64, device=dev)
    info = torch.empty(1, dtype=torch.int32, device=dev)
    check(
        lib.kb2_spectral_prepare(cov.data_ptr(), x.data_ptr(), y.data_ptr(), n, float(cond_max),
                                 int(bool(fit_intercept)), evals.data_ptr(), proj.data_ptr(), gls.data_ptr(),
                                 logdet_term.data_ptr(), info.data_ptr(), _stream()), 'kb2_spectral_prepare')
    _count(2)
    return evals, proj, gls, logdet_term, info


def mvn_loglike(theta, proj, logdet_term):
    lib = require_cuda()
    theta = theta.contiguous()
    W, ndim = theta.shape
    out = torch.empty(W, dtype=torch.float64, device=theta.device)
    check(
        lib.kb2_mvn_loglike(theta.data_ptr(), W, ndim, proj.data_ptr(), proj.shape[1], logdet_term.data_ptr(),
                            out.data_ptr(), _stream()), 'kb2_mvn_loglike')
    _count()
    return out


def ensemble_sample(gls, ndim, n_walkers, n_steps, logdet_term, seed):
    lib = require_cuda()
    dev = gls.device
    ws = torch.empty(lib.kb2_sampler_workspace_bytes(n_steps, n_walkers), dtype=torch.uint8, device=dev)
    chain = torch.empty((n_steps, n_walkers, ndim), dtype=torch.float64, device=dev)
    logp = torch.empty((n_steps, n_walkers), dtype=torch.float64, device=dev)
    check(
        lib.kb2_ensemble_sample(gls.data_ptr(), ndim, n_walkers, n_steps, logdet_term.data_ptr(),
                                int(seed) & 0xFFFFFFFFFFFFFFFF, ws.data_ptr(),
                                chain.data_ptr(), logp.data_ptr(), _stream()), 'kb2_ensemble_sample')
    _count(3)
    return chain, logp


def fp64_peak(grid, iters, sink):
    lib = require_cuda()
    check(lib.kb2_fp64_peak(grid, iters, sink.data_ptr(), _stream()), 'kb2_fp64_peak')
    _count()


def stream_copy(src, dst):
    lib = require_cuda()
    check(lib.kb2_stream_copy(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()), 'kb2_stream_copy')
    _count()


def device_info():
    lib = require_cuda()
    import ctypes
    vals = [ctypes.c_int(0) for _ in range(4)]
    check(lib.kb2_device_info(*[ctypes.byref(v) for v in vals]), 'kb2_device_info')
    return {'sm_count': vals[0].value, 'max_smem_optin': vals[1].value, 'cc': (vals[2].value, vals[3].value)}


class _PinnedRing:
    """Page-locked staging ring for small host->device copies.  cudaMemcpyAsync from PAGEABLE memory synchronises the
    stream before it starts, which stops the host from running ahead of the device and exposes every launch latency
    of the step; a copy from this ring is enqueued like a kernel.  A slot is reused one lap later."""

    def __init__(self, nbytes=8 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.view = self.buf.numpy()
        self.size = nbytes
        self.off = 0

    def put(self, a, device):
        n = a.nbytes
        start = (self.off + 63) & ~63
        if start + n > self.size:
            # wrap: everything queued from this lap must have left the ring before it is overwritten
            # (any stream may have used the ring: a device-wide wait, once per lap of 8 MB of small arrays)
            torch.cuda.synchronize(device)
            start = 0
        self.view[start:start + n] = a.reshape(-1).view(np.uint8)
        self.off = start + n
        src = self.buf[start:start + n].view(torch.from_numpy(a).dtype).reshape(a.shape)
        out = torch.empty(a.shape, dtype=src.dtype, device=device)
        out.copy_(src, non_blocking=True)
        return out


_RINGS = {}
_PINNED_POOL = {}  # size -> free page-locked buffers (cudaHostAlloc is slow: results buffers are recycled)


def pinned_take(nbytes):
    """A page-locked uint8 buffer of at least `nbytes` (rounded up to 64 KiB) from the pool."""
    size = (int(nbytes) + 65535) & ~65535
    free = _PINNED_POOL.setdefault(size, [])
    return free.pop() if free else torch.empty(size, dtype=torch.uint8).pin_memory()


def pinned_give(buf):
    free = _PINNED_POOL.setdefault(buf.numel(), [])
    if len(free) < 8:
        free.append(buf)


def small_to_device(a, device, np_dtype=np.float64):
    """Small host array -> CUDA tensor through a pinned staging ring: enqueued on the current stream, no stream
    synchronisation (see _PinnedRing)."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np_dtype))
    device = torch.device(device)
    if a.nbytes == 0 or a.nbytes > (1 << 20):
        return torch.from_numpy(a).to(device, non_blocking=True)
    key = (device.type, device.index)
    ring = _RINGS.get(key)
    if ring is None:
        ring = _RINGS[key] = _PinnedRing()
    return ring.put(a, device)


_CONST_CACHE = {}  # (bytes, dtype, shape, device) -> (tensor, upload event); read-only inputs only


def const_to_device(a, device, np_dtype=np.float64):
    """As small_to_device for arrays the kernels only READ (index lists, grids, counts, lattices): identical content is
    uploaded once and shared afterwards -- repeated analyses of trajectories of one shape upload nothing.  The caller must
    not write to the returned tensor."""
    a = np.ascontiguousarray(np.asarray(a, dtype=np_dtype))
    if a.nbytes == 0 or a.nbytes > (64 << 10):
        return small_to_device(a, device, np_dtype)
    device = torch.device(device)
    key = (a.tobytes(), a.dtype.str, a.shape, device.type, device.index)
    hit = _CONST_CACHE.get(key)
    if hit is not None:
        t, uploaded = hit
        stream = current_stream(device)
        if uploaded is not None:
            if uploaded.query():
                _CONST_CACHE[key] = (t, None)  # landed: later hits need not ask again
            else:  # uploaded on another stream and still in flight
                stream.wait_event(uploaded)
        t.record_stream(stream)  # (an evicted entry must not be reused by its allocation stream under this reader)
        return t
    t = small_to_device(a, device, np_dtype)
    uploaded = torch.cuda.Event()
    uploaded.record(current_stream(device))
    if len(_CONST_CACHE) >= 256:
        _CONST_CACHE.pop(next(iter(_CONST_CACHE)))
    _CONST_CACHE[key] = (t, uploaded)
    return t


def to_device(a, device, dtype=None):
    """numpy / torch -> contiguous CUDA tensor (float32 stays float32 unless dtype is given)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    if t.device.type == 'cpu' and not t.is_pinned() and 0 < t.numel() * t.element_size() <= (1 << 20):
        return small_to_device(t.contiguous().numpy(), device, t.contiguous().numpy().dtype)
    return t.to(device, non_blocking=True).contiguous()
