"""ctypes binding of include/kinisi_b200.h.  There is no CPU fallback: if the shared library is missing
the import fails loudly."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_uint64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('KB2_LIB_PATH') or os.path.join(HERE, 'libkinisi_b200.so')  # (the override: A/B probes of two builds)

F32, F64 = 0, 1
MODE_FRACTIONAL, MODE_DISPLACEMENT = 0, 1
ABI_VERSION = 11

# name -> (restype, argtypes); mirrors include/kinisi_b200.h one to one
SIGNATURES = {
    'kb2_abi_version': (c_int, []),
    'kb2_last_error': (c_char_p, []),
    'kb2_device_info': (c_int, [POINTER(c_int)] * 4),
    'kb2_invert_lattice': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'kb2_frames_to_atoms': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    'kb2_frames_to_atoms_fractional': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int64, c_int64,
                                               c_int, c_void_p]),
    'kb2_molecule_centres': (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    'kb2_frame_sums': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                               c_void_p, c_int64, c_void_p]),
    'kb2_scan_frames': (c_int, [c_void_p, c_int, c_int64, c_int, c_double, c_void_p, c_double, c_void_p]),
    'kb2_unwrap_cumsum_workspace_bytes': (c_size_t, [c_int, c_int64]),
    'kb2_unwrap_cumsum': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                  c_void_p, c_double, c_void_p, c_int64, c_void_p, c_void_p]),
    'kb2_atom_sum': (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    'kb2_self_moments_workspace_bytes': (c_size_t, [c_int]),
    'kb2_self_moments': (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    'kb2_self_moments_runs_plan_pack': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'kb2_self_moments_runs_planned': (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p]),
    'kb2_self_moments_runs_workspace_bytes': (c_size_t, [c_int, c_int]),
    'kb2_self_moments_runs_blocks': (c_int, [c_int]),
    'kb2_self_moments_runs': (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p]),
    'kb2_self_moments_win_workspace_bytes': (c_size_t, [c_int]),
    'kb2_self_moments_win_blocks': (c_int, [c_int]),
    'kb2_self_moments_win_plan_pack': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'kb2_self_moments_win_planned': (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    'kb2_self_moments_win_plan': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int]),
    'kb2_self_moments_tile_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int]),
    'kb2_self_moments_tile_blocks': (c_int, [c_int]),
    'kb2_self_moments_tile_plan_pack': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'kb2_self_moments_tile_planned': (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p]),
    'kb2_self_moments_tile_plan': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_int)]),
    'kb2_self_moments_tile_plan_cost': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'kb2_self_moments_choose': (c_int, [c_void_p, c_int, c_int, c_void_p]),
    'kb2_self_moments_plan': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      POINTER(c_int)]),
    'kb2_disp_moments': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'kb2_moment_stats': (c_int, [c_void_p] * 5 + [c_int] + [c_void_p] * 4 + [c_void_p]),
    'kb2_cov_build': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'kb2_gls_prepare': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'kb2_spectral_prepare_max_n': (c_int, []),
    'kb2_spectral_prepare': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    'kb2_mvn_loglike': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'kb2_sampler_workspace_bytes': (c_size_t, [c_int, c_int]),
    'kb2_ensemble_sample': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    'kb2_lag_squares': (c_int, [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p, c_void_p]),
    'kb2_disp_squares': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'kb2_resample_means': (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_uint64, c_int, c_void_p, c_void_p]),
    'kb2_reblock_levels': (c_int, [c_int64]),
    'kb2_reblock_workspace_bytes': (c_size_t, [c_int64]),
    'kb2_reblock': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    'kb2_peer_slots': (c_int, []),
    'kb2_peer_window_bytes': (c_size_t, [c_int, c_int64]),
    'kb2_peer_window_create': (c_int, [c_int, c_int64, POINTER(c_void_p), c_void_p]),
    'kb2_peer_window_open': (c_int, [c_void_p, POINTER(c_void_p)]),
    'kb2_peer_window_close': (c_int, [c_void_p]),
    'kb2_peer_window_destroy': (c_int, [c_void_p]),
    'kb2_peer_allreduce_f64': (c_int, [c_void_p, c_int64, POINTER(c_void_p), c_int, c_int, c_int64, c_uint64, c_void_p]),
    'kb2_peer_rows_slices': (c_int, [c_int64]),
    'kb2_peer_allreduce_rows_f64': (c_int, [c_void_p, c_int, c_int64, POINTER(c_void_p), c_int, c_int, c_int64, c_uint64,
                                            c_void_p]),
    'kb2_frame_sums_joined': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_int64, POINTER(c_void_p), c_int, c_int, c_int64, c_uint64, c_void_p, c_void_p]),
    'kb2_moment_stats_joined': (c_int, [c_void_p] * 5 + [c_int] + [c_void_p] * 4 + [POINTER(c_void_p), c_int, c_int, c_int64, c_uint64,
                                                                               c_void_p]),
    'kb2_peer_status': (c_int, []),
    'kb2_fp64_peak': (c_int, [c_int, c_int, c_void_p, c_void_p]),
    'kb2_stream_copy': (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class KinisiB200Error(RuntimeError):
    pass


def load():
    """Load libkinisi_b200.so (built by kinisi_b200._build.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f'{LIB_PATH} is missing: build the CUDA extension first '
                          '(python -m kinisi_b200._build). kinisi_b200 has no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the library is stale
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.kb2_abi_version() != ABI_VERSION:
        raise ImportError(f'{LIB_PATH}: ABI version {lib.kb2_abi_version()} != {ABI_VERSION}; rebuild it')
    _lib = lib
    return lib


def check(rc, fn):
    if rc != 0:
        msg = load().kb2_last_error().decode('utf-8', 'replace')
        raise KinisiB200Error(f'{fn} failed ({rc}): {msg}')
