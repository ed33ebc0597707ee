"""The N>1 path on CPU: two gloo ranks, atoms sharded, per-frame framework sums and per-dt moment sums
joined by kinisi_b200.dist's all-reduce.  The per-rank partial sums come from the oracle here (there is
no GPU); the sharding, the all-reduce plumbing and the finalisation are the product's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as td
    from kinisi_b200 import dist
    from oracle import kinisi_oracle as ko
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n_mobile, n_fw, n_frames = 11, 5, 200
        frac, latt = ko.synthetic_walk(n_mobile, n_frames, box=9.0, sigma=0.2, seed=3, n_framework=n_fw)
        u = np.diff(np.concatenate([np.zeros((n_mobile + n_fw, 1, 3)), ko.get_disp(frac, latt)], axis=1), axis=1)
        # shard mobile atoms and framework atoms separately, as bench.py does
        m_lo, m_hi = dist.shard_bounds(n_mobile)
        f_lo, f_hi = dist.shard_bounds(n_fw)
        assert dist.world_size() == world and dist.rank() == rank
        n_sel = int(round(dist.allreduce_sum_scalar(m_hi - m_lo)))
        n_fwg = int(round(dist.allreduce_sum_scalar(f_hi - f_lo)))
        assert (n_sel, n_fwg) == (n_mobile, n_fw)
        # framework drift: per-frame sums of the local framework atoms, all-reduced, then scanned
        fw_sum = torch.from_numpy(u[n_mobile + f_lo:n_mobile + f_hi].sum(axis=0).T.copy())
        dist.allreduce_sum_(fw_sum, 'frames')
        drift = np.cumsum(fw_sum.numpy().T, axis=0) / n_fwg
        dc_local = np.cumsum(u[m_lo:m_hi], axis=1) - drift[None]
        lags = ko.time_intervals(n_frames, 0.001, 1, n_steps=25)
        S1, S2, _, _ = ko.raw_sums_streaming(dc_local, lags)
        sums = torch.from_numpy(np.stack([S1, S2], axis=1))
        dist.allreduce_sum_(sums, 'moments')
        # collective sum for MSTD
        S = torch.from_numpy(dc_local.sum(axis=0).T.copy())
        dist.allreduce_sum_(S, 'frames')
        if rank == 0:
            np.savez(os.path.join(out_dir, 'out.npz'), sums=sums.numpy(), S=S.numpy(), lags=lags)
    finally:
        td.destroy_process_group()


def test_two_rank_atom_sharding(tmp_path):
    from oracle import kinisi_oracle as ko
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / 'out.npz')
    n_mobile, n_fw, n_frames = 11, 5, 200
    frac, latt = ko.synthetic_walk(n_mobile, n_frames, box=9.0, sigma=0.2, seed=3, n_framework=n_fw)
    parsed = ko.parse(ko.get_disp(frac, latt), list(range(n_mobile)), list(range(n_mobile, n_mobile + n_fw)), 0.001, 1,
                      n_steps=25)
    S1, S2, _, _ = ko.raw_sums_streaming(parsed['dc_sel'], parsed['lags'])
    np.testing.assert_array_equal(out['lags'], parsed['lags'])
    np.testing.assert_allclose(out['sums'][:, 0], S1, rtol=1e-12)
    np.testing.assert_allclose(out['sums'][:, 1], S2, rtol=1e-12)
    np.testing.assert_allclose(out['S'].T, parsed['dc_sel'].sum(axis=0), rtol=1e-10, atol=1e-10)
    # finalisation as the product does it (kb2_moment_stats restated on the host)
    cnt = n_mobile * (n_frames - parsed['lags'] + 1.0)
    ref = ko.bootstrap_streaming(parsed, 'msd')
    n = out['sums'][:, 0] / cnt
    v = (out['sums'][:, 1] - out['sums'][:, 0] * n) / (cnt - 1) / parsed['n_o']
    np.testing.assert_allclose(n, ref['n'], rtol=1e-11)
    np.testing.assert_allclose(v, ref['v'], rtol=1e-10)
