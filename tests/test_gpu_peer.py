"""Two ranks on ONE GPU (the box `pytest -m gpu` runs on has a single B200): the joins over ranks through peer memory
(csrc/k7_peer.cu, kinisi_b200.dist._PeerChannel) -- CUDA IPC windows, flags, slot credits -- and an atom-sharded analysis
against the same analysis on one rank.  The process group (gloo) is only the rendezvous for the IPC handles; NCCL refuses two
ranks on one device, the peer-memory kernel does not care.  tools/peer_check.py and tools/dist_check.py are the same checks
on 2 and 8 GPUs (profiles/r02_peer_check_n*.txt, r02_dist_check_n*.txt)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _local_array(n, seed, r):
    g = torch.Generator(device='cpu')
    g.manual_seed(1000 * seed + r)
    return torch.randn(n, dtype=torch.float64, generator=g) * (1.0 + r)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['KB2_PEER_TIMEOUT_S'] = '20'
    import torch.distributed as td
    from kinisi_b200 import diffusion, dist
    from kinisi_b200.parser import Parser
    from oracle import kinisi_oracle as ko  # (checker only: seeded input generator)
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # (1) joins of the sizes the path exchanges and odd ones, on three streams, more in flight than a window has slots
        sizes = [1, 2, 3, 200, 1023, 1024, 1025, 4097, 60048, 65535]
        streams = [torch.cuda.Stream(dev) for _ in range(3)]
        for rep in range(2):
            outs = []
            for i, n in enumerate(sizes * 2):
                seed = rep * 100 + i
                with torch.cuda.stream(streams[i % 3]):
                    x = _local_array(n, seed, rank).to(dev)
                    dist.allreduce_sum_(x, 'frames' if n > 1000 else 'moments')
                    outs.append((n, seed, x))
            torch.cuda.synchronize()
            dist.check_peer_status()
            for n, seed, x in outs:
                want = _local_array(n, seed, 0)
                for r in range(1, world):
                    want = want + _local_array(n, seed, r)  # the kernel adds in rank order
                assert torch.equal(x.cpu(), want), (n, seed)
        assert dist._PEER.get(('frames', 0)) and dist._PEER.get(('moments', 0)), 'the peer-memory road was not taken'
        # a [3][pitch] array is cut by columns (the cut of the fused frame-sum kernel)
        for pitch in (16, 1024, 1040, 3008, 20000):
            x = _local_array(3 * pitch, pitch, rank).reshape(3, pitch).to(dev)
            dist.allreduce_sum_(x, 'frames')
            want = _local_array(3 * pitch, pitch, 0)
            for r in range(1, world):
                want = want + _local_array(3 * pitch, pitch, r)
            assert torch.equal(x.cpu().reshape(-1), want), pitch
        # (2) one trajectory on one rank and sharded by atom over the ranks.  The framework frame sums are joined inside the
        # kernel that makes them: 20 framework atoms -> the bulk-copy kernel on both ranks, 6 -> the register kernel,
        # 1 -> rank 1 has none and joins the same call with the stand-alone kernel
        for n_fw, T in ((20, 3000), (6, 1500), (1, 2100)):
            n_mob = 48
            frac, latt = ko.synthetic_walk(n_mob, T, box=14.0, sigma=0.1, seed=17 + n_fw, n_framework=n_fw, dtype=np.float32)
            q_all = np.where(np.arange(n_mob) % 3 == 0, -2.0, 1.0)
            res = {}
            for mode in ('single', 'sharded'):
                if mode == 'single':
                    rows = np.arange(n_mob + n_fw)
                    idx, fw, q, kw = np.arange(n_mob), np.arange(n_mob, n_mob + n_fw), q_all, {}
                else:
                    m_lo, m_hi = dist.shard_bounds(n_mob)
                    f_lo, f_hi = dist.shard_bounds(n_fw)
                    rows = np.concatenate([np.arange(m_lo, m_hi), n_mob + np.arange(f_lo, f_hi)]).astype(np.int64)
                    idx, fw = np.arange(m_hi - m_lo), (m_hi - m_lo) + np.arange(f_hi - f_lo)
                    q, kw = q_all[m_lo:m_hi], dict(distributed=True, global_counts=(n_mob, n_fw))
                coords = torch.from_numpy(np.ascontiguousarray(frac[rows])).to(dev)
                seq0 = dist._PEER[('frames', 0)].seq
                p = Parser(Parser.get_disp(coords, latt), idx, fw, 0.001, 1, n_steps=40, progress=False, **kw)
                if mode == 'sharded':
                    assert dist._PEER[('frames', 0)].seq == seq0 + 1  # one join, fused into the frame-sum kernel where there are atoms
                out = {}
                for kind, cls, args in (('msd', diffusion.MSDBootstrap, ()), ('mstd', diffusion.MSTDBootstrap, ()),
                                        ('mscd', diffusion.MSCDBootstrap, (q, ))):
                    b = cls(p.delta_t, p.disp_3d, *args, p._n_o, progress=False)
                    out[kind] = np.stack([b.n, b.v])
                res[mode] = out
            torch.cuda.synchronize()
            dist.check_peer_status()
            for kind in ('msd', 'mstd', 'mscd'):
                np.testing.assert_allclose(res['sharded'][kind], res['single'][kind], rtol=1e-10, err_msg=f'{kind} n_fw={n_fw}')
        np.save(os.path.join(out_dir, f'ok{rank}.npy'), np.array([1]))
    finally:
        td.destroy_process_group()


def test_two_ranks_join_through_peer_memory(tmp_path):
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    world = 2
    ctx = mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=False)
    import time
    deadline = time.time() + 240
    try:
        while not ctx.join(timeout=5):  # raises if a rank failed
            assert time.time() < deadline, 'the ranks did not finish'
    finally:
        for proc in ctx.processes:
            if proc.is_alive():
                proc.kill()
    for r in range(world):
        assert os.path.exists(tmp_path / f'ok{r}.npy'), f'rank {r} failed'


def _absent_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['KB2_PEER_TIMEOUT_S'] = '2'
    import torch.distributed as td
    from kinisi_b200 import dist
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    td.init_process_group('gloo', rank=rank, world_size=world)
    try:
        x = torch.ones(300, dtype=torch.float64, device=dev)
        dist.allreduce_sum_(x, 'moments')  # both ranks: the windows exist and work
        torch.cuda.synchronize()
        assert float(x[0]) == world
        td.barrier()
        if rank == 0:
            y = torch.ones(300, dtype=torch.float64, device=dev)
            dist.allreduce_sum_(y, 'moments')  # rank 1 never issues this one
            torch.cuda.synchronize()  # returns after KB2_PEER_TIMEOUT_S, not never
            raised = False
            try:
                dist.check_peer_status()
            except RuntimeError as exc:
                raised = 'timed out waiting for a peer' in str(exc)
            assert raised, 'the missing peer went unnoticed'
        np.save(os.path.join(out_dir, f'absent{rank}.npy'), np.array([1]))
        td.barrier()
    finally:
        td.destroy_process_group()


def test_a_peer_that_never_arrives_is_reported_not_waited_for(tmp_path):
    """Every wait of the join kernels is bounded: a rank that does not issue its side of a join makes the others' kernel
    give up after KB2_PEER_TIMEOUT_S and raise a status word, instead of hanging the device."""
    import time
    world = 2
    ctx = mp.spawn(_absent_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=False)
    deadline = time.time() + 120
    try:
        while not ctx.join(timeout=5):
            assert time.time() < deadline, 'the ranks did not finish'
    finally:
        for proc in ctx.processes:
            if proc.is_alive():
                proc.kill()
    for r in range(world):
        assert os.path.exists(tmp_path / f'absent{r}.npy'), f'rank {r} failed'
