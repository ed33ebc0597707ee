"""Multi-GPU parity (plan item v): the same trajectory analysed on one rank and sharded by atom over
WORLD_SIZE ranks (NCCL) must agree to reduction-order tolerance.  Run under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from kinisi_b200 import diffusion, dist
from kinisi_b200.parser import Parser
from oracle import kinisi_oracle as ko  # (checker only: seeded input generator)

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
td.init_process_group('nccl', device_id=dev)
n_mob, n_fw, T = 96, 40, 6000
frac, latt = ko.synthetic_walk(n_mob, T, box=14.0, sigma=0.1, seed=17, n_framework=n_fw, dtype=np.float32)
q_all = np.where(np.arange(n_mob) % 3 == 0, -2.0, 1.0)
results = {}
for mode in ('single', 'sharded'):
    for source in ('device', 'host'):
        if mode == 'single':
            rows = np.arange(n_mob + n_fw)
            idx, fw, q, distributed = np.arange(n_mob), np.arange(n_mob, n_mob + n_fw), q_all, False
        else:
            m_lo, m_hi = dist.shard_bounds(n_mob)
            f_lo, f_hi = dist.shard_bounds(n_fw)
            rows = np.concatenate([np.arange(m_lo, m_hi), n_mob + np.arange(f_lo, f_hi)])
            idx, fw = np.arange(m_hi - m_lo), (m_hi - m_lo) + np.arange(f_hi - f_lo)
            q, distributed = q_all[m_lo:m_hi], True
        coords = torch.from_numpy(np.ascontiguousarray(frac[rows]))
        coords = coords.to(dev) if source == 'device' else coords.pin_memory()
        if mode == 'single' and rank != 0:
            continue  # the single-rank reference runs on rank 0 only (no collectives inside: distributed=False)
        p = Parser(Parser.get_disp(coords, latt), idx, fw, 0.001, 1, n_steps=60, progress=False, distributed=distributed)
        out = {}
        for kind, cls, args in (('msd', diffusion.MSDBootstrap, ()), ('mstd', diffusion.MSTDBootstrap, ()),
                                ('mscd', diffusion.MSCDBootstrap, (q, ))):
            b = cls(p.delta_t, p.disp_3d, *args, p._n_o, progress=False)
            out[kind] = (b.n.copy(), b.v.copy(), b.ngp.copy())
        b = diffusion.MSDBootstrap(p.delta_t, p.disp_3d, p._n_o, progress=False)
        b.diffusion(float(b.dt[6]), progress=False, random_state=np.random.RandomState(2))
        out['D'] = (b.D.n, b.D.samples.std(), b._theta_ml.cpu().numpy())
        results[(mode, source)] = out
td.barrier()
if rank == 0:
    ref = results[('single', 'device')]
    worst = 0.0
    for key in (('single', 'host'), ('sharded', 'device'), ('sharded', 'host')):
        got = results[key]
        for kind in ('msd', 'mstd', 'mscd'):
            for a, b_ in zip(ref[kind][:2], got[kind][:2]):
                worst = max(worst, float(np.max(np.abs(a - b_) / np.abs(a))))
            np.testing.assert_allclose(got[kind][2], ref[kind][2], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(got['D'][2][[0, 1, 2, 3, 4, 7]], ref['D'][2][[0, 1, 2, 3, 4, 7]], rtol=1e-8)
        assert abs(got['D'][0] - ref['D'][0]) < 0.5 * ref['D'][1]
    assert worst < 1e-11, worst
    print(f'dist_check ok: world {world}, worst relative difference of n / v over MSD, MSTD, MSCD = {worst:.2e}; '
          f'D single {ref["D"][0]:.6g} sharded {results[("sharded", "device")]["D"][0]:.6g}')
td.destroy_process_group()
