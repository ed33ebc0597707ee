"""The joins through peer memory (csrc/k7_peer.cu) against NCCL: arrays of the sizes the path exchanges ([K][2] moment sums,
[3][pitch] frame sums) and odd ones, issued on several streams at once and more of them in flight than the window has slots;
every rank must hold the same bits, equal to the float64 sum in rank order, and close to NCCL's all-reduce.  Then the time
of one join beside NCCL's.  Run under torchrun:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_check.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as td
from kinisi_b200 import dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
td.init_process_group('nccl', device_id=dev)
assert dist.active()


def local_array(n, seed, r):
    g = torch.Generator(device='cpu')
    g.manual_seed(1000 * seed + r)
    return torch.randn(n, dtype=torch.float64, generator=g) * (1.0 + r)


sizes = [1, 2, 3, 200, 2047, 2048, 2049, 4097, 60000, 60048, 65535, 150048, 300048]
streams = [torch.cuda.Stream(dev) for _ in range(4)]
worst_nccl, calls = 0.0, 0
for rep in range(3):
    outs = []
    for i, n in enumerate(sizes * 2):  # 26 joins in flight on four streams: more than the eight slots of a window
        seed = rep * 100 + i
        with torch.cuda.stream(streams[i % 4]):
            x = local_array(n, seed, rank).to(dev)
            y = x.clone()
            dist.allreduce_sum_(x, 'frames' if n > 1000 else 'moments')
            outs.append((n, seed, x, y))
            calls += 1
    torch.cuda.synchronize()
    dist.check_peer_status()
    for n, seed, x, y in outs:
        want = local_array(n, seed, 0)
        for r in range(1, world):
            want = want + local_array(n, seed, r)  # the order of the kernel: rank 0 first
        got = x.cpu()
        assert torch.equal(got, want), (n, seed, float((got - want).abs().max()))
        td.all_reduce(y)
        worst_nccl = max(worst_nccl, float(((y.cpu() - got).abs() / (got.abs() + 1e-300)).clamp(max=1.0).max()) if n > 3 else 0.0)
        ref = got.to(dev).clone()
        td.broadcast(ref, 0)
        assert torch.equal(ref.cpu(), got), 'ranks differ'
assert ('frames', local) in dist._PEER and dist._PEER[('frames', local)], 'the peer-memory road was not taken'


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    td.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n, host * 1e6


lines = []
for n in (200, 60048, 300048):
    x = torch.ones(n, dtype=torch.float64, device=dev)
    dev_us, host_us = timeit(lambda: dist.allreduce_sum_(x, 'frames'))
    os.environ['KB2_PEER_AR'] = '0'
    d2, h2 = timeit(lambda: dist.allreduce_sum_(x, 'frames'))
    os.environ['KB2_PEER_AR'] = '1'
    lines.append(f'  n = {n:7d} doubles: peer kernel {dev_us:6.1f} us per join back to back (host {host_us:5.1f} us per call); '
                 f'NCCL on the caller\'s stream {d2:6.1f} us (host {h2:5.1f} us)')
td.barrier()
if rank == 0:
    print(f'peer_check ok: world {world}, {calls} joins on 4 streams bit-identical on all ranks and equal to the rank-ordered sum; '
          f'worst relative difference from NCCL {worst_nccl:.1e}')
    print('\n'.join(lines))
td.destroy_process_group()
