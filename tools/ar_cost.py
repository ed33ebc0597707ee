"""Host and device cost of one small all-reduce through torch.distributed (run under torchrun)."""
import os, time, torch, torch.distributed as td
rank = int(os.environ['RANK']); torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
td.init_process_group('nccl')
for n in (200, 60006):
    t = torch.ones(n, dtype=torch.float64, device='cuda')
    for _ in range(20): td.all_reduce(t)
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(200): td.all_reduce(t)
    host = time.perf_counter() - t0
    z.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f'n={n}: host {host / 200 * 1e6:.1f} us per call, device {a.elapsed_time(z) / 200 * 1e3:.1f} us per call')
td.destroy_process_group()
