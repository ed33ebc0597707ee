"""Host-side H2D bandwidth against CPU affinity: which NUMA node the pinned staging memory lands on decides the PCIe rate."""
import os, time, torch, pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
print('cpus', os.cpu_count(), 'affinity now', len(os.sched_getaffinity(0)))
try:
    n = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    cpus = [i * 64 + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
    print('gpu0 cpu affinity', cpus[:4], '...', cpus[-4:], len(cpus))
except Exception as e:
    print('affinity query failed', e); cpus = None
for node in sorted(os.listdir('/sys/devices/system/node')) if os.path.isdir('/sys/devices/system/node') else []:
    if node.startswith('node'):
        print(node, open(f'/sys/devices/system/node/{node}/cpulist').read().strip())
print(os.popen('nvidia-smi topo -m 2>/dev/null | head -12').read())

def bw(tag):
    x = torch.empty(368 << 20, dtype=torch.uint8).pin_memory()
    x.fill_(1)
    d = torch.empty_like(x, device='cuda')
    for _ in range(2): d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(5): d.copy_(x, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(tag, 'GB/s', 5 * x.numel() / a.elapsed_time(b) / 1e6)
    del x, d

torch.cuda.init()
bw('default')
all_cpus = sorted(os.sched_getaffinity(0))
if cpus:
    os.sched_setaffinity(0, set(cpus) & set(all_cpus) or set(all_cpus))
    bw('gpu-local cpus')
    other = set(all_cpus) - set(cpus)
    if other:
        os.sched_setaffinity(0, other)
        bw('remote cpus')
