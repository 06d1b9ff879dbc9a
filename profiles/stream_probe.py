"""Streaming-bandwidth probe on the box: what plain read/write mixes reach on this B200 (GB/s, CUDA events)."""
import torch
n = 256 * 1024 * 1024  # floats (1 GiB per tensor)
a = torch.randn(n, device='cuda'); b = torch.randn(n, device='cuda'); c = torch.empty(n, device='cuda')
def timeit(f, bytes_, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return bytes_ / best / 1e6
print('copy   1R1W  %.0f GB/s' % timeit(lambda: c.copy_(a), 8 * n))
print('add    2R1W  %.0f GB/s' % timeit(lambda: torch.add(a, b, out=c), 12 * n))
print('read   1R    %.0f GB/s' % timeit(lambda: a.sum(), 4 * n))
print('fill   1W    %.0f GB/s' % timeit(lambda: c.fill_(1.0), 4 * n))
d = torch.randn(n, device='cuda')
print('addcmul 3R1W %.0f GB/s' % timeit(lambda: torch.addcmul(a, b, d, out=c), 16 * n))
