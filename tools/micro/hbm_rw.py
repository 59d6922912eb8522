"""Pure-read, pure-write and copy HBM bandwidth with plain torch kernels (context for the roofline denominators)."""
import torch
dev = torch.device('cuda', 0)
n = 1 << 30
x = torch.empty(n, dtype=torch.bfloat16, device=dev); y = torch.empty_like(x)
def t(fn, nbytes, name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f'{name:12s} {nbytes / best / 1e6:8.1f} GB/s')
t(lambda: x.fill_(1.0), 2 * n, 'write (fill)')
t(lambda: y.copy_(x), 4 * n, 'copy')
t(lambda: x.view(torch.int32).sum(), 2 * n, 'read (sum)')
