"""HBM write / copy bandwidth on this box with torch ops (CUDA events): the ceiling for the 407 MB of observation rows."""
import torch
dev = torch.device('cuda', 0)
n = 407437312 // 4
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty(n, dtype=torch.float32, device=dev)
big = torch.empty(1 << 28, dtype=torch.float32, device=dev)   # 1 GiB
big2 = torch.empty(1 << 28, dtype=torch.float32, device=dev)
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
t = timeit(lambda: x.fill_(1.0)); print(f'fill 407 MB: {t*1e3:.1f} us  {x.numel()*4/t/1e6:.0f} GB/s')
t = timeit(lambda: x.zero_()); print(f'zero 407 MB (memset): {t*1e3:.1f} us  {x.numel()*4/t/1e6:.0f} GB/s')
t = timeit(lambda: big.fill_(1.0)); print(f'fill 1 GiB: {t*1e3:.1f} us  {big.numel()*4/t/1e6:.0f} GB/s')
t = timeit(lambda: y.copy_(x)); print(f'copy 407 MB: {t*1e3:.1f} us  {2*x.numel()*4/t/1e6:.0f} GB/s (read+write)')
t = timeit(lambda: big2.copy_(big)); print(f'copy 1 GiB: {t*1e3:.1f} us  {2*big.numel()*4/t/1e6:.0f} GB/s (read+write)')
t = timeit(lambda: x.sum()); print(f'sum (read) 407 MB: {t*1e3:.1f} us  {x.numel()*4/t/1e6:.0f} GB/s')
