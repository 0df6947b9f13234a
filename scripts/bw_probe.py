"""Raw HBM bandwidth probes with torch ops (denominators for the dequant roofline): fill, copy, 1:4 read:write mix."""
import torch
dev = torch.device("cuda")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for mb in (117, 512, 2048):
    n = mb * 1024 * 1024 // 2
    x = torch.empty(n, dtype=torch.float16, device=dev); y = torch.empty(n, dtype=torch.float16, device=dev)
    q = torch.empty(n // 4, dtype=torch.float16, device=dev)
    bufs = [torch.empty(n, dtype=torch.float16, device=dev) for _ in range(4)] if mb < 1024 else [y]
    i = [0]
    def fill():
        bufs[i[0] % len(bufs)].zero_(); i[0] += 1
    def copy():
        bufs[i[0] % len(bufs)].copy_(x); i[0] += 1
    def widen():      # reads n/4 elements of int32? emulate 1:4 read:write: y = repeat(q)
        torch.repeat_interleave(q, 4, output_size=n)
    ms = timeit(fill); print(f"fill  {mb} MB: {ms*1e3:.1f} us  write {n*2/ms/1e6:.0f} GB/s")
    ms = timeit(copy); print(f"copy  {mb} MB: {ms*1e3:.1f} us  r+w {2*n*2/ms/1e6:.0f} GB/s")
    s = timeit(lambda: x.sum()); print(f"read  {mb} MB: {s*1e3:.1f} us  read {n*2/s/1e6:.0f} GB/s")
