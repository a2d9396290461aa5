import torch
for mb in (32, 256, 1024):
    n = mb << 20
    bufs = [torch.empty(n // 4, dtype=torch.float32, device="cuda") for _ in range(8 if mb <= 256 else 2)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("fill_", lambda b: b.fill_(1.5)), ("zero_", lambda b: b.zero_())):
        for b in bufs: fn(b)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0.record()
            for b in bufs: fn(b)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / len(bufs))
        print(f"{name} {mb:5d} MiB: {1e3*best:8.1f} us  {n/best/1e6:8.1f} GB/s")
a = torch.empty(256 << 18, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for _ in range(10):
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
print(f"copy 256 MiB: {1e3*best:.1f} us  {2*a.numel()*4/best/1e6:.1f} GB/s (read + write)")
