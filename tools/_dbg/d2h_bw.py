import time, torch
n = 32 << 20
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = torch.empty(n, dtype=torch.uint8).pin_memory()
for nchunk in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(nchunk)]
    ev = torch.cuda.Event()
    def run():
        cur = torch.cuda.current_stream()
        ev.record(cur)
        step = n // nchunk
        for i, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
        for s in streams:
            s.synchronize()
    for _ in range(5): run()
    best = 1e9
    for _ in range(30):
        torch.cuda.synchronize(); t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0; best = min(best, dt)
    print(f"chunks={nchunk}: {1e3*best:.3f} ms  {n/best/1e9:.1f} GB/s")
# plain single copy on the current stream
best = 1e9
for _ in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.current_stream().synchronize(); dt = time.perf_counter() - t0; best = min(best, dt)
print(f"single copy_: {1e3*best:.3f} ms  {n/best/1e9:.1f} GB/s")
