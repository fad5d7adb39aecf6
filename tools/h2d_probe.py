import torch, time
x = torch.empty(33177600 // 4 * 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in (1, 4):
    hs = [torch.empty_like(x).pin_memory() for _ in range(n)]
    ds = [torch.empty_like(x, device="cuda") for _ in range(n)]
    ss = [torch.cuda.Stream() for _ in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(10):
        for i in range(n):
            with torch.cuda.stream(ss[i]):
                ds[i].copy_(hs[i], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {n} stream(s): {10 * n * x.numel() * 4 / dt / 1e9:.1f} GB/s")
    t0 = time.perf_counter()
    for r in range(10):
        for i in range(n):
            with torch.cuda.stream(ss[i]):
                hs[i].copy_(ds[i], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"D2H {n} stream(s): {10 * n * x.numel() * 4 / dt / 1e9:.1f} GB/s")
