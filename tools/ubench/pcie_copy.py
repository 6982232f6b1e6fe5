import torch, time
for mb in (1, 8, 64, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for direction in ("d2h", "h2d"):
        for _ in range(3):
            torch.cuda.synchronize(); t = time.perf_counter()
            if direction == "d2h": h.copy_(d, non_blocking=True)
            else: d.copy_(h, non_blocking=True)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{direction} {mb:4d} MB: {dt*1e3:7.3f} ms  {n/dt/1e9:6.1f} GB/s")
