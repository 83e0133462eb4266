import torch, time
for mb in (8, 64, 128):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(mb << 20, dtype=torch.uint8, device='cuda')
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            e0.record()
            if direction == "h2d": d.copy_(h, non_blocking=True)
            else: h.copy_(d, non_blocking=True)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(mb, "MB", direction, f"{mb / 1024 / (best * 1e-3):.1f} GB/s")
