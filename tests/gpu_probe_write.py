"""How fast can this GPU absorb pure writes (L2-resident and HBM-sized)?  Calibrates the GEMM epilogue's store rate."""
import torch
dev = torch.device("cuda:0")
for mb in (8, 16, 32, 64, 128, 512, 2048):
    x = torch.empty(mb * 1024 * 1024 // 2, device=dev, dtype=torch.bfloat16)
    y = torch.empty_like(x)
    for name, fn in (("fill", lambda: x.fill_(1.0)), ("copy", lambda: y.copy_(x))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"{name} {mb:5d} MB: {us:8.1f} us  {mb*1.048576/us*1e3:8.1f} GB/s written", flush=True)
