"""What read bandwidth do simple streaming kernels reach on this GPU?  (calibration for the GroupNorm kernels)"""
import torch
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for mb in (113, 226, 1024):
    x = torch.randn(mb * 1024 * 1024 // 2, device=dev).bfloat16()
    y = torch.empty_like(x)
    # rotate over several buffers so nothing is L2 resident
    xs = [x.clone() for _ in range(3)]
    i = [0]
    def s():
        i[0] = (i[0] + 1) % 3
        return xs[i[0]].sum(dtype=torch.float32)
    def amax():
        i[0] = (i[0] + 1) % 3
        return xs[i[0]].amax()
    def cp():
        i[0] = (i[0] + 1) % 3
        y.copy_(xs[i[0]])
    def silu():
        i[0] = (i[0] + 1) % 3
        torch.nn.functional.silu(xs[i[0]], inplace=False)
    for name, fn, traffic in (("sum", s, 1), ("amax", amax, 1), ("copy", cp, 2), ("silu", silu, 2)):
        us = timeit(fn)
        print(f"{name:5s} {mb:5d} MB: {us:8.1f} us  {traffic*mb*1.048576/us*1e3/1e3:6.2f} TB/s", flush=True)
