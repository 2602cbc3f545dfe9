import torch, time
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for mb in (39, 157, 630, 2500):
    n = mb * 1000 * 1000 // 4
    x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")
    us_fill = t(lambda: x.fill_(1.0)); us_copy = t(lambda: y.copy_(x)); us_read = t(lambda: x.sum())
    print(f"{mb:5d} MB: fill {us_fill:7.1f} us = {mb/us_fill*1e3:6.0f} GB/s write | copy {us_copy:7.1f} us = {2*mb/us_copy*1e3:6.0f} GB/s | sum {us_read:7.1f} us = {mb/us_read*1e3:6.0f} GB/s read")
