import torch, time
x = torch.empty(1_200_000_000, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print("memset  4.8 GB: %.3f ms  %.2f TB/s write" % (ms, 4.8 / ms))
ms = t(lambda: x.fill_(1.5)); print("fill    4.8 GB: %.3f ms  %.2f TB/s write" % (ms, 4.8 / ms))
ms = t(lambda: y.copy_(x)); print("copy  2x4.8 GB: %.3f ms  %.2f TB/s r+w" % (ms, 9.6 / ms))
ms = t(lambda: x.sum()); print("read    4.8 GB: %.3f ms  %.2f TB/s read" % (ms, 4.8 / ms))
