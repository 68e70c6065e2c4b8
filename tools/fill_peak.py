import torch
dev="cuda:0"
x = torch.empty(10*1024**3, dtype=torch.uint8, device=dev)
y = torch.empty(5*1024**3, dtype=torch.uint8, device=dev)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms = t(lambda: x.fill_(7))
print("fill_ 10 GiB: %.3f ms  %.1f GB/s" % (ms, x.numel()/ms/1e6))
ms = t(lambda: x.zero_())
print("zero_ (memset) 10 GiB: %.3f ms  %.1f GB/s" % (ms, x.numel()/ms/1e6))
ms = t(lambda: y.copy_(x[:y.numel()]))
print("copy 5 GiB: %.3f ms  %.1f GB/s (r+w)" % (ms, 2*y.numel()/ms/1e6))
xf = x.view(torch.float32)
ms = t(lambda: xf.fill_(1.5))
print("fill_ f32 10 GiB: %.3f ms  %.1f GB/s" % (ms, x.numel()/ms/1e6))
