import torch, time
dev = torch.device("cuda:0")
for mb in (49, 98, 196):
    n = mb * 1000 * 1000 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb} MB pinned: {ms:.3f} ms = {mb / ms:.1f} GB/s")
# 4 separate tensors like the bench (2 x 78.6 MB + 2 x 19.7 MB)
hs = [torch.empty(s, dtype=torch.float32).pin_memory() for s in (16*1228800, 16*1228800, 16*307200, 16*307200)]
ds = [torch.empty_like(h, device=dev) for h in hs]
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    for h, d in zip(hs, ds): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"bench-shaped upload (4 tensors, 196.6 MB): {ms:.3f} ms = {196.6 / ms:.1f} GB/s")
