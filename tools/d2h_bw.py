import torch, time
n = 1813512192 // 4
d = torch.empty(n, device="cuda"); h = torch.empty(n).pin_memory()
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("contiguous D2H GB/s", n * 4 / e0.elapsed_time(e1) / 1e6)
