"""Write-only HBM bandwidth on this GPU (the ceiling kernel (b) works against when the covariances are
materialised): torch fill of a 4 GB buffer, CUDA events, best of 5."""
import json
import torch

x = torch.empty(1 << 29, dtype=torch.float64, device="cuda")   # 4 GiB
res = {}
for name, fn in (("fill", lambda: x.fill_(1.0)), ("zero", lambda: x.zero_())):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[name + "_GBps"] = x.numel() * 8 / best / 1e6
y = torch.empty_like(x)
y.copy_(x); torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
res["copy_read_plus_write_GBps"] = 2 * x.numel() * 8 / best / 1e6
print(json.dumps(res))
