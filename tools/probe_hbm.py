"""Pure-write, pure-read and copy HBM bandwidth with plain torch ops (what bounds the training tile stores)."""
import torch
n = 2 << 30      # 8 GB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
w = t(lambda: a.zero_())
r = t(lambda: a.sum())
c = t(lambda: b.copy_(a))
print(f"write {4 * n / w / 1e9:.0f} GB/s   read {4 * n / r / 1e9:.0f} GB/s   copy {8 * n / c / 1e9:.0f} GB/s (read+write)")
