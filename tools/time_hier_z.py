import sys, torch
sys.path.insert(0, "/root/repo")
from nerf_projects_b200 import ops
R, S, N = 32768, 64, 128
g = torch.Generator().manual_seed(0)
z = (torch.rand(R, S, generator=g) * 4 + 2).sort(-1)[0].cuda()
w = (torch.rand(R, S, generator=g) ** 5).cuda()
u = torch.rand(R, N, generator=g).cuda()
for det in (True, False):
    for _ in range(3): ops.hierarchical_z(z, w, N, det=det, u=None if det else u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.hierarchical_z(z, w, N, det=det, u=None if det else u)
    e1.record(); torch.cuda.synchronize()
    print(f"hierarchical_z det={det}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per 32768-ray chunk")
