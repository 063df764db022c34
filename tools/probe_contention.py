"""Who competes for shared memory?  Bulk-copy stream (16 KB chunks, ring of 4) concurrent with back-to-back UMMAs
(SS: A,B from smem; TS: A from TMEM) and epilogue-like 16-byte stores.  Prints cycles per chunk / per MMA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_projects_b200 import _lib as L

lib = L.load()
src = torch.randint(0, 255, (1200 * 1024,), dtype=torch.uint8, device="cuda")
out = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
for mode, name in ((0, "no MMA"), (1, "SS MMA"), (2, "TS MMA")):
    for sts in (0, 8):
        out.zero_()
        L.check(lib.nerfb_contention_probe(src.data_ptr(), src.numel(), 2000, mode, sts, 148, out.data_ptr(), None), "probe")
        torch.cuda.synchronize()
        o = out.view(148, 4).float()
        print(f"{name:7s} sts_warps={sts}: {o[:,0].mean():7.0f} cyc/chunk ({16384 / o[:,0].mean():5.1f} B/clk)   "
              f"{o[:,1].mean():6.1f} cyc/MMA (floor 128)   stores/thread {o[:,3].mean():.0f}")
