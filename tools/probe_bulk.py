"""cp.async.bulk (1-D TMA) global->shared throughput per SM: cycles per chunk for several ring depths /
piece counts, all SMs reading the same L2-resident 1.2 MB buffer (the weight-stream pattern)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_projects_b200 import _lib as L

lib = L.load()
src = torch.randint(0, 255, (1200 * 1024,), dtype=torch.uint8, device="cuda")
out = torch.zeros(148, dtype=torch.int64, device="cuda")
for nblocks in (148, 8):
    for chunk in (16384, 32768):
        for depth in (1, 2, 4, 6):
            for pieces in (1, 4, 16):
                if depth * chunk > 200 * 1024:
                    continue
                L.check(lib.nerfb_bulk_probe(src.data_ptr(), src.numel(), 400, chunk, pieces, depth, nblocks,
                                             out.data_ptr(), None), "probe")
                torch.cuda.synchronize()
                c = out[:nblocks].float()
                print(f"blocks {nblocks:3d} chunk {chunk:6d} depth {depth} pieces {pieces:2d}: {c.mean():8.0f} cyc/chunk "
                      f"(min {c.min():.0f} max {c.max():.0f})  -> {chunk / c.mean():6.1f} B/clk/SM")
