"""Timeline of the fused MLP kernel (block 0): prints per-event cycle offsets (debug aid).

The clock64 hooks are compiled in only with -DNERFB_TRACE=1; build and select that variant first:
    NERFB_VARIANT=trace NERFB_DEFINES="-DNERFB_TRACE=1" python -m nerf_projects_b200.build
    NERFB_VARIANT=trace python tools/trace_mlp.py [fwd|save|chain]"""
import sys, os
if os.environ.get("NERFB_VARIANT") != "trace":
    sys.exit("run with NERFB_VARIANT=trace (see the docstring): the product library has no trace hooks")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nerf_projects_b200 as nb
from nerf_projects_b200 import mlp, _lib
from nerf_projects_b200 import synthetic as SY

net = nb.NeRF(8, 256, 63, 27, 5, [4], True).cuda()
net.load_state_dict(SY.random_init_state_dict(0, use_viewdirs=True, output_ch=5))
R, S = 32768, 192
g = torch.Generator().manual_seed(0)
rays = torch.cat([torch.randn(R, 3, generator=g), torch.randn(R, 3, generator=g), torch.full((R, 1), 2.), torch.full((R, 1), 6.),
                  torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)], -1).cuda()
z = (torch.rand(R, S, generator=g) * 4 + 2).sort(-1)[0].cuda()
MODE = sys.argv[1] if len(sys.argv) > 1 else "fwd"      # fwd | save | chain
from nerf_projects_b200 import ops
n = R * S
if MODE != "fwd":
    R = 4096; n = R * S
    rays, z = rays[:R].contiguous(), z[:R].contiguous()
    acts = torch.empty(ops.train_store_bytes(n, 0), dtype=torch.uint8, device="cuda")
    gstore = torch.empty(ops.train_store_bytes(n, 1), dtype=torch.uint8, device="cuda")
    masks = torch.empty(ops.train_store_bytes(n, 2), dtype=torch.uint8, device="cuda")
    raw = torch.empty(n, 4, device="cuda")
    d_pad = torch.randn((n + 511) // 512 * 512, 4, device="cuda")
    packed, packed_bwd = mlp.packed_weights(net), mlp.packed_weights_bwd(net)
def run():
    if MODE == "fwd":
        mlp.forward_bf16(net, R, S, rays, z)
    elif MODE == "save":
        ops.mlp_fwd_save_bf16(R, S, rays, z, None, None, packed, raw, acts, masks)
    else:
        ops.mlp_bwd_chain_bf16(n, d_pad, packed_bwd, masks, gstore)
if MODE == "chain":
    ops.mlp_fwd_save_bf16(R, S, rays, z, None, None, packed, raw, acts, masks)
for _ in range(2):
    run()
torch.cuda.synchronize()
buf = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
_lib.load().nerfb_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
_lib.load().nerfb_set_trace(None)
b = buf.cpu().numpy().reshape(8, 4096)
MASK = (1 << 44) - 1
# clock alignment: every traced thread records tag 0xFFF right after the cluster barrier; shift the peer's clock
# so that its marker coincides with the leader's
def marker(region):
    for v in b[region]:
        if v and (int(v) >> 44) == 0xFFF:
            return int(v) & MASK
    return None
m0 = marker(0)
shift = {}
m4 = marker(4)
for region in range(8):
    shift[region] = (m0 - m4) if region >= 4 else 0
ev = []
names = ("LOAD", "MMA", "EPI", "OBS", "pLOAD", "pMMA", "pEPI", "pRELAY")
for region, name in enumerate(names):
    for v in b[region]:
        if v == 0 or (int(v) >> 44) == 0xFFF:
            continue
        ev.append(((int(v) & MASK) + shift[region], name, int(v) >> 44))
ev.sort()
starts = [t for t, n, tag in ev if n == "MMA" and tag == 0x300]
lo, hi = starts[2], starts[4]
print("iteration length (cycles):", [starts[i + 1] - starts[i] for i in range(len(starts) - 1)][:8])
last = {}
for t, n, tag in ev:
    if t < lo or t >= hi:
        continue
    kind = tag >> 8
    desc = {1: "load issued q=%d", 2: "relay: own full seen, arrive sent q=%d", 3: "act_ready waited op=%d",
            5: "weights ready op=%d", 6: "issued+commits op=%d", 7: "mma_done waited l,t=%d", 8: "epilogue end l,t=%d",
            9: "signalled t=%d", 10: "load copies issued q=%d", 11: "OBSERVER: weights barrier complete q=%d"}[kind] % (tag & 0xFF)
    print(f"{t - lo:8d}  +{t - last.get(n, t):6d}  {n:6s} {desc}")
    last[n] = t
