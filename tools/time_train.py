"""Per-kernel CUDA-event timing of one training step's MLP backward (GPU box only).

    python tools/time_train.py [rays] [reps] [nv]       (nv: the network without view directions, out_ch 5)
Prints ms and achieved TFLOP/s / GB/s of: fused forward, forward-with-save, data-gradient chain,
weight-gradient kernel on a [rays x 192] fine-network batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nerf_projects_b200 as nb  # noqa: E402
from nerf_projects_b200 import mlp, ops  # noqa: E402
from nerf_projects_b200 import synthetic as SY  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
S = 192
dev = "cuda"
net = nb.NeRF(8, 256, 63, 27, 5, [4], True).to(dev)
net.load_state_dict(SY.random_init_state_dict(1, use_viewdirs=True, output_ch=5))
gen = torch.Generator().manual_seed(0)
rays = torch.cat([torch.randn(R, 3, generator=gen), torch.randn(R, 3, generator=gen), torch.full((R, 1), 2.0),
                  torch.full((R, 1), 6.0), torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=-1)],
                 -1).to(dev)
z = (torch.rand(R, S, generator=gen) * 4 + 2).sort(-1)[0].to(dev)
n = R * S
npad = (n + 511) // 512 * 512
d_pad = torch.zeros(npad, 4, device=dev)
d_pad[:n] = torch.randn(n, 4, generator=gen).to(dev)
packed, packed_bwd = mlp.packed_weights(net), mlp.packed_weights_bwd(net)
acts = torch.empty(ops.train_store_bytes(n, 0), dtype=torch.uint8, device=dev)
gstore = torch.empty(ops.train_store_bytes(n, 1), dtype=torch.uint8, device=dev)
masks = torch.empty(ops.train_store_bytes(n, 2), dtype=torch.uint8, device=dev)
raw = torch.empty(n, 4, device=dev)
grads = {k: torch.zeros_like(p) for k, p in net.named_parameters()}
gptrs = mlp._param_ptrs(grads)

legs = {
    "fwd": (lambda: ops._mlp_fwd_bf16(R, S, rays, z, None, None, True, 5, packed, raw), 1186816, 20),
    "fwd_save": (lambda: ops.mlp_fwd_save_bf16(R, S, rays, z, None, None, packed, raw, acts, masks), 1186816, 5 * 1024 + 20),
    "bwd_chain": (lambda: ops.mlp_bwd_chain_bf16(n, d_pad, packed_bwd, masks, gstore), 2 * 557696,
                  9 * 32 + 38 * 128 + 16),
    "wgrad": (lambda: ops.mlp_wgrad_bf16(n, acts, gstore, d_pad, gptrs), 1186816, 10.4 * 1024),
}
# the data-gradient chain + weight gradients in one cooperative launch, for several SM splits (chain SMs of 148)
for cs in ((60, 68, 76, 84, 92) if R >= 2048 and "--no-fused" not in sys.argv else ()):
    legs[f"bwd_fused_{cs}"] = ((lambda cs=cs: ops.mlp_bwd_fused_bf16(n, d_pad, packed_bwd, masks, acts, gstore, gptrs, cs)),
                               2 * 557696 + 1186816, 9 * 32 + 38 * 128 + 16 + 5 * 1024 + 300)
if len(sys.argv) > 3 and sys.argv[3] == "nv":
    # network without view directions: output_linear head (492 544 MACs forward, 7 chain layers, 33 + 32 stored panels)
    net = nb.NeRF(8, 256, 63, 0, 5, [4], False).to(dev)
    net.load_state_dict(SY.random_init_state_dict(1, use_viewdirs=False, input_ch_views=0, output_ch=5))
    packed, packed_bwd = mlp.packed_weights(net), mlp.packed_weights_bwd(net)
    raw5 = torch.empty(n, 5, device=dev)
    d4 = torch.zeros(npad, device=dev)
    gptrs_nv = mlp._param_ptrs(mlp._zero_grads(net))
    rays8 = rays[:, :8].contiguous()
    fl = 2 * (63 * 256 + 4 * 65536 + 319 * 256 + 2 * 65536 + 256 * 5)
    legs = {
        "fwd": (lambda: ops._mlp_fwd_bf16(R, S, rays8, z, None, None, False, 5, packed, raw5), fl, 24),
        "fwd_save": (lambda: ops.mlp_fwd_save_nv_bf16(R, S, rays8, z, None, 5, packed, raw5, acts, masks), fl,
                     33 * 128 + 8 * 32 + 24),
        "bwd_chain": (lambda: ops.mlp_bwd_chain_nv_bf16(n, d_pad, d4, packed_bwd, masks, gstore), 2 * 7 * 65536,
                      8 * 32 + 32 * 128 + 20),
        "wgrad": (lambda: ops.mlp_wgrad_nv_bf16(n, 5, acts, gstore, d_pad, d4, gptrs_nv), fl,
                  (32 + 4 * 7 + 2 + 4) * 128 + 20),
    }
for name, (fn, flop, bytes_per_row) in legs.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:10s} {ms:8.3f} ms   {n * flop / ms / 1e9:8.1f} TFLOP/s   {n * bytes_per_row / ms / 1e6:8.1f} GB/s "
          f"(algorithmic)   rows={n}")
