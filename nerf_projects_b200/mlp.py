"""MLP evaluation over ray samples: run_network + NeRF.forward
(ipynb:790-855, nerf.py:57-111) and its backward w.r.t. the parameters.

Two precisions, both hand-written CUDA:
  * ``bf16``  -- one fused tcgen05/TMEM kernel per network pass
                 (``nerfb_mlp_fwd_bf16``), canonical 8x256 architecture only;
  * ``fp32``  -- validation mode: encoding kernel + one SIMT SGEMM per layer,
                 any D/W/skips.  Also provides the parameter gradients (with
                 recompute, chunked over rays) for both modes.
"""
from __future__ import annotations

import contextlib
import os
import threading
from typing import Dict, List, Optional

import torch

from . import _lib as L
from . import ops

_PRECISION = os.environ.get("NERFB200_PRECISION", "bf16")
ROWS_PER_CHUNK = 1 << 18      # rows (ray samples) processed per fp32 chunk


def set_precision(mode: str):
    """'bf16' (tensor cores, default) or 'fp32' (validation mode, 1e-4 parity).  Sets the process-wide default;
    `precision(mode)` overrides it for the current thread inside a `with` block."""
    global _PRECISION
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = mode


_tls = threading.local()


def get_precision() -> str:
    return getattr(_tls, "mode", None) or _PRECISION


@contextlib.contextmanager
def precision(mode: str):
    """Scoped, thread-local precision: `with nb.precision("fp32"): render(...)` leaves the process default untouched."""
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    prev = getattr(_tls, "mode", None)
    _tls.mode = mode
    try:
        yield
    finally:
        _tls.mode = prev


def param_list(net) -> List[torch.Tensor]:
    return list(net.parameters())


def _enc_L(input_ch: int) -> int:
    if input_ch == 3:
        return -1
    assert (input_ch - 3) % 6 == 0, f"input width {input_ch} is not 3+6L"
    return (input_ch - 3) // 6


# ------------------------------------------------------------------ fp32

def _fwd_fp32_chunk(net, R, S, rays, z, pts, viewdirs, keep: bool):
    """Forward of one chunk.  Returns (raw [R*S,ch], saved or None)."""
    dev = net.pts_linears[0].weight.device
    N = R * S
    W, ic, icv = net.W, net.input_ch, net.input_ch_views
    Lp = _enc_L(ic)
    use_vd = bool(net.use_viewdirs)
    Lv = _enc_L(icv) if use_vd else None
    new = lambda c: torch.empty((N, c), dtype=torch.float32, device=dev)

    if (net.D - 1) in net.skips:
        raise L.NerfB200Error("a skip connection after the last pts layer is not supported "
                              "(the reference module cannot run it either)")
    X: List[torch.Tensor] = []            # X[i] = input buffer of pts layer i
    catv = new(W + icv) if use_vd else None
    x0 = new(ic)
    ops.embed_samples(R, S, rays, z, pts, viewdirs, Lp, Lv, x0, 0, catv, W if use_vd else 0)
    cur = x0
    pingpong = [None, None]
    for i, lin in enumerate(net.pts_linears):
        if i in net.skips:
            nxt = new(ic + W)             # [gamma(x) | h_i]  (nerf.py:79-80)
            ops.embed_samples(R, S, rays, z, pts, None, Lp, None, nxt, 0, None, 0)
            out = nxt[:, ic:]
        else:
            if keep:
                nxt = new(W)
            else:
                if pingpong[i % 2] is None:
                    pingpong[i % 2] = new(W)
                nxt = pingpong[i % 2]
            out = nxt
        ops.linear_fwd(cur, lin.weight, lin.bias, out, True)
        X.append(cur)
        cur = nxt
    h_last = cur
    saved = None
    if use_vd:
        raw = new(4)
        ops.linear_fwd(h_last, net.alpha_linear.weight, net.alpha_linear.bias, raw[:, 3:], False)
        ops.linear_fwd(h_last, net.feature_linear.weight, net.feature_linear.bias, catv[:, :W], False)
        hv = new(W // 2)
        ops.linear_fwd(catv, net.views_linears[0].weight, net.views_linears[0].bias, hv, True)
        ops.linear_fwd(hv, net.rgb_linear.weight, net.rgb_linear.bias, raw[:, :3], False)
        if keep:
            saved = dict(X=X, h_last=h_last, catv=catv, hv=hv)
    else:
        raw = new(net.output_ch)
        ops.linear_fwd(h_last, net.output_linear.weight, net.output_linear.bias, raw, False)
        if keep:
            saved = dict(X=X, h_last=h_last)
    return raw, saved


def _bwd_fp32_chunk(net, saved, d_raw, grads: Dict[str, torch.Tensor]):
    """Accumulates d loss / d params of one chunk into ``grads`` (name -> tensor)."""
    dev = d_raw.device
    N = d_raw.shape[0]
    W, ic = net.W, net.input_ch
    new = lambda c: torch.empty((N, c), dtype=torch.float32, device=dev)
    X, h_last = saved["X"], saved["h_last"]
    g = lambda name: grads[name]
    if net.use_viewdirs:
        catv, hv = saved["catv"], saved["hv"]
        ops.linear_bwd_weight(d_raw[:, :3], hv, g("rgb_linear.weight"), g("rgb_linear.bias"))
        d_hv = new(W // 2)
        ops.linear_bwd_data(d_raw[:, :3], net.rgb_linear.weight, d_hv, relu_src=hv)
        ops.linear_bwd_weight(d_hv, catv, g("views_linears.0.weight"), g("views_linears.0.bias"))
        d_feat = new(W)
        ops.linear_bwd_data(d_hv, net.views_linears[0].weight[:, :W], d_feat)
        ops.linear_bwd_weight(d_feat, h_last, g("feature_linear.weight"), g("feature_linear.bias"))
        ops.linear_bwd_weight(d_raw[:, 3:4], h_last, g("alpha_linear.weight"), g("alpha_linear.bias"))
        d_h = new(h_last.shape[1])
        ops.linear_bwd_data(d_feat, net.feature_linear.weight, d_h)
        ops.linear_bwd_data(d_raw[:, 3:4], net.alpha_linear.weight, d_h, relu_src=h_last, accumulate=True)
    else:
        oc = net.output_ch
        ops.linear_bwd_weight(d_raw[:, :oc], h_last, g("output_linear.weight"), g("output_linear.bias"))
        d_h = new(h_last.shape[1])
        ops.linear_bwd_data(d_raw[:, :oc], net.output_linear.weight, d_h, relu_src=h_last)
    if d_h.shape[1] != W:       # skip after the last layer: keep only the h part
        d_h = d_h[:, ic:]
    for i in range(net.D - 1, -1, -1):
        lin = net.pts_linears[i]
        ops.linear_bwd_weight(d_h, X[i], g(f"pts_linears.{i}.weight"), g(f"pts_linears.{i}.bias"))
        if i == 0:
            break
        off = ic if (i - 1) in net.skips else 0
        d_prev = new(W)
        ops.linear_bwd_data(d_h, lin.weight[:, off:], d_prev, relu_src=X[i][:, off:])
        d_h = d_prev


def _slice_rows(t, r0, r1, S=None):
    if t is None:
        return None
    return t[r0:r1] if S is None else t[r0 * S:r1 * S]


def forward_fp32(net, R, S, rays, z, pts=None, viewdirs=None):
    ch = 4 if net.use_viewdirs else net.output_ch
    dev = net.pts_linears[0].weight.device
    step = max(1, ROWS_PER_CHUNK // S)
    if R <= step:
        return _fwd_fp32_chunk(net, R, S, rays, z, pts, viewdirs, keep=False)[0]
    raw = torch.empty((R * S, ch), dtype=torch.float32, device=dev)
    for r0 in range(0, R, step):
        r1 = min(R, r0 + step)
        out, _ = _fwd_fp32_chunk(net, r1 - r0, S, _slice_rows(rays, r0, r1), _slice_rows(z, r0, r1),
                                 _slice_rows(pts, r0, r1, S), _slice_rows(viewdirs, r0, r1), keep=False)
        raw[r0 * S:r1 * S].copy_(out)
    return raw


def backward_fp32(net, R, S, rays, z, pts, viewdirs, d_raw) -> Dict[str, torch.Tensor]:
    """Parameter gradients of sum(raw * d_raw) with forward recompute per chunk."""
    grads = {n: torch.zeros_like(p) for n, p in net.named_parameters()}
    step = max(1, (ROWS_PER_CHUNK // 4) // S)
    for r0 in range(0, R, step):
        r1 = min(R, r0 + step)
        _, saved = _fwd_fp32_chunk(net, r1 - r0, S, _slice_rows(rays, r0, r1), _slice_rows(z, r0, r1),
                                   _slice_rows(pts, r0, r1, S), _slice_rows(viewdirs, r0, r1), keep=True)
        _bwd_fp32_chunk(net, saved, d_raw[r0 * S:r1 * S], grads)
    return grads


# ------------------------------------------------------------------ bf16

_P_INDEX = {}
for _i in range(8):
    _P_INDEX[f"pts_linears.{_i}.weight"] = 2 * _i
    _P_INDEX[f"pts_linears.{_i}.bias"] = 2 * _i + 1
for _k, _name in enumerate(["views_linears.0", "feature_linear", "alpha_linear", "rgb_linear", "output_linear"]):
    _P_INDEX[_name + ".weight"] = 16 + 2 * _k
    _P_INDEX[_name + ".bias"] = 17 + 2 * _k


def packed_weights(net) -> torch.Tensor:
    """bf16 UMMA smem images of the network, re-packed only when a parameter changed."""
    params = dict(net.named_parameters())
    key = tuple((p.data_ptr(), p._version) for p in params.values())
    if net._packed is not None and net._packed_key == key:
        return net._packed
    dev = net.pts_linears[0].weight.device
    nbytes = ops.packed_weights_bytes(net.use_viewdirs)
    if net._packed is None or net._packed.numel() != nbytes or net._packed.device != dev:
        net._packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ptrs = [None] * L.NERFB_NUM_PARAMS
    for name, p in params.items():
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise L.NerfB200Error(f"parameter {name} must be contiguous fp32")
        ptrs[_P_INDEX[name]] = p.data_ptr()
    ops.pack_weights(ptrs, net.use_viewdirs, net.output_ch, net._packed)
    net._packed_key = key
    return net._packed


def forward_bf16(net, R, S, rays, z, pts=None, viewdirs=None):
    if not net.is_canonical():
        raise L.NerfB200Error("the bf16 tensor-core kernel supports the canonical NeRF MLP only "
                              "(D=8, W=256, skips=[4], multires 10/4); use set_precision('fp32')")
    dev = net.pts_linears[0].weight.device
    ch = 4 if net.use_viewdirs else net.output_ch
    raw = torch.empty((R * S, ch), dtype=torch.float32, device=dev)
    ops.mlp_fwd_bf16(R, S, rays, z, pts, viewdirs, net.use_viewdirs, net.output_ch, packed_weights(net), raw)
    return raw


def forward(net, R, S, rays, z, pts=None, viewdirs=None):
    if get_precision() == "bf16":
        return forward_bf16(net, R, S, rays, z, pts, viewdirs)
    return forward_fp32(net, R, S, rays, z, pts, viewdirs)


# ------------------------------------------------- bf16 backward (tensor cores)

TRAIN_ROWS_PER_CHUNK = 1 << 20     # samples per backward chunk: 5.2 GB + 5.0 GB of tile stores
_workspaces = {}


def packed_weights_bwd(net) -> torch.Tensor:
    """Transposed bf16 weight images for the data-gradient chain (repacked when a parameter changed)."""
    params = dict(net.named_parameters())
    key = tuple((p.data_ptr(), p._version) for p in params.values())
    cached = getattr(net, "_packed_bwd", None)
    if cached is not None and getattr(net, "_packed_bwd_key", None) == key:
        return cached
    dev = net.pts_linears[0].weight.device
    nbytes = ops.packed_weights_bwd_bytes()
    if cached is None or cached.numel() != nbytes or cached.device != dev:
        cached = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if net.use_viewdirs:
        ops.pack_weights_bwd(_param_ptrs(params), cached)
    else:
        ops.pack_weights_bwd_nv(_param_ptrs(params), net.output_ch, cached)
    net._packed_bwd, net._packed_bwd_key = cached, key
    return cached


def _param_ptrs(named):
    ptrs = [None] * L.NERFB_NUM_PARAMS
    for name, p in named.items():
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise L.NerfB200Error(f"tensor {name} must be contiguous fp32")
        ptrs[_P_INDEX[name]] = p.data_ptr()
    return ptrs


def _train_workspace(dev, nrows, need_acts=True):
    """(acts, gstore, masks) tile stores for up to nrows samples, cached per device (grown on demand)."""
    need = (ops.train_store_bytes(nrows, 0) if need_acts else 0, ops.train_store_bytes(nrows, 1),
            ops.train_store_bytes(nrows, 2) if need_acts else 0)
    ws = _workspaces.get(dev)
    if ws is None or any(w.numel() < n for w, n in zip(ws, need)):
        sizes = tuple(max(n, ws[i].numel() if ws else 0) for i, n in enumerate(need))
        ws = None
        _workspaces.pop(dev, None)
        ws = tuple(torch.empty(n, dtype=torch.uint8, device=dev) for n in sizes)
        _workspaces[dev] = ws
    return ws


def release_workspaces():
    """Free the cached training tile stores (up to ~10 GB per device; they are re-created on demand)."""
    _workspaces.clear()


def _zero_grads(net) -> Dict[str, torch.Tensor]:
    """Zeroed gradient tensors for every parameter as views of ONE flat buffer (one fill instead of 24; every view
    starts on a 128-byte boundary, which the vector reductions of the weight-gradient kernel need)."""
    named = list(net.named_parameters())
    offs, total = [], 0
    for _, p in named:
        offs.append(total)
        total += (p.numel() + 31) // 32 * 32
    flat = torch.zeros(total, dtype=torch.float32, device=named[0][1].device)
    return {n: flat[o:o + p.numel()].view_as(p) for (n, p), o in zip(named, offs)}


def supports_bf16_backward(net) -> bool:
    """Canonical 8x256 network, with view directions or with the plain output_linear head (4 or 5 channels)."""
    return bool(net.is_canonical() and (net.use_viewdirs or net.output_ch in (4, 5)))


def _raw_ch(net) -> int:
    return 4 if net.use_viewdirs else net.output_ch


def _padded_d_raw(net, d_raw, n, dev):
    """d_raw [n,ch] -> ([npad,4] channels 0..3, [npad] channel 4 or None), zero beyond n (npad: whole 512-row super-tiles)."""
    npad = (n + 511) // 512 * 512
    ch = _raw_ch(net)
    d_raw = L.f32c(d_raw).view(-1, ch)
    if ch == 4 and npad == n and d_raw.data_ptr() % 16 == 0:
        return d_raw, None                 # whole super-tiles already (e.g. 4096 rays x 64 | 192 samples): no copy
    d_pad = torch.zeros((npad, 4), dtype=torch.float32, device=dev)
    d_pad[:n].copy_(d_raw[:, :4])
    d4 = None
    if ch == 5:
        d4 = torch.zeros(npad, dtype=torch.float32, device=dev)
        d4[:n].copy_(d_raw[:, 4])
    return d_pad, d4


def _fwd_save(net, R, S, rays, z, pts, viewdirs, packed, raw, acts, masks):
    if net.use_viewdirs:
        ops.mlp_fwd_save_bf16(R, S, rays, z, pts, viewdirs, packed, raw, acts, masks)
    else:
        ops.mlp_fwd_save_nv_bf16(R, S, rays, z, pts, net.output_ch, packed, raw, acts, masks)


# Chain + weight gradients in ONE cooperative launch (nerf_bwd_fused_kernel: dY handed over through L2 with per-tile
# completion flags).  Correct (tests/test_gpu_tc.py::test_fused_backward_launch_equals_separate_kernels) but MEASURED SLOWER
# than the two launches: 3.27 ms against 0.905 + 1.543 ms on 786 432 samples at the best SM split (profiles/README.md) --
# the weight-gradient CTAs are bound by their own per-stage rate (~1.5 us per 64-sample stage), not by HBM, so giving half
# of the SMs to the chain doubles their time.  Kept as an opt-in experiment: NERFB200_FUSED_BWD=1.
FUSED_BWD = os.environ.get("NERFB200_FUSED_BWD", "0") == "1"
FUSED_BWD_CHAIN_SMS = int(os.environ.get("NERFB200_FUSED_BWD_CHAIN_SMS", "68"))
FUSED_BWD_MIN_ROWS = 1 << 17       # below ~2 chain iterations the single launch has nothing to overlap


def _chain_and_wgrad(net, n, d_pad, d4, packed_bwd, acts, masks, gstore, gptrs):
    if net.use_viewdirs:
        if FUSED_BWD and n >= FUSED_BWD_MIN_ROWS:
            ops.mlp_bwd_fused_bf16(n, d_pad, packed_bwd, masks, acts, gstore, gptrs, FUSED_BWD_CHAIN_SMS)
            return
        ops.mlp_bwd_chain_bf16(n, d_pad, packed_bwd, masks, gstore)
        ops.mlp_wgrad_bf16(n, acts, gstore, d_pad, gptrs)
    else:
        ops.mlp_bwd_chain_nv_bf16(n, d_pad, d4, packed_bwd, masks, gstore)
        ops.mlp_wgrad_nv_bf16(n, net.output_ch, acts, gstore, d_pad, d4, gptrs)


def backward_bf16(net, R, S, rays, z, pts, viewdirs, d_raw) -> Dict[str, torch.Tensor]:
    """Parameter gradients of sum(raw * d_raw) on the tensor cores: per ray chunk the fused forward
    is re-run with its layer inputs stored (recompute), then the data-gradient chain, then the
    weight-gradient GEMMs.  Nothing is kept between forward and backward except rays, z and d_raw."""
    if not supports_bf16_backward(net):
        raise L.NerfB200Error("the tensor-core backward supports the canonical 8x256 network only")
    dev = net.pts_linears[0].weight.device
    named = dict(net.named_parameters())
    grads = _zero_grads(net)
    gptrs = _param_ptrs(grads)
    packed, packed_bwd = packed_weights(net), packed_weights_bwd(net)
    nchunk = max(1, -(-(R * S) // TRAIN_ROWS_PER_CHUNK))       # even split: no tiny tail chunk
    step = -(-R // nchunk)
    acts, gstore, masks = _train_workspace(dev, min(R, step) * S)
    ch = _raw_ch(net)
    d_raw = L.f32c(d_raw).view(-1, ch)
    for r0 in range(0, R, step):
        r1 = min(R, r0 + step)
        n = (r1 - r0) * S
        d_pad, d4 = _padded_d_raw(net, d_raw[r0 * S:r1 * S], n, dev)
        raw_tmp = torch.empty((n, ch), dtype=torch.float32, device=dev)
        _fwd_save(net, r1 - r0, S, _slice_rows(rays, r0, r1), _slice_rows(z, r0, r1), _slice_rows(pts, r0, r1, S),
                  _slice_rows(viewdirs, r0, r1), packed, raw_tmp, acts, masks)
        _chain_and_wgrad(net, n, d_pad, d4, packed_bwd, acts, masks, gstore, gptrs)
    return grads


def _save_budget_bytes(dev=None) -> int:
    """Bytes the training forward may keep for the backward: NERFB200_SAVE_BUDGET_GB if set, else the smaller of 24 GB and
    a quarter of the device's memory (a 4096-ray step needs 10 GB; larger steps recompute chunk by chunk)."""
    env = os.environ.get("NERFB200_SAVE_BUDGET_GB")
    if env is not None:
        return int(float(env) * (1 << 30))
    total = torch.cuda.get_device_properties(dev if dev is not None else torch.cuda.current_device()).total_memory
    return min(24 << 30, total // 4)




def can_save_forward(net, nrows: int) -> bool:
    """Training forward may keep its layer inputs (5 KB per sample) for the backward instead of
    recomputing them, when the tile stores of this pass fit the budget."""
    return (get_precision() == "bf16" and supports_bf16_backward(net)
            and ops.train_store_bytes(nrows, 0) + ops.train_store_bytes(nrows, 1)
            <= _save_budget_bytes(net.pts_linears[0].weight.device))


def forward_save_bf16(net, R, S, rays, z, pts=None, viewdirs=None):
    """Fused forward that also stores every layer input (bf16 operand images) and the ReLU sign bits
    -> (raw, (acts, masks))."""
    dev = net.pts_linears[0].weight.device
    n = R * S
    raw = torch.empty((n, _raw_ch(net)), dtype=torch.float32, device=dev)
    acts = torch.empty(ops.train_store_bytes(n, 0), dtype=torch.uint8, device=dev)
    masks = torch.empty(ops.train_store_bytes(n, 2), dtype=torch.uint8, device=dev)
    _fwd_save(net, R, S, rays, z, pts, viewdirs, packed_weights(net), raw, acts, masks)
    return raw, (acts, masks)


def backward_from_saved(net, nrows: int, saved, d_raw) -> Dict[str, torch.Tensor]:
    """Parameter gradients from what a forward_save_bf16 pass stored (no recompute)."""
    acts, masks = saved
    dev = acts.device
    grads = _zero_grads(net)
    d_pad, d4 = _padded_d_raw(net, d_raw, nrows, dev)
    _, gstore, _ = _train_workspace(dev, nrows, need_acts=False)
    _chain_and_wgrad(net, nrows, d_pad, d4, packed_weights_bwd(net), acts, masks, gstore, _param_ptrs(grads))
    return grads


def backward(net, R, S, rays, z, pts, viewdirs, d_raw):
    """bf16 mode: tensor-core backward for the canonical 8x256 network (either head), fp32 CUDA-core kernels
    otherwise (other architectures, validation mode)."""
    if get_precision() == "bf16" and supports_bf16_backward(net):
        return backward_bf16(net, R, S, rays, z, pts, viewdirs, d_raw)
    return backward_fp32(net, R, S, rays, z, pts, viewdirs, d_raw)
