// Fused NeRF MLP on tcgen05 tensor cores (sm_100a): forward, forward-with-save (training) and the
// data-gradient chain of the backward, one templated persistent kernel.
//
// Replaces run_network + NeRF.forward (ipynb:790-855, nerf.py:57-111) for all
// R*S ray samples of a chunk: sample positions o + d*z, the sin/cos positional
// encoding (in registers), the 8x256 skip-connected MLP, the density head, the
// feature / view-dependent layers and the rgb head run inside ONE persistent
// kernel; activations live in shared memory (bf16) and TMEM (fp32
// accumulators) and never touch HBM.  Output: raw[R*S, 4] = [rgb logits, sigma].
// MODE_FWD_SAVE additionally bulk-stores every layer input for the weight-gradient
// kernel (wgrad_tc.cu); MODE_BWD runs d_raw -> d_hv -> d_feature -> dY_7..dY_0 through
// the transposed weights with the same machinery.
//
// Design (DESIGN.md section 4 has the full derivation and the measurements):
//   * one 2-CTA cluster per SM pair, tcgen05.mma.cta_group::2 (UMMA M=256,
//     N=256, K=16): each CTA owns 128 rows (samples) of A and HALF of the weight
//     rows of B, so a whole 256x256 layer is 64 KB per CTA and stays resident
//     in a 4-slot ring while BOTH row tiles (A,B) of the CTA pass over it:
//     every weight byte read from L2 feeds 512 samples;
//   * two row tiles per CTA ping-pong: while the tensor core runs layer l of
//     tile B, the 8 epilogue warps drain tile A's accumulator from TMEM
//     (bias, ReLU, bf16, 128B-swizzled store = next layer's A operand);
//   * weights are pre-packed (nerfb_pack_weights) as the exact shared-memory
//     images (K-major, SWIZZLE_128B) and streamed with tensor-map TMA copies
//     (cp.async.bulk.tensor.2d, .cta_group::2) whose bytes complete on the
//     LEADER CTA's mbarrier from both CTAs -- no relay between the CTAs;
//   * the MMA issuer is one elected lane running a static schedule of BLOCKS
//     (one tile's pass over one layer) with minimal work between MMA groups;
//   * the density head is evaluated in fp32 from the un-rounded layer-7
//     accumulators, the rgb head in fp32 from the fp32 view
//     layer activations;
//   * every barrier wait is bounded by a watchdog (tc05.cuh).
#include <cstdlib>
#include <mutex>
#include <cuda.h>
#include "common.cuh"
#include "tc05.cuh"
#include "mlp_layout.cuh"
#include "composite.cuh"
#include "wgrad_body.cuh"

namespace nerfb {
using namespace tc05;

tc05::Watch get_watch();  // umma_selftest.cu

// ------------------------------------------------------------- schedule
// A static program shared by the three asynchronous roles.  Generated on the
// host (build_schedule) and passed by value.
// bytes: bytes per CTA of the chunk (8192 or 16384), plus -- when the load also brings the layer's bias tile -- bit 0 set,
// bit 1 = bias buffer (0/1) and bits 2..5 = index of the tile in the blob's bias area
struct LoadEnt { uint32_t off; uint32_t bytes; };          // blob offset of the chunk, bytes per CTA (+ bias tile bits)
enum : uint8_t { F_FIRST = 1, F_LAST = 2, F_REL_SLOT = 4, F_REL_ENC = 8, F_N128 = 16,
                 F_BIAS0 = 32, F_BIAS1 = 64 };   // block starts with the layer's bias MMA, tile in bias buffer 0 / 1
// One op = 4 MMAs (K = 64) of one tile over one weight chunk.  Consecutive ops of the same tile and layer form a
// BLOCK (one pass of the tile over the layer): `tile` carries the block length in its high nibble on the block's
// first op; the MMA issuer handles a block as one unit with minimal work between its chunks.
struct OpEnt { uint8_t load, tile, asrc, flags; };          // asrc 0..3 = act chunk, 4 = enc
struct Sched {
  int nloads, nops, nlayers;
  LoadEnt loads[44];
  OpEnt ops[84];
};

static void build_schedule(bool vd, Sched& s) {
  s.nloads = s.nops = 0;
  s.nlayers = vd ? 10 : 8;
  auto chunk_off = [](int id) -> uint32_t {
    return id < kNumFullChunks ? (uint32_t)id * kChunkFull
                               : (uint32_t)kNumFullChunks * kChunkFull + (uint32_t)(id - kNumFullChunks) * kChunkViews;
  };
  auto add_load = [&](int id, bool n128) -> uint8_t {
    s.loads[s.nloads] = LoadEnt{chunk_off(id), (uint32_t)(n128 ? kChunkBytes / 2 : kChunkBytes)};
    return (uint8_t)s.nloads++;
  };
  auto add_op = [&](uint8_t load, int tile, int asrc, int flags) {
    s.ops[s.nops++] = OpEnt{load, (uint8_t)tile, (uint8_t)asrc, (uint8_t)flags};
  };
  int blk_first[48], blk_count[48], nblk = 0;
  auto block = [&](int first, int count) { blk_first[nblk] = first; blk_count[nblk++] = count; };
  int nbias = 0;
  // Every layer's bias tile (mlp_layout.cuh) arrives with the layer's FIRST load (same barrier) into bias buffer
  // nbias % 2, and each tile's block starts with the bias MMA.  (Reuse of a buffer two layers later is safe: that load is
  // issued after the loader waited for a ring slot released by an op of the layer in between, i.e. after every MMA of
  // this layer; tests/test_schedule_cpu.py checks the distances.)
  auto bias_on = [&](uint8_t first_load) -> int {
    const int buf = nbias & 1;
    s.loads[first_load].bytes |= 1u | ((uint32_t)buf << 1) | ((uint32_t)nbias << 2);
    ++nbias;
    return buf ? F_BIAS1 : F_BIAS0;
  };
  // plain K=256 layer: chunks id0..id0+3, both tiles pass over the resident layer
  auto plain = [&](int id0, bool n128) {
    uint8_t l[4];
    for (int k = 0; k < 4; ++k) l[k] = add_load(id0 + k, n128);
    const int fb = bias_on(l[0]);
    const int f = (n128 ? F_N128 : 0);
    block(s.nops, 4);
    for (int k = 0; k < 4; ++k) add_op(l[k], 0, k, f | (k == 0 ? (F_FIRST | fb) : 0) | (k == 3 ? F_LAST : 0));
    block(s.nops, 4);
    for (int k = 0; k < 4; ++k) add_op(l[k], 1, k, f | F_REL_SLOT | (k == 0 ? (F_FIRST | fb) : 0) | (k == 3 ? F_LAST : 0));
  };
  // layer with an extra K=64 chunk read from the enc buffer (skip concat / view concat):
  // tile A consumes it first, tile B last, so the 4-slot ring never needs 5 live chunks.
  auto with_enc = [&](int id_enc, int id0, bool n128) {
    const int f = n128 ? F_N128 : 0;
    const uint8_t g = add_load(id_enc, n128);
    const int fb = bias_on(g);
    uint8_t l[4];
    for (int k = 0; k < 4; ++k) l[k] = add_load(id0 + k, n128);
    const uint8_t g2 = add_load(id_enc, n128);
    block(s.nops, 5);
    add_op(g, 0, 4, f | F_FIRST | fb | F_REL_SLOT | F_REL_ENC);
    for (int k = 0; k < 4; ++k) add_op(l[k], 0, k, f | (k == 3 ? F_LAST : 0));
    block(s.nops, 5);
    for (int k = 0; k < 4; ++k) add_op(l[k], 1, k, f | F_REL_SLOT | (k == 0 ? (F_FIRST | fb) : 0));
    add_op(g2, 1, 4, f | F_LAST | F_REL_SLOT | F_REL_ENC);
  };
  {  // layer 0: K = 64 (gamma(x), written by the epilogue warps into chunk 0 of the tile's own activation buffer),
     // one weight chunk used by both tiles
    const uint8_t g0 = add_load(0, false);
    const int fb = bias_on(g0);
    block(s.nops, 1);
    add_op(g0, 0, 0, F_FIRST | F_LAST | fb);
    block(s.nops, 1);
    add_op(g0, 1, 0, F_FIRST | F_LAST | F_REL_SLOT | fb);
  }
  for (int l = 1; l <= 4; ++l) plain(1 + 4 * (l - 1), false);
  with_enc(17, 18, false);          // layer 5: [gamma(x) | h4]
  plain(22, false);                 // layer 6
  plain(26, false);                 // layer 7
  if (vd) {
    plain(30, false);               // feature_linear
    with_enc(34, 35, true);         // views_linears[0]: [feature | gamma(d)], N = 128
  }
  for (int i = 0; i < nblk; ++i) s.ops[blk_first[i]].tile |= (uint8_t)(blk_count[i] << 4);
}

// backward chain: d_hv -> d_feature -> d_h7 -> ... -> d_h0 with transposed weights
// (vd == false: output_linear's data gradient is the first-stage operand, so the chain is W7^T .. W1^T only)
static void build_schedule_bwd(bool vd, Sched& s) {
  s.nloads = s.nops = 0;
  s.nlayers = vd ? 9 : 7;
  auto add_load = [&](int id) -> uint8_t {
    s.loads[s.nloads] = LoadEnt{(uint32_t)id * kChunkFull, (uint32_t)kChunkBytes};
    return (uint8_t)s.nloads++;
  };
  auto add_op = [&](uint8_t load, int tile, int asrc, int flags) {
    s.ops[s.nops++] = OpEnt{load, (uint8_t)tile, (uint8_t)asrc, (uint8_t)flags};
  };
  int blk_first[48], blk_count[48], nblk = 0;
  auto block = [&](int first, int count) { blk_first[nblk] = first; blk_count[nblk++] = count; };
  if (vd) {  // views^T: K = 128 (two chunks)
    const uint8_t l0 = add_load(0), l1 = add_load(1);
    block(s.nops, 2);
    add_op(l0, 0, 0, F_FIRST); add_op(l1, 0, 1, F_LAST);
    block(s.nops, 2);
    add_op(l0, 1, 0, F_FIRST | F_REL_SLOT); add_op(l1, 1, 1, F_LAST | F_REL_SLOT);
  }
  for (int j = vd ? 1 : 2; j <= 8; ++j) {      // j = 1: feature^T, j = 2..8: W7^T .. W1^T
    uint8_t l[4];
    for (int k = 0; k < 4; ++k) l[k] = add_load(2 + 4 * (j - 1) + k);
    block(s.nops, 4);
    for (int k = 0; k < 4; ++k) add_op(l[k], 0, k, (k == 0 ? F_FIRST : 0) | (k == 3 ? F_LAST : 0));
    block(s.nops, 4);
    for (int k = 0; k < 4; ++k) add_op(l[k], 1, k, F_REL_SLOT | (k == 0 ? F_FIRST : 0) | (k == 3 ? F_LAST : 0));
  }
  for (int i = 0; i < nblk; ++i) s.ops[blk_first[i]].tile |= (uint8_t)(blk_count[i] << 4);
}

struct MlpArgs {
  const float* rays; int rw;
  const float* z; const float* pts; const float* viewdirs;
  int S; long long nrows; int use_viewdirs; int out_ch;
  const uint8_t* packed; float* raw;
  int nst, nclusters;
  Watch watch;
  long long* trace;   // debug timeline (nerfb_set_trace), 3 regions of kTraceLen entries
  uint8_t* acts;      // MODE_FWD_SAVE (out): [ntiles][kFwdTileBytes]
  uint8_t* masks;     // MODE_FWD_SAVE (out) / MODE_BWD (in): ReLU sign bits, [ntiles][kMaskTileBytes]
  uint8_t* gstore;    // MODE_BWD: [ntiles][kBwdTileBytes]
  const float* d_raw; // MODE_BWD: [nrows,4]
  const float* d_raw4; // MODE_BWD without view directions: gradient of output channel 4 [nrows] or null
  int param_off;      // byte offset of the fp32 parameter block inside `packed`
  int bias_off;       // byte offset of the bias-tile area inside `packed` (forward modes)
  int* flags;         // MODE_BWD inside the fused backward launch: completion flags [ntiles][kFlagGroups][2] for the
  int epoch;          //   weight-gradient CTAs (wgrad_body.cuh), set to `epoch` when a half's bulk stores have completed
  // MODE_FWD_COMP: compositing folded into the forward
  int lindisp;        // z == nullptr: coarse depths computed in-kernel from the rays' near / far (z_at)
  float zstep;        // fl(1 / (S - 1)): torch.linspace's step for those depths
  float* comp_rec;    // [R][S/32][kRecFloats] segment records
  float* comp_wpart;  // [R*S] alpha * (exclusive product inside the segment), or null (final pass)
  // row index -> ray index without a 64-bit division per row (filled by launch_mlp): for nrows < 2^32 and S >= 2,
  // row / S = (t + ((row - t) >> 1)) >> div_shift with t = umulhi(div_mul, row) (Granlund-Montgomery, exact); else div_mul = 0
  uint32_t div_mul; int div_shift;
};

// ray index of a row (row < nrows)
__device__ __forceinline__ long long row_to_ray(const MlpArgs& a, long long grow) {
  if (a.div_mul != 0u) {
    const uint32_t n = (uint32_t)grow;
    const uint32_t t = __umulhi(a.div_mul, n);
    return (long long)((t + ((n - t) >> 1)) >> a.div_shift);
  }
  return a.S == 1 ? grow : grow / a.S;
}
constexpr int kTraceLen = 4096;

// The fp32 parameter block of the network a launch evaluates, ALSO in the constant bank: the final epilogue
// (views bias, rgb head) reads it with compile-time offsets, so the values are constant-bank operands of the FMAs
// instead of 64 broadcast LDS.128 per thread (that epilogue was bound by the shared-memory pipe).  One block per
// process; launch_mlp copies into it in stream order and chains launches on different streams with an event.
__constant__ float c_par[kParamFloats + 4];
// Training tile stores are bulk (TMA) copies out of the activation buffers.  Writing them with 16-byte stores from the
// epilogue registers was measured and rejected (forward+save 1.69 ms vs 1.02 ms, chain 2.09 vs 1.66 ms; profiles/README.md).
static long long* g_trace = nullptr;
static cudaEvent_t g_const_event = nullptr;   // recorded after every fused-MLP launch (guards c_par)

// ------------------------------------------------------------- encoding
// value of column c of gamma(p) (embedder.py:54-65): [p | sin f0 p | cos f0 p | sin f1 p | ...]
// A thread produces 32 consecutive columns (HH = 0: columns 0..31, octaves 0..4; HH = 1: columns 32..63, octaves
// 4..9).  Per dimension ONE accurate sin/cos at the thread's first octave -- p*2^k0 reduced to revolutions with a
// two-term 1/2pi, MUFU on the reduced argument (~1e-7 even at 2^9*|p| ~ 3000 rad) -- and the higher octaves by the
// double-angle recurrence (s' = 2sc, c' = 1 - 2s^2): at most 5 doublings, error <= ~3e-6, far below the bf16
// rounding (4e-3) of the operand it feeds.  (The fp32 validation mode uses embed.cu, not this.)
template <int L, int HH>
__device__ __forceinline__ void encode_cols(const float (&p)[3], uint32_t (&packed)[16]) {
  constexpr float kInv2PiHi = 0.15915494f, kInv2PiLo = 6.4206383e-09f, k2Pi = 6.2831855f;
  constexpr int c_lo = 32 * HH, c_hi = 32 * HH + 31;
  constexpr int k0 = c_lo < 3 ? 0 : (c_lo - 3) / 6;                        // first octave this thread touches
  constexpr int k1 = (c_hi < 3 + 6 * L ? (c_hi - 3) / 6 : L - 1);          // last one
  float sn[3][10], cs[3][10];
  if (k0 <= k1) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float sc = (float)(1 << k0);
      const float thi = p[j] * kInv2PiHi;
      const float tlo = fmaf(p[j], kInv2PiHi, -thi) + p[j] * kInv2PiLo;
      const float a = thi * sc;
      const float f = ((a - rintf(a)) + tlo * sc) * k2Pi;
      sn[j][k0] = __sinf(f);
      cs[j][k0] = __cosf(f);
#pragma unroll
      for (int k = k0 + 1; k <= k1; ++k) {
        sn[j][k] = 2.f * sn[j][k - 1] * cs[j][k - 1];
        cs[j][k] = fmaf(-2.f * sn[j][k - 1], sn[j][k - 1], 1.f);
      }
    }
  }
  float e[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int c = 32 * HH + i;
    if (c < 3) {
      e[i] = p[c];
    } else if (c < 3 + 6 * L) {
      const int k = (c - 3) / 6, r = (c - 3) % 6, j = r % 3;
      e[i] = (r < 3) ? sn[j][k] : cs[j][k];
    } else {
      e[i] = 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
}

__device__ __forceinline__ void store_row32(uint8_t* chunk, int row, int unit0, const uint32_t (&w)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(chunk + sw128_offset(row, unit0 + i)) =
        make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

// rgb head of one sample from its 64 view-layer columns [64*HH, 64*HH+64): x = relu(acc + bias), partial
// rgb[k] += x * W_rgb[k][col]; all weights are constant-bank operands (c_par, compile-time offsets).
template <int HH, bool SAVE>
__device__ __forceinline__ void rgb_head_half(const uint32_t (&v)[32], int b, float (&ca)[3], float (&cb)[3],
                                              float (&xs)[32], uint32_t& signbits) {
  uint32_t m = 0u;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int c = 64 * HH + 32 * b + j;          // b is a compile-time constant at every call site (unrolled)
    const float pre = __uint_as_float(v[j]);      // bias already in the accumulator (bias MMA)
    (void)c;
    if (SAVE) m = __funnelshift_l(__float_as_uint(pre), m, 1);
    const float x = fmaxf(pre, 0.f);
    xs[j] = x;
    if (j & 1) {
      cb[0] = fmaf(x, c_par[kOffRgbW + c], cb[0]);
      cb[1] = fmaf(x, c_par[kOffRgbW + 128 + c], cb[1]);
      cb[2] = fmaf(x, c_par[kOffRgbW + 256 + c], cb[2]);
    } else {
      ca[0] = fmaf(x, c_par[kOffRgbW + c], ca[0]);
      ca[1] = fmaf(x, c_par[kOffRgbW + 128 + c], ca[1]);
      ca[2] = fmaf(x, c_par[kOffRgbW + 256 + c], ca[2]);
    }
  }
  signbits = m;
}

// ------------------------------------------------- compositing (MODE_FWD_COMP)
// raw2outputs (ipynb:254-349) folded into the forward: a warp composites the 32 consecutive samples of ONE ray (S is a
// multiple of 32) whose [rgb logits, sigma] the final epilogue has left in shared memory.  Inside the segment: alpha, the
// exclusive product P_i of (1 - alpha + 1e-10), w'_i = alpha_i * P_i, and the sums of w', w' rgb, w' z; the segment's
// record {Q = product over the segment, sums} goes to global memory (32 bytes per 32 samples).  The transmittance in
// front of a segment, T = product of the earlier segments' Q, multiplies the sums when the ray is finished:
// nerfb_composite_finish (final pass) or nerfb_resample_records (coarse pass, which also needs the weights T * w').
//
// WHO does it: two otherwise idle warps per CTA (warp 3: tile 0, warp 2: tile 1), fed through a 4 KB stage in shared
// memory that the final epilogue fills, and started only inside the next super-tile's middle layers.  The fused kernel
// is bound by the instruction issue of its eight epilogue warps (ncu source page: they wait for the tensor pipe in only
// ~7 % of their samples), so every instruction added on an SM scheduler costs throughput -- measured A/B
// (profiles/README.md): one compositor warp -2.0 % of the kernel (stubbing its body out gave it all back; an unrolled
// IEEE / fp64 body and a rolled fp32 one cost the same), the same work inside the epilogue warps -3.5 %, spread over
// three of their layers -4 %; a last-arriver finish inside the kernel (arrival counters, __threadfence) ~4 k cycles per
// segment.  Hence: a small rolled fp32 loop with the fast exp (MUFU) forms, on two schedulers.  This path exists in
// bf16 mode only, whose tolerance (2e-2) is five orders of magnitude above these rounding differences; the fp64-scan
// kernels (composite.cu) remain the validation-mode / training path.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ float z_at_step(int i, int S, float step, float near_v, float far_v, int lindisp) {
  // z_at (common.cuh) with torch.linspace's step = fl(1/(S-1)) precomputed on the host: bit-identical depths
  const float t = (S <= 1) ? 0.f : (i < S / 2 ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(S - 1 - i), 1.0f));
  const float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near_v, omt), __fmul_rn(far_v, t));
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near_v), omt), __fmul_rn(__fdiv_rn(1.0f, far_v), t)));
}

// the four 32-sample segments of one 128-row tile; g0 = global index of the tile's first sample
__device__ __noinline__ void composite_tile(const MlpArgs& a, const float4* __restrict__ stage, long long g0, int lane) {
  const int S = a.S;
  if (g0 >= a.nrows) return;                  // warp-uniform (nrows = R*S is a multiple of 32)
  uint32_t r = (uint32_t)g0 / (uint32_t)S;    // COMP launches are limited to < 2^31 samples
  int i0 = (int)((uint32_t)g0 - r * (uint32_t)S);
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    const long long grow0 = g0 + 32 * q;
    if (grow0 >= a.nrows) return;
    const float4 v = stage[32 * q + lane];
    const float* rp = a.rays + (size_t)r * a.rw;
    const bool last_seg = (i0 + 32 == S);
    float zi, zl = 0.f;
    if (a.z != nullptr) {
      zi = __ldg(a.z + grow0 + lane);
      if (!last_seg) zl = __ldg(a.z + grow0 + 32);                   // first depth of the ray's next segment
    } else {
      const float near_v = __ldg(rp + 6), far_v = __ldg(rp + 7);
      zi = z_at_step(i0 + lane, S, a.zstep, near_v, far_v, a.lindisp);
      if (!last_seg) zl = z_at_step(i0 + 32, S, a.zstep, near_v, far_v, a.lindisp);
    }
    const float nrm = ray_norm(rp + 3);
    float zn = __shfl_down_sync(0xffffffffu, zi, 1);
    if (lane == 31) zn = zl;
    const float dist = ((lane == 31 && last_seg) ? 1e10f : (zn - zi)) * nrm;
    const float alpha = 1.0f - __expf(-fmaxf(v.w, 0.f) * dist);
    float p = (1.0f - alpha) + 1e-10f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float pq = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p *= pq;
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0f;
    const float Q = __shfl_sync(0xffffffffu, p, 31);
    const float pw = alpha * excl;
    if (a.comp_wpart != nullptr) a.comp_wpart[grow0 + lane] = pw;
    float A = pw, C0 = pw * fast_sigmoid(v.x), C1 = pw * fast_sigmoid(v.y), C2 = pw * fast_sigmoid(v.z), D = pw * zi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      A += __shfl_xor_sync(0xffffffffu, A, o);
      C0 += __shfl_xor_sync(0xffffffffu, C0, o);
      C1 += __shfl_xor_sync(0xffffffffu, C1, o);
      C2 += __shfl_xor_sync(0xffffffffu, C2, o);
      D += __shfl_xor_sync(0xffffffffu, D, o);
    }
    if (lane == 0) {
      float4* rq = reinterpret_cast<float4*>(a.comp_rec + ((size_t)r * (S >> 5) + (i0 >> 5)) * kRecFloats);
      rq[0] = make_float4(Q, 0.f, A, C0);
      rq[1] = make_float4(C1, C2, D, 0.f);
    }
    i0 += 32;
    if (i0 == S) { i0 = 0; ++r; }
  }
}

// -------------------------------------------------------------- kernel
template <int MODE>
__device__ __forceinline__ void mlp_body(const MlpArgs& a, const Sched& sc, const CUtensorMap& tm128, const CUtensorMap& tm64,
                                         const CUtensorMap& tm32) {
  constexpr bool SAVE = (MODE == MODE_FWD_SAVE || MODE == MODE_FWD_SAVE_NV);
  constexpr bool BWD = (MODE == MODE_BWD || MODE == MODE_BWD_NV);
  constexpr bool NV = (MODE == MODE_FWD_SAVE_NV || MODE == MODE_BWD_NV);   // training kernels of the no-view-directions network
  constexpr bool COMP = (MODE == MODE_FWD_COMP || MODE == MODE_FWD_COMP_NV);   // forward + compositing (inference)
  constexpr bool kVD = !(NV || MODE == MODE_FWD_NV || MODE == MODE_FWD_COMP_NV);   // view-dependent head (compile time)
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const uint32_t sbase = smem_u32(smem);

  float* s_par = reinterpret_cast<float*>(smem + kSmParams);
  float* s_sig = reinterpret_cast<float*>(smem + kSmSig);
  auto bar = [&](int i) -> uint32_t { return sbase + kSmBars + 8 * i; };
  const int B_FULL = 0, B_EMPTY = kNS, B_ACT = 2 * kNS, B_DONE = 2 * kNS + 2, B_ENC = 2 * kNS + 4;
  const int B_STFULL = 2 * kNS + 5, B_STFREE = 2 * kNS + 7;      // MODE_FWD_COMP: [rgb, sigma] stage of tile t published / consumed

  if ((sbase & 1023u) != 0) watchdog_fire(a.watch, 0xA11, sbase, 0);
  // debug timeline: blocks 0 (leader) and 1 (peer) of cluster 0; 4 regions each: loader, MMA, epilogue, relay/observer
  // in-kernel clock64 timeline (tools/trace_mlp.py): compiled in only with -DNERFB_TRACE=1 (variant build) -- as a run-time
  // flag its predicated hooks were 0.8-1.1 % of the executed instructions of every launch
#ifdef NERFB_TRACE
  const bool tracing = (a.trace != nullptr) && blockIdx.x < 2;
#else
  constexpr bool tracing = false;
#endif
  const int tr_base = blockIdx.x * 4;
  int tr_n = 0;
  auto tr = [&](int region, int tag) {
    if (tracing && tr_n < kTraceLen) a.trace[(tr_base + region) * kTraceLen + tr_n++] = ((long long)tag << 44) | (clock64() & 0xFFFFFFFFFFFll);
  };

  // ---- one-time setup
  if (tid == 0) {
    for (int s = 0; s < kNS; ++s) {
      // ONE barrier per slot, in the leader CTA: its loader arrives with expect_tx for BOTH halves and the
      // tensor-map copies of both CTAs complete their bytes on it (.cta_group::2) -- no relay hop
      mbar_init(bar(B_FULL + s), 1);
      mbar_init(bar(B_EMPTY + s), 1);
    }
    mbar_init(bar(B_ACT + 0), 2);
    mbar_init(bar(B_ACT + 1), 2);
    mbar_init(bar(B_DONE + 0), 1);
    mbar_init(bar(B_DONE + 1), 1);
    mbar_init(bar(B_ENC), 1);
    mbar_init(bar(B_STFULL + 0), 1);
    mbar_init(bar(B_STFULL + 1), 1);
    mbar_init(bar(B_STFREE + 0), 1);
    mbar_init(bar(B_STFREE + 1), 1);
    fence_barrier_init();
  }
  // The schedule tables are read from SHARED memory in the hot loops: a dynamically indexed read of the
  // kernel-parameter constant bank misses the (tiny, instruction-shared) constant cache almost every time here,
  // ~500 cycles per table entry in the loader and the MMA issuer (timeline in profiles/).
  for (int i = tid; i < sc.nloads; i += kThreads)
    reinterpret_cast<uint2*>(smem + kSmLoads)[i] = make_uint2(sc.loads[i].off, sc.loads[i].bytes);
  for (int i = tid; i < sc.nops; i += kThreads) {
    const OpEnt e = sc.ops[i];
    reinterpret_cast<uint32_t*>(smem + kSmOps)[i] =
        (uint32_t)e.load | ((uint32_t)e.tile << 8) | ((uint32_t)e.asrc << 16) | ((uint32_t)e.flags << 24);
  }
  {
    // (the first 9.5 KB of the block's shared-memory image -- the biases, now added by MMAs -- hold the two bias-tile
    //  buffers and the constant A operand of the bias MMAs instead)
    const float* gpar = reinterpret_cast<const float*>(a.packed + a.param_off);
    for (int i = kOffHeadW + tid; i < kParamFloats; i += kThreads) s_par[i] = __ldg(gpar + i);
  }
  if (tid < 16) {     // ones tile: [8 rows x 16] bf16, columns 0 and 1 = 1.0 (K group 0 = bytes 0..127, K group 1 = 128..255)
    *reinterpret_cast<uint4*>(smem + kSmOnes + 16 * tid) = make_uint4(tid < 8 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    tmem_alloc<2>(sbase + kSmTmemSlot, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kSmTmemSlot);

  const int niter = (a.nst - cluster_id + a.nclusters - 1) / a.nclusters;   // super-tiles of this cluster
  if (lane == 0 && (warp == 0 || warp == 1 || warp == 4)) tr(warp == 0 ? 0 : (warp == 1 ? 1 : 2), 0xFFF);  // clock alignment marker

  if (warp == 0) {
    // ===================== weight loader (both CTAs: own half of every chunk)
    if (lane == 0) {
      long long gl = 0;
      const uint32_t full_leader = mapa(bar(B_FULL), 0);      // shared::cluster address of the leader's w_full[0]
      for (int it = 0; it < niter; ++it) {
        for (int q = 0; q < sc.nloads; ++q, ++gl) {
          const int slot = (int)(gl % kNS);
          if (gl >= kNS) mbar_wait(bar(B_EMPTY + slot), (uint32_t)(((gl / kNS) + 1) & 1), a.watch, 0x100 + q);
          tr(0, 0x100 + q);
          const uint2 le2 = reinterpret_cast<const uint2*>(smem + kSmLoads)[q];
          const LoadEnt le{le2.x, le2.y & ~63u};
          const bool with_bias = (le2.y & 1u) != 0;
          if (rank == 0) mbar_arrive_expect_tx(bar(B_FULL + slot), 2 * le.bytes + (with_bias ? 2 * kBiasTileHalf : 0));
          // my half of the chunk = rows [row0, row0 + bytes/128) of the blob viewed as [rows][128 B]
          const int row0 = (int)((le.off + rank * le.bytes) >> 7);
          tma_load_2d_cg2(sbase + kSmRing + slot * kChunkBytes, le.bytes == (uint32_t)kChunkBytes ? &tm128 : &tm64,
                          full_leader + 8 * slot, 0, row0);
          if (with_bias) {      // the layer's bias tile: my 4 KB half, same barrier as the chunk
            const uint32_t buf = (le2.y >> 1) & 1u, id = (le2.y >> 2) & 15u;
            tma_load_2d_cg2(sbase + kSmBiasBuf + buf * kBiasTileHalf, &tm32, full_leader + 8 * slot, 0,
                            (int)(((uint32_t)a.bias_off + id * kBiasTileBytes + rank * kBiasTileHalf) >> 7));
          }
          tr(0, 0xA00 + q);
        }
      }
    }
  } else if (COMP && (warp == 3 || warp == 2)) {
    // ===================== compositors: warp 3 takes tile 0, warp 2 tile 1 (two SM schedulers share the cost); each
    // consumes the [rgb logits, sigma] stage of its tile once the epilogue publishes it
    const int t = 3 - warp;
    constexpr bool vdc = kVD;
    const float4* stage = (vdc ? reinterpret_cast<const float4*>(s_par + kOffHeadW + 256)
                               : reinterpret_cast<const float4*>(smem + kSmEnc + 8192)) + t * 128;
    for (int it = 0; it < niter; ++it) {
      const long long st = cluster_id + (long long)it * a.nclusters;
      mbar_wait(bar(B_STFULL + t), (uint32_t)(it & 1), a.watch, 0xC00 + t);
      composite_tile(a, stage, st * kSuperRows + rank * kTileRows + 2 * t * kTileRows, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_STFREE + t));
    }
  } else if (warp == 3) {
    // ===================== (idle; debug observer of the leader's weight barriers when tracing)
    if (lane == 0 && rank == 0 && tracing) {
      // debug observer (leader's otherwise idle warp): when does each weights barrier really complete?
      long long gl = 0;
      for (int it = 0; it < niter; ++it)
        for (int q = 0; q < sc.nloads; ++q, ++gl) {
          const int slot = (int)(gl % kNS);
          mbar_wait(bar(B_FULL + slot), (uint32_t)((gl / kNS) & 1), a.watch, 0x280 + q);
          tr(3, 0xB00 + q);
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA)
    // tcgen05.mma issue is back-pressured by execution (the queue is ~1 MMA deep, see the
    // timeline in profiles/), so every cycle spent between two MMAs idles the tensor core:
    //  * the WHOLE warp runs this loop with warp-uniform control flow and operands derived
    //    only from uniform sources (kernel params, constants), so descriptors live in
    //    uniform registers; a divergent `if (lane == 0)` made ptxas wrap every UTCHMMA in a
    //    waterfall loop of 7 R2UR.BROADCAST (~170 cycles per MMA instead of 128);
    //  * ONE weights barrier per slot, no wait at all for the second tile's pass over a
    //    resident chunk, and the barrier probes of op o+1 are issued BEFORE the MMAs of op o.
    // ONE elected lane runs the whole persistent loop (the compiler then knows every operand is uniform and
    // keeps descriptors in uniform registers); the other lanes wait at the final __syncwarp.  The unit of
    // work is a BLOCK -- one tile's pass over the chunks of a layer: between two chunks of a block there is
    // only a table read (shared memory), at most one barrier probe, a few descriptor adds, 4 UTCHMMA and
    // the commits.  (The per-op bookkeeping of the first version cost ~500 cycles between MMA groups --
    // ~100 dependent instructions on one warp -- and idled the tensor core in every first-tile pass.)
    if (rank == 0 && elect_one()) {
      // the whole TMEM (512 columns) is allocated, so its base address is 0: keep it a constant
      if (tmem_base != 0) watchdog_fire(a.watch, 0xA12, tmem_base, 0);
      const uint32_t idesc256 = make_idesc_bf16(256, 256), idesc128 = make_idesc_bf16(256, 128);
      const uint64_t desc_hi = make_sdesc_sw128(0);            // everything but the start address
      const uint64_t d_ones = make_sdesc_nosw(sbase + kSmOnes, 128, 0);    // every 8-row group reads the same 8 rows
      const uint64_t d_bias0 = make_sdesc_nosw(sbase + kSmBiasBuf, 128, 256);
      const uint64_t d_bias1 = make_sdesc_nosw(sbase + kSmBiasBuf + kBiasTileHalf, 128, 256);
      const uint32_t* s_ops = reinterpret_cast<const uint32_t*>(smem + kSmOps);
      const int nops = sc.nops, nloads = sc.nloads;
      uint32_t act_cnt0 = 0, act_cnt1 = 0;
      int confirmed = 0;                                       // loads [0, confirmed) are known to have landed
      int gl0 = 0;
      for (int it = 0; it < niter; ++it, gl0 += nloads) {
        for (int o = 0; o < nops;) {
          uint32_t e = s_ops[o];
          const int tile = (e >> 8) & 0xF, blk = (e >> 12) & 0xF;
          if (blk == 0) watchdog_fire(a.watch, 0xA13, (uint32_t)o, e);
          if ((e >> 24) & F_FIRST) {
            const uint32_t cnt = tile ? act_cnt1 : act_cnt0;
            mbar_wait(bar(B_ACT + tile), cnt & 1, a.watch, 0x300 + o);
            if (tile) ++act_cnt1; else ++act_cnt0;
            tc_fence_after();
            tr(1, 0x300 + o);
          }
          const uint32_t d_tmem = (uint32_t)tile * 256u;
          const uint32_t act_base = sbase + kSmAct + tile * 4 * kChunkBytes;
          for (int k = 0; k < blk; ++k) {
            const uint32_t en = (k + 1 < blk) ? s_ops[o + k + 1] : 0u;   // next entry: its LDS overlaps the MMA issue
            const int asrc = (e >> 16) & 0xFF;
            const uint32_t flags = e >> 24;
            const int gl = gl0 + (int)(e & 0xFF);
            const int slot = gl % kNS;
            if (gl >= confirmed) {
              mbar_wait(bar(B_FULL + slot), (uint32_t)((gl / kNS) & 1), a.watch, 0x400 + o + k);
              confirmed = gl + 1;
              tc_fence_after();
              tr(1, 0x500 + o + k);
            }
            const uint32_t a_addr = (asrc < 4) ? (act_base + asrc * kChunkBytes) : (sbase + kSmEnc);
            const uint64_t da = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFF);
            const uint64_t db = desc_hi | (uint64_t)(((sbase + kSmRing + slot * kChunkBytes) >> 4) & 0x3FFF);
            const uint32_t idesc = (flags & F_N128) ? idesc128 : idesc256;
            const bool bias_mma = (flags & (F_BIAS0 | F_BIAS1)) != 0;
            if (bias_mma) umma_bf16<2>(d_tmem, d_ones, (flags & F_BIAS1) ? d_bias1 : d_bias0, idesc, 0u);
            umma_bf16<2>(d_tmem, da, db, idesc, ((flags & F_FIRST) && !bias_mma) ? 0u : 1u);
            umma_bf16<2>(d_tmem, da + 2, db + 2, idesc, 1u);
            umma_bf16<2>(d_tmem, da + 4, db + 4, idesc, 1u);
            umma_bf16<2>(d_tmem, da + 6, db + 6, idesc, 1u);
            if (flags & F_REL_SLOT) umma_commit<2>(bar(B_EMPTY + slot));
            if (flags & F_REL_ENC) umma_commit<2>(bar(B_ENC));
            if (flags & F_LAST) umma_commit<2>(bar(B_DONE + tile));
            e = en;
          }
          tr(1, 0x600 + o);
          o += blk;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue / encoding warps
    const int et = tid - 128;
    const int q = warp & 3, hh = (warp - 4) >> 2;
    const int row = 32 * q + lane;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;
    uint32_t done_cnt0 = 0, done_cnt1 = 0;      // (two scalars, not an array indexed by the run-time tile number: registers)
    uint32_t enc_cnt = 0;
    constexpr bool vd = kVD;

    auto enc_wait = [&]() {
      if (enc_cnt > 0) mbar_wait(bar(B_ENC), (enc_cnt - 1) & 1, a.watch, 0x600 + enc_cnt);
      ++enc_cnt;
    };
    // Publish tile t's freshly written operand chunks: make them visible to the async proxy,
    // (optionally) tell the MMA issuer, and -- training modes -- bulk-store `npanels` panels of
    // act_t to `gdst`.  The pre-barrier wait covers the PREVIOUS bulk store (other tile, issued a
    // whole epilogue ago), so after the barrier every thread may overwrite that tile's buffer.
    // Training stores are issued per 16 KB panel by the leader thread of the column half that wrote it (et == 0 for
    // columns 0..127, et == 128 for 128..255): the FIRST panel of a half goes out mid-epilogue (store_first_panel),
    // the second one here -- a 64 KB store issued only at the end finished reading shared memory too late and stalled
    // the next publish on its wait_group.read.
    const bool half_leader = (q == 0 && lane == 0);
    // Fused backward launch (a.flags != nullptr): every (tile, panel group, column half) this thread stores is TWO bulk
    // groups; when the second one is committed, everything older than the 4 most recent groups -- i.e. the entry before the
    // previous one -- has completed and its flag is released to the weight-gradient CTAs (two entries of lag: the wait
    // never stalls the epilogue).
    int sig_old = -1, sig_prev = -1;
    auto release_flag = [&](int idx) {
      asm volatile("fence.proxy.async;" ::: "memory");      // completed async-proxy writes before the generic-proxy flag
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + idx), "r"(a.epoch) : "memory");
    };
    auto entry_committed = [&](const uint8_t* gdst) {
      const size_t off = (size_t)(gdst - a.gstore);
      const size_t T = off / kBwdTileBytes;
      const uint32_t in = (uint32_t)(off - T * kBwdTileBytes);
      const int grp = in < (uint32_t)kGFeat ? 0 : (in < (uint32_t)kGY ? 1 : 2 + (int)((in - kGY) / (4 * kPanel)));
      bulk_wait4();
      if (sig_old >= 0) release_flag(sig_old);
      sig_old = sig_prev;
      sig_prev = (int)((T * kFlagGroups + grp) * 2 + hh);
    };
    auto publish = [&](int t, bool signal, uint8_t* gdst, int npanels) {
      tc_fence_before();
      fence_proxy_async_smem();
      // what must have left shared memory before the barrier: the OTHER tile's stores (its buffer is overwritten by
      // the next epilogue).  This tile's own first-panel store, issued half an epilogue ago, may still be in flight.
      if ((SAVE || BWD) && half_leader) { if (gdst != nullptr && npanels == 4) bulk_wait_read1(); else bulk_wait_read0(); }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (et == 0 && signal) { mbar_arrive_remote(bar(B_ACT + t), 0); tr(2, 0x900 + t); }
      if ((SAVE || BWD) && half_leader && gdst != nullptr) {
        const int pnl = (npanels == 4) ? 2 * hh + 1 : hh;          // 4 panels: my half's second; 2 panels: my chunk
        if (BWD && a.flags != nullptr && npanels == 2) {          // one panel per half: as two groups, like the others
          bulk_s2g(gdst + pnl * kPanel, sbase + kSmAct + (t * 4 + pnl) * kChunkBytes, kPanel / 2);
          bulk_commit();
          bulk_s2g(gdst + pnl * kPanel + kPanel / 2, sbase + kSmAct + (t * 4 + pnl) * kChunkBytes + kPanel / 2, kPanel / 2);
          bulk_commit();
        } else {
          bulk_s2g(gdst + pnl * kPanel, sbase + kSmAct + (t * 4 + pnl) * kChunkBytes, kPanel);
          bulk_commit();
        }
        if (BWD && a.flags != nullptr) entry_committed(gdst);
      }
    };
    // first panel (columns 128*hh .. 128*hh+63) of a 4-panel tile: complete after the second 32-column batch of my half
    auto store_first_panel = [&](int t, uint8_t* gdst) {
      fence_proxy_async_smem();
      if (hh == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
      else asm volatile("bar.sync 3, 128;" ::: "memory");
      if (half_leader) {
        bulk_s2g(gdst + 2 * hh * kPanel, sbase + kSmAct + (t * 4 + 2 * hh) * kChunkBytes, kPanel);
        bulk_commit();
      }
    };
    auto signal_ready = [&](int t) { publish(t, true, nullptr, 0); };
    // sample position of my row in tile t
    auto load_point = [&](long long grow, float (&p)[3]) {
      p[0] = p[1] = p[2] = 0.f;
      if (grow < a.nrows) {
        if (a.pts != nullptr) {
          p[0] = __ldg(a.pts + grow * 3 + 0); p[1] = __ldg(a.pts + grow * 3 + 1); p[2] = __ldg(a.pts + grow * 3 + 2);
        } else {
          const long long r = row_to_ray(a, grow);
          const float* ray = a.rays + r * a.rw;
          const float zz = (COMP && a.z == nullptr)
                               ? z_at_step((int)(grow - r * a.S), a.S, a.zstep, __ldg(ray + 6), __ldg(ray + 7), a.lindisp)
                               : __ldg(a.z + grow);
#pragma unroll
          for (int j = 0; j < 3; ++j) p[j] = __fadd_rn(__ldg(ray + j), __fmul_rn(__ldg(ray + 3 + j), zz));
        }
      }
    };
    // the same in two steps: the global loads now (7 registers), the arithmetic when the point is needed -- the prefetch
    // for the next super-tile (layer 6) must not wait for its own loads (an FMUL on the loaded ray was a top stall site)
    auto load_point_raw = [&](long long grow, float (&q)[7]) {
#pragma unroll
      for (int j = 0; j < 7; ++j) q[j] = 0.f;
      if (grow < a.nrows) {
        if (a.pts != nullptr) {
          q[0] = __ldg(a.pts + grow * 3 + 0); q[1] = __ldg(a.pts + grow * 3 + 1); q[2] = __ldg(a.pts + grow * 3 + 2);
        } else {
          const long long r = row_to_ray(a, grow);
          const float* ray = a.rays + r * a.rw;
          if (COMP && a.z == nullptr) {
            // coarse depths computed in-kernel: keep near / far and the sample index (as a float: exact below 2^24)
            q[6] = (float)(int)(grow - r * a.S);
          } else {
            q[6] = __ldg(a.z + grow);
          }
#pragma unroll
          for (int j = 0; j < 3; ++j) { q[j] = __ldg(ray + j); q[3 + j] = __ldg(ray + 3 + j); }
        }
      }
    };
    auto finish_point = [&](long long grow, const float (&q)[7], float (&p)[3]) {
      p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
      if (grow < a.nrows && a.pts == nullptr) {
        float zz = q[6];
        if (COMP && a.z == nullptr) {
          const long long r = row_to_ray(a, grow);
          const float* ray = a.rays + r * a.rw;
          zz = z_at_step((int)q[6], a.S, a.zstep, __ldg(ray + 6), __ldg(ray + 7), a.lindisp);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) p[j] = __fadd_rn(q[j], __fmul_rn(q[3 + j], zz));
      }
    };
    auto load_dir = [&](long long grow, float (&p)[3]) {
      p[0] = p[1] = p[2] = 0.f;
      if (grow < a.nrows) {
        const long long r = row_to_ray(a, grow);
        const float* src = a.viewdirs != nullptr ? a.viewdirs + r * 3 : a.rays + r * a.rw + 8;
        p[0] = __ldg(src); p[1] = __ldg(src + 1); p[2] = __ldg(src + 2);
      }
    };
    // gsave != nullptr (MODE_FWD_SAVE): also write my half row of the panel image to global
    // gamma(x) of tile t for LAYER 0 goes into chunk 0 of the tile's own activation buffer (free at that time), so
    // it can be produced right after the tile's last epilogue of the previous super-tile, under the other tile's
    // tail, instead of serialising both tiles through the shared enc buffer at the start of an iteration
    auto write_first_layer_x = [&](int t, const float (&p)[3], uint8_t* gsave) {
      uint32_t w[16];
      if (hh == 0) encode_cols<10, 0>(p, w); else encode_cols<10, 1>(p, w);
      store_row32(smem + kSmAct + t * 4 * kChunkBytes, row, 4 * hh, w);
      if (SAVE && gsave != nullptr) store_row32(gsave, row, 4 * hh, w);
    };
    auto write_enc_x = [&](long long grow, uint8_t* gsave) {
      float p[3];
      load_point(grow, p);
      uint32_t w[16];
      if (hh == 0) encode_cols<10, 0>(p, w); else encode_cols<10, 1>(p, w);
      store_row32(smem + kSmEnc, row, 4 * hh, w);
      if (SAVE && gsave != nullptr) store_row32(gsave, row, 4 * hh, w);
    };
    auto write_enc_d = [&](long long grow, uint8_t* gsave) {
      float p[3];
      load_dir(grow, p);
      uint32_t w[16];
      if (hh == 0) encode_cols<4, 0>(p, w); else encode_cols<4, 1>(p, w);
      store_row32(smem + kSmEnc, row, 4 * hh, w);
      if (SAVE && gsave != nullptr) store_row32(gsave, row, 4 * hh, w);
    };

    if constexpr (!BWD) {
    if (niter > 0) {      // layer-0 operands of the first super-tile
      const long long st0 = cluster_id;
      for (int t = 0; t < 2; ++t) {
        const long long g0 = st0 * kSuperRows + rank * kTileRows + 2 * t * kTileRows + row;
        float p0[3];
        load_point(g0, p0);
        write_first_layer_x(t, p0, SAVE ? a.acts + (size_t)(st0 * 4 + 2 * t + rank) * kFwdTileBytes + kFXg : nullptr);
        signal_ready(t);
      }
    }
    // rays / depths of my rows in the NEXT super-tile (loaded layers ahead, see load_point_raw); one named array per tile:
    // indexed by the run-time tile number they would live in local memory, and the store would wait for the loads
    float qnext0[7] = {}, qnext1[7] = {};
    for (int it = 0; it < niter; ++it) {
      const long long st = cluster_id + (long long)it * a.nclusters;
      // my row of tile t, and (training) the tile's stores: computed from the tile-0 value where they are used -- as
      // arrays indexed by the run-time tile number they lived in local memory, and the address arithmetic behind the LDL was
      // a long-scoreboard stall site of the forward-with-save epilogue (profiles/README.md)
      const long long grow0 = st * kSuperRows + rank * kTileRows + row;
      uint8_t* const gtile0 = SAVE ? a.acts + (size_t)(st * 4 + rank) * kFwdTileBytes : nullptr;
      uint8_t* const mtile0 = SAVE ? a.masks + (size_t)(st * 4 + rank) * kMaskTileBytes : nullptr;
#define GROW(t) (grow0 + (long long)(t) * (2 * kTileRows))
#define GTILE(t) (gtile0 + (size_t)(t) * (2 * (size_t)kFwdTileBytes))
#define MTILE(t) (mtile0 + (size_t)(t) * (2 * (size_t)kMaskTileBytes))
      for (int l = 0; l < sc.nlayers; ++l) {
        for (int t = 0; t < 2; ++t) {
          mbar_wait(bar(B_DONE + t), (t == 0 ? done_cnt0 : done_cnt1) & 1, a.watch, 0x700 + l * 2 + t);
          if (t == 0) ++done_cnt0; else ++done_cnt1;
          if (et == 0) tr(2, 0x700 + l * 2 + t);
          tc_fence_after();
          const uint32_t taddr = tmem_base + lane_addr + t * 256;
          if (l <= 8 && !(l == 7 && !vd && !(SAVE && NV))) {
            // ---- hidden layer: bias (+ReLU) -> bf16 -> next A operand; layer 7 also feeds the density head
            const float* hw = s_par + kOffHeadW + 128 * hh;
            uint8_t* act_t = smem + kSmAct + t * 4 * kChunkBytes;
            float sig = 0.f;
            float part[5] = {0.f, 0.f, 0.f, 0.f, 0.f};   // MODE_FWD_SAVE without view directions: output_linear partials
            const bool out_head = SAVE && NV && l == 7;
            uint32_t sbits[4] = {0u, 0u, 0u, 0u};     // MODE_FWD_SAVE: sign bits of my 4 x 32 pre-activations
            // software pipeline: the TMEM load of batch b+1 is in flight while batch b is processed
            uint32_t v[2][32];
#ifndef NERFB_EXP_NOTMEM
            tmem_ld32(taddr + 128 * hh, v[0]);
#else
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[0][i] = row * 77 + i; v[1][i] = row * 31 + i; }
#endif
#pragma unroll
            for (int b = 0; b < 4; ++b) {
#ifndef NERFB_EXP_NOTMEM
              tmem_ld_wait();
              if (b < 3) tmem_ld32(taddr + 128 * hh + 32 * (b + 1), v[(b + 1) & 1]);
#endif
              const uint32_t (&vv)[32] = v[b & 1];
              // (the layer's bias is already in the accumulator: bias MMA / column 63 of the encoding chunk)
              float x[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(vv[j]);
              if (SAVE && l < 8) {
                // one funnel shift per unit (column j ends at bit 31 - j), as four independent 8-bit chains so the
                // shifts do not serialise on one register
                uint32_t m4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
#pragma unroll
                  for (int k = 0; k < 4; ++k) m4[k] = __funnelshift_l(__float_as_uint(x[8 * k + jj]), m4[k], 1);
                sbits[b] = (m4[0] << 24) | (m4[1] << 16) | (m4[2] << 8) | m4[3];
              }
              if (SAVE && NV && out_head) {
                // output_linear (out_ch x 256) in fp32 from the un-rounded layer-7 activations, as in the plain forward
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float xr = fmaxf(x[j], 0.f);
#pragma unroll
                  for (int k = 0; k < 5; ++k) part[k] = fmaf(xr, hw[256 * k + 32 * b + j], part[k]);
                }
              } else if (l == 7) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 ww = *reinterpret_cast<const float4*>(hw + 32 * b + 4 * j4);
                  sig = fmaf(fmaxf(x[4 * j4 + 0], 0.f), ww.x, sig); sig = fmaf(fmaxf(x[4 * j4 + 1], 0.f), ww.y, sig);
                  sig = fmaf(fmaxf(x[4 * j4 + 2], 0.f), ww.z, sig); sig = fmaf(fmaxf(x[4 * j4 + 3], 0.f), ww.w, sig);
                }
              }
              uint32_t w[16];
              if (l != 8) {
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2_relu(x[2 * i], x[2 * i + 1]);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
              }
              const int col0 = 128 * hh + 32 * b;
#ifndef NERFB_EXP_NOSTS
              store_row32(act_t + (col0 >> 6) * kChunkBytes, row, (col0 & 63) >> 3, w);
              if (SAVE && b == 1) store_first_panel(t, GTILE(t) + (l < 8 ? kFH + l * 4 * kPanel : kFFeat));
#else
              if (w[0] == 0x12345678u) store_row32(act_t + (col0 >> 6) * kChunkBytes, row, (col0 & 63) >> 3, w);
#endif
            }
            if (l == 7) s_sig[(t * 2 + hh) * 128 + row] = sig;
            if (SAVE && l < 8)
              *reinterpret_cast<uint4*>(MTILE(t) + ((size_t)l * 128 + row) * kMaskRowBytes + 16 * hh) =
                  make_uint4(sbits[0], sbits[1], sbits[2], sbits[3]);
            if (SAVE && NV && out_head) {
              // the two column halves meet in the (idle) enc buffer: [tile][half][row][8]
              float* xch = reinterpret_cast<float*>(smem + kSmEnc);
#pragma unroll
              for (int k = 0; k < 5; ++k) xch[((t * 2 + hh) * 128 + row) * 8 + k] = part[k];
              asm volatile("bar.sync 1, 256;" ::: "memory");
              if (hh == 0 && GROW(t) < a.nrows) {
                for (int k = 0; k < a.out_ch; ++k)
                  a.raw[GROW(t) * a.out_ch + k] = xch[((t * 2 + 0) * 128 + row) * 8 + k] +
                                                  xch[((t * 2 + 1) * 128 + row) * 8 + k] + s_par[kOffHeadB + k];
              }
            }
          } else if (l == 7) {
            // ---- no view directions: output_linear (out_ch x 256) in fp32 from layer-7 accumulators
            float part[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              uint32_t v[32];
              tmem_ld32(taddr + 128 * hh + 32 * b, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float x = fmaxf(__uint_as_float(v[j]), 0.f);      // bias: in the accumulator (bias MMA)
#pragma unroll
                for (int k = 0; k < 5; ++k) part[k] = fmaf(x, s_par[kOffHeadW + 256 * k + 128 * hh + 32 * b + j], part[k]);
              }
            }
            if constexpr (COMP) {
              // enc buffer (idle until the next super-tile's skip layer): bytes 0..4095 = the hh == 1 partial sums
              // [tile][row] float4, bytes 8192..12287 = the stage [tile][row] float4 the compositor warp reads
              float4* xch = reinterpret_cast<float4*>(smem + kSmEnc) + (t * 128 + row);
              float4* stg = reinterpret_cast<float4*>(smem + kSmEnc + 8192) + (t * 128 + row);
              if (it > 0) mbar_wait(bar(B_STFREE + t), (uint32_t)((it - 1) & 1), a.watch, 0xC10 + t);
              if (hh == 1) *xch = make_float4(part[0], part[1], part[2], part[3]);
              asm volatile("bar.sync 1, 256;" ::: "memory");
              if (hh == 0) {
                const float4 o = *xch;
                *stg = make_float4(part[0] + o.x + s_par[kOffHeadB + 0], part[1] + o.y + s_par[kOffHeadB + 1],
                                   part[2] + o.z + s_par[kOffHeadB + 2], part[3] + o.w + s_par[kOffHeadB + 3]);
              }
            } else {
            // exchange the two column halves through the (idle) enc buffer: [tile][half][row][8]
            float* xch = reinterpret_cast<float*>(smem + kSmEnc);
#pragma unroll
            for (int k = 0; k < 5; ++k) xch[((t * 2 + hh) * 128 + row) * 8 + k] = part[k];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (hh == 0 && GROW(t) < a.nrows) {
              for (int k = 0; k < a.out_ch; ++k)
                a.raw[GROW(t) * a.out_ch + k] = xch[((t * 2 + 0) * 128 + row) * 8 + k] +
                                                xch[((t * 2 + 1) * 128 + row) * 8 + k] + s_par[kOffHeadB + k];
            }
            }
          } else {
            // ---- l == 9: view layer epilogue + rgb head -> raw.  Both column halves work: thread (row, hh) takes
            // view-layer columns 64*hh .. 64*hh+63; the hh == 1 partial rgb sums cross over through a scratch area
            // (the unused output_linear rows of the parameter block) and hh == 0 writes the sample's raw[4].
            {
              // (two accumulator sets per colour and both TMEM loads issued up front: this epilogue is bound by
              //  instruction latency, not by TMEM bandwidth)
              float ca[3] = {0.f, 0.f, 0.f}, cb[3] = {0.f, 0.f, 0.f};
              uint32_t sbits[2] = {0u, 0u};
              uint32_t v2[2][32];
              tmem_ld32(taddr + 64 * hh, v2[0]);
              tmem_ld32(taddr + 64 * hh + 32, v2[1]);
              tmem_ld_wait();
#pragma unroll
              for (int b = 0; b < 2; ++b) {
                float xs[32];
                if (hh == 0) rgb_head_half<0, SAVE>(v2[b], b, ca, cb, xs, sbits[b]);
                else rgb_head_half<1, SAVE>(v2[b], b, ca, cb, xs, sbits[b]);
                if (SAVE) {      // hv (bf16) for the rgb-head / views-layer gradients
                  uint32_t w[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(xs[2 * i], xs[2 * i + 1]);
                  store_row32(GTILE(t) + kFHv + hh * kPanel, row, 4 * b, w);
                }
              }
              const float c0 = ca[0] + cb[0], c1 = ca[1] + cb[1], c2 = ca[2] + cb[2];
              if (SAVE)
                *reinterpret_cast<uint2*>(MTILE(t) + ((size_t)8 * 128 + row) * kMaskRowBytes + 8 * hh) =
                    make_uint2(sbits[0], sbits[1]);
              if constexpr (COMP) {
                // stage [tile][row] float4 in the unused output_linear rows of the parameter block: the hh == 1 partial
                // sums cross over through it, then it holds the sample's [rgb logits, sigma] for the compositor warp
                float4* stg = reinterpret_cast<float4*>(s_par + kOffHeadW + 256) + (t * 128 + row);
                if (it > 0) mbar_wait(bar(B_STFREE + t), (uint32_t)((it - 1) & 1), a.watch, 0xC10 + t);
                if (hh == 1) *stg = make_float4(c0, c1, c2, 0.f);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (hh == 0) {
                  const float4 o = *stg;
                  const float sigma = s_sig[(t * 2 + 0) * 128 + row] + s_sig[(t * 2 + 1) * 128 + row] + s_par[kOffHeadB];
                  *stg = make_float4(c0 + o.x + s_par[kOffRgbB], c1 + o.y + s_par[kOffRgbB + 1],
                                     c2 + o.z + s_par[kOffRgbB + 2], sigma);
                }
              } else {
              float* xch = s_par + kOffHeadW + 256 + (t * 128 + row) * 3;
              if (hh == 1) { xch[0] = c0; xch[1] = c1; xch[2] = c2; }
              asm volatile("bar.sync 1, 256;" ::: "memory");
              if (hh == 0 && GROW(t) < a.nrows) {
                const float sigma = s_sig[(t * 2 + 0) * 128 + row] + s_sig[(t * 2 + 1) * 128 + row] + s_par[kOffHeadB];
                *reinterpret_cast<float4*>(a.raw + GROW(t) * 4) =
                    make_float4(c0 + xch[0] + s_par[kOffRgbB], c1 + xch[1] + s_par[kOffRgbB + 1],
                                c2 + xch[2] + s_par[kOffRgbB + 2], sigma);
              }
              }
            }
          }
          // ---- operands the NEXT layer needs besides the activations
          if (l == 4) {
            if (COMP && !vd && it > 0) {      // the encoding overwrites BOTH tiles' stage (enc buffer): both must be consumed
              mbar_wait(bar(B_STFREE + 0), (uint32_t)((it - 1) & 1), a.watch, 0xC20);
              mbar_wait(bar(B_STFREE + 1), (uint32_t)((it - 1) & 1), a.watch, 0xC21);
            }
            enc_wait(); write_enc_x(GROW(t), nullptr);
          }
          if (l == 6 && it + 1 < niter)       // global loads whose latency must not sit in front of the next encode
          {
            const long long gn = (st + a.nclusters) * kSuperRows + rank * kTileRows + 2 * t * kTileRows + row;
            if (t == 0) load_point_raw(gn, qnext0); else load_point_raw(gn, qnext1);
          }
          if (l == 8) { enc_wait(); write_enc_d(GROW(t), SAVE ? GTILE(t) + kFDg : nullptr); }
          if (et == 0) tr(2, 0x800 + l * 2 + t);
          if (l + 1 < sc.nlayers) {
            // training: h_l (l < 8) / feature (l == 8) go to the tile store as written
            uint8_t* gdst = nullptr;
            if (SAVE) gdst = GTILE(t) + (l < 8 ? kFH + l * 4 * kPanel : kFFeat);
            publish(t, true, gdst, 4);
            // MODE_FWD_COMP: the stage this tile's final epilogue wrote in the PREVIOUS super-tile is handed to its
            // compositor warp only now, inside the middle layers, not during the epilogue-bound super-tile boundary
            // (publish() contains the CTA barrier that orders the 128 writers before this arrive)
            if (COMP && l == 1 && it > 0 && et == 0) mbar_arrive(bar(B_STFULL + t));
          } else {
            // last layer of this tile.  Without view directions h7 is the last stored layer input: the second panels of
            // both halves go out here (nothing to signal, no MMA follows).
            if (SAVE && NV) publish(t, false, GTILE(t) + kFH + 7 * 4 * kPanel, 4);
            if (it + 1 < niter) {
              // the tile's buffers are free -> layer-0 operand of the NEXT super-tile right away
              if (SAVE) {           // ... once the bulk stores of the last stored layer have left the buffer
                if (half_leader) bulk_wait_read0();
                asm volatile("bar.sync 1, 256;" ::: "memory");
              }
              const long long stn = st + a.nclusters;
              float pn[3];
              float qn[7];
#pragma unroll
              for (int j = 0; j < 7; ++j) qn[j] = t == 0 ? qnext0[j] : qnext1[j];
              finish_point(stn * kSuperRows + rank * kTileRows + 2 * t * kTileRows + row, qn, pn);
              write_first_layer_x(t, pn,
                                  SAVE ? a.acts + (size_t)(stn * 4 + 2 * t + rank) * kFwdTileBytes + kFXg : nullptr);
              signal_ready(t);
            } else {
              tc_fence_before();
              asm volatile("bar.sync 1, 256;" ::: "memory");
              if (COMP && et == 0) mbar_arrive(bar(B_STFULL + t));      // last super-tile: nothing left to hide behind
            }
          }
        }
      }
    }
    } else {
#undef GROW
#undef GTILE
#undef MTILE
      // ================= MODE_BWD: data-gradient chain  d_raw -> d_hv -> d_feature -> dY_7 .. dY_0
      // Layer j of the schedule multiplies by a transposed weight matrix; its epilogue applies the ReLU
      // mask of the layer it lands on (sign bits collected by the forward-with-save pass: 16 bytes per
      // thread and tile-layer, expanded with prmt), writes the bf16 gradient as the next A operand and
      // bulk-stores it for the weight-gradient kernel.
      // Without view directions (MODE_BWD_NV) the first stage is d_h7 = (d_raw W_out) * [h7 > 0] on CUDA cores and the
      // chain is W7^T .. W1^T (schedule layers j = 2..8).
      constexpr bool vdb = !NV;
      constexpr int j_first = vdb ? 0 : 2, j_end = 9;
      const float* w_rgb = s_par + kOffRgbW;
      const float* w_alpha = s_par + kOffHeadW;     // alpha_linear.weight, or output_linear.weight [5][256] (rows >= out_ch zero)
      // d_raw row and first-stage sign bits of my rows in the next super-tile, requested layers ahead.  One named struct
      // per tile: as arrays indexed by the run-time tile number they lived in local memory, and the store behind each
      // load waited for it -- the prefetch stalled where it was issued
      struct FirstStage {
        float4 dr;
        uint2 mhv;       // view layer: my 64 columns
        uint4 mh7;       // no view directions: h7, my 128 columns
        float d4;        // no view directions, out_ch = 5: gradient of output channel 4
      };
      FirstStage fs0 = {}, fs1 = {};
      auto prefetch_into = [&](long long stn, int t, FirstStage& f) {
        const long long gn = stn * kSuperRows + rank * kTileRows + 2 * t * kTileRows + row;
        const uint8_t* mt = a.masks + (size_t)(stn * 4 + 2 * t + rank) * kMaskTileBytes;
        f.dr = make_float4(0.f, 0.f, 0.f, 0.f);
        f.d4 = 0.f;
        if (gn < a.nrows) {
          f.dr = __ldg(reinterpret_cast<const float4*>(a.d_raw) + gn);
          if (!vdb && a.d_raw4 != nullptr) f.d4 = __ldg(a.d_raw4 + gn);
        }
        if (vdb) f.mhv = __ldg(reinterpret_cast<const uint2*>(mt + ((size_t)8 * 128 + row) * kMaskRowBytes) + hh);
        else f.mh7 = __ldg(reinterpret_cast<const uint4*>(mt + ((size_t)7 * 128 + row) * kMaskRowBytes + 16 * hh));
      };
      auto prefetch_first_stage = [&](long long stn, int t) {
        if (t == 0) prefetch_into(stn, 0, fs0); else prefetch_into(stn, 1, fs1);
      };
      for (int it = 0; it < niter; ++it) {
        const long long st = cluster_id + (long long)it * a.nclusters;
        const long long grow0 = st * kSuperRows + rank * kTileRows + row;
        const uint8_t* const mtile0 = a.masks + (size_t)(st * 4 + rank) * kMaskTileBytes;
        uint8_t* const gtile0 = a.gstore + (size_t)(st * 4 + rank) * kBwdTileBytes;
#define GROW(t) (grow0 + (long long)(t) * (2 * kTileRows))
#define GTILE(t) (gtile0 + (size_t)(t) * (2 * (size_t)kBwdTileBytes))
#define MTILE(t) (mtile0 + (size_t)(t) * (2 * (size_t)kMaskTileBytes))
        float dsig0 = 0.f, dsig1 = 0.f;
        if (it == 0) {        // first super-tile: nothing was prefetched yet
          for (int t = 0; t < 2; ++t) prefetch_first_stage(st, t);
        }
        for (int t = 0; t < 2; ++t) {
          const float4 dr = t == 0 ? fs0.dr : fs1.dr;
          if (t == 0) dsig0 = dr.w; else dsig1 = dr.w;
          if (!vdb) {
            // ---- d_h7 = (sum_k d_raw[k] W_out[k,:]) * [h7 > 0]: my 128 columns (chunks 2hh, 2hh+1) of the K=256 operand
            const float d4 = t == 0 ? fs0.d4 : fs1.d4;
            const uint4 m7 = t == 0 ? fs0.mh7 : fs1.mh7;
            const uint32_t m7v[4] = {m7.x, m7.y, m7.z, m7.w};
            uint8_t* act_t = smem + kSmAct + t * 4 * kChunkBytes;
            uint8_t* gdst = GTILE(t) + kGY + 7 * 4 * kPanel;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const float* wo = w_alpha + 128 * hh + 32 * b;
              uint32_t w[16];
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                const float dk[5] = {dr.x, dr.y, dr.z, dr.w, d4};
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wo + 256 * k + 4 * i4);
                  g[0] = fmaf(dk[k], w4.x, g[0]); g[1] = fmaf(dk[k], w4.y, g[1]);
                  g[2] = fmaf(dk[k], w4.z, g[2]); g[3] = fmaf(dk[k], w4.w, g[3]);
                }
                w[2 * i4] = pack_bf16x2(g[0], g[1]);
                w[2 * i4 + 1] = pack_bf16x2(g[2], g[3]);
              }
              apply_sign_mask32(w, m7v[b]);
              const int col0 = 128 * hh + 32 * b;
              store_row32(act_t + (col0 >> 6) * kChunkBytes, row, (col0 & 63) >> 3, w);
              if (b == 1) store_first_panel(t, gdst);
            }
            publish(t, true, gdst, 4);
            continue;
          }
          // ---- d_hv = (d_rgb W_rgb) * [hv > 0]: my 64 columns (chunk hh) of the K=128 operand
          // hv sign bits: 4 words for columns 0..127; my chunk uses words 2*hh and 2*hh+1
          const uint2 mhv = t == 0 ? fs0.mhv : fs1.mhv;
          uint8_t* act_t = smem + kSmAct + t * 4 * kChunkBytes;
#pragma unroll
          for (int half = 0; half < 2; ++half) {          // 32 columns at a time
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int c = 64 * hh + 32 * half + 2 * i;
              const float g0 = fmaf(dr.z, w_rgb[256 + c], fmaf(dr.y, w_rgb[128 + c], dr.x * w_rgb[c]));
              const float g1 = fmaf(dr.z, w_rgb[257 + c], fmaf(dr.y, w_rgb[129 + c], dr.x * w_rgb[c + 1]));
              w[i] = pack_bf16x2(g0, g1);
            }
            apply_sign_mask32(w, half == 0 ? mhv.x : mhv.y);
            store_row32(act_t + hh * kChunkBytes, row, 4 * half, w);
          }
          publish(t, true, GTILE(t) + kGHv, 2);
        }
        // the 16 bytes of sign bits a thread needs for a tile-layer are requested one whole epilogue ahead (the chain
        // is epilogue-bound, so a load issued right before its use would expose the full DRAM latency every time)
        auto load_mask_words = [&](int jj, int tt) -> uint4 {
          return __ldg(reinterpret_cast<const uint4*>(MTILE(tt) + ((size_t)(8 - jj) * 128 + row) * kMaskRowBytes + 16 * hh));
        };
        uint4 mw_next = load_mask_words(vdb ? 1 : 2, 0);         // first masked tile-layer: j = 1 (2 without views), t = 0
        for (int j = j_first; j < j_end; ++j) {
          const int m = 8 - j;                       // mask layer: j=1 -> h7 ... j=8 -> h0 ; j=0: feature (no ReLU)
          for (int t = 0; t < 2; ++t) {
            const int colbase = 128 * hh;
#ifdef NERFB_EXP_NOMASK
            const bool use_mask = false;
#else
            const bool use_mask = j > 0;
#endif
            uint4 mw = make_uint4(0u, 0u, 0u, 0u);
            if (use_mask) {
              mw = mw_next;
              const int jn = t == 0 ? j : j + 1, tn = t ^ 1;          // next masked tile-layer
              if (jn < j_end) mw_next = load_mask_words(jn, tn);
            }
            mbar_wait(bar(B_DONE + t), (t == 0 ? done_cnt0 : done_cnt1) & 1, a.watch, 0x700 + j * 2 + t);
            if (t == 0) ++done_cnt0; else ++done_cnt1;
            if (et == 0) tr(2, 0x700 + j * 2 + t);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_addr + t * 256;
            uint8_t* act_t = smem + kSmAct + t * 4 * kChunkBytes;
            const float ds = t == 0 ? dsig0 : dsig1;
            const uint32_t mwv[4] = {mw.x, mw.y, mw.z, mw.w};
            uint32_t v[2][32];
            tmem_ld32(taddr + colbase, v[0]);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              tmem_ld_wait();
              if (b < 3) tmem_ld32(taddr + colbase + 32 * (b + 1), v[(b + 1) & 1]);
              const uint32_t (&vv)[32] = v[b & 1];
              uint32_t w[16];
              if (j == 1) {       // + d_sigma * w_alpha (vector loads: the shared-memory pipe is what this epilogue waits on)
                const float* wa = w_alpha + colbase + 32 * b;
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                  const float4 a4 = *reinterpret_cast<const float4*>(wa + 4 * i4);
                  w[2 * i4] = pack_bf16x2(fmaf(ds, a4.x, __uint_as_float(vv[4 * i4])),
                                          fmaf(ds, a4.y, __uint_as_float(vv[4 * i4 + 1])));
                  w[2 * i4 + 1] = pack_bf16x2(fmaf(ds, a4.z, __uint_as_float(vv[4 * i4 + 2])),
                                              fmaf(ds, a4.w, __uint_as_float(vv[4 * i4 + 3])));
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = pack_bf16x2(__uint_as_float(vv[2 * i]), __uint_as_float(vv[2 * i + 1]));
              }
              if (use_mask) apply_sign_mask32(w, mwv[b]);
              if (b == 3 && j == 6 && it + 1 < niter)       // next super-tile's first-stage inputs, two layers early
                prefetch_first_stage(st + a.nclusters, t);
              const int col0 = colbase + 32 * b;
              store_row32(act_t + (col0 >> 6) * kChunkBytes, row, (col0 & 63) >> 3, w);
              if (b == 1) store_first_panel(t, GTILE(t) + (j == 0 ? kGFeat : kGY + m * 4 * kPanel));
            }
            if (et == 0) tr(2, 0x800 + j * 2 + t);
            uint8_t* gdst = GTILE(t) + (j == 0 ? kGFeat : kGY + m * 4 * kPanel);
            publish(t, j + 1 < j_end, gdst, 4);
          }
        }
      }
    }
    if ((SAVE || BWD) && half_leader) {
      bulk_wait0();
      if (BWD && a.flags != nullptr) {
        if (sig_old >= 0) release_flag(sig_old);
        if (sig_prev >= 0) release_flag(sig_prev);
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

#undef GROW
#undef GTILE
#undef MTILE

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
nerf_mlp_kernel(const __grid_constant__ MlpArgs a, const __grid_constant__ Sched sc,
                const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm64,
                const __grid_constant__ CUtensorMap tm32) {
  mlp_body<MODE>(a, sc, tm128, tm64, tm32);
}

// The backward of a network pass in ONE launch: the first `chain_ctas` CTAs (2-CTA clusters) run the data-gradient chain,
// the others the weight-gradient jobs, which take the chain's dY panels through per-tile completion flags while they are
// still in L2 (wgrad_body.cuh).  The weight-gradient kernel alone is HBM-bound on reading dY and the layer inputs
// (10.6 KB per sample); here dY is read from L2 and the tensor work of both halves overlaps.  All CTAs must be
// co-resident (consumers spin on flags): the launch is cooperative and the grid is the number of SMs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
nerf_bwd_fused_kernel(const __grid_constant__ MlpArgs a, const __grid_constant__ Sched sc,
                      const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm64,
                      const __grid_constant__ CUtensorMap tm32, const __grid_constant__ WArgs w, const int chain_ctas) {
  if ((int)blockIdx.x < chain_ctas) mlp_body<MODE_BWD>(a, sc, tm128, tm64, tm32);
  else wgrad_body<true>(w, (int)blockIdx.x - chain_ctas);
}

// ------------------------------------------------------------- packing
struct PackChunk {
  const float* w;      // weight matrix [rows, ld]
  int ld, col0, ncols; // source columns col0 .. col0+ncols (rest of the 64 are zero)
  int nrows;           // 256 or 128 (two CTA halves of nrows/2)
  uint32_t off;        // blob offset
};
struct PackArgs {
  PackChunk ch[kNumFullChunks + kNumViewChunks];
  int nchunks;
  const float* bias_tile[kNumBiasTilesVD];   // the layers' biases in blob order
  int bias_rows[kNumBiasTilesVD];            // 256, or 128 for the views layer
  int nbias;
  uint32_t bias_area;
  const float* src[8];  // fp32 block sources, see pack kernel
  int out_ch, use_viewdirs;
  const float* bias[10];
};

__device__ void fill_param_block(const PackArgs& pa, float* par) {
  for (int i = threadIdx.x; i < kParamFloats; i += blockDim.x) {
    float v = 0.f;
    if (i < kOffBiasFeat) {
      v = pa.bias[i >> 8][i & 255];
    } else if (i < kOffBiasViews) {
      if (pa.use_viewdirs) v = pa.bias[8][i - kOffBiasFeat];
    } else if (i < kOffHeadW) {
      if (pa.use_viewdirs) v = pa.bias[9][i - kOffBiasViews];
    } else if (i < kOffRgbW) {
      const int j = i - kOffHeadW;              // head weights [k][256]
      const int rows = pa.use_viewdirs ? 1 : pa.out_ch;
      if (j < rows * 256) v = pa.src[0][j];
    } else if (i < kOffHeadB) {
      if (pa.use_viewdirs) v = pa.src[2][i - kOffRgbW];   // rgb_linear.weight [3,128]
    } else if (i < kOffRgbB) {
      const int j = i - kOffHeadB;
      const int rows = pa.use_viewdirs ? 1 : pa.out_ch;
      if (j < rows) v = pa.src[1][j];
    } else {
      const int j = i - kOffRgbB;
      if (pa.use_viewdirs && j < 3) v = pa.src[3][j];
    }
    par[i] = v;
  }
}

__global__ void pack_weights_kernel(const __grid_constant__ PackArgs pa, uint8_t* __restrict__ blob, int chunk_area) {
  const int c = blockIdx.y;
  if (c < pa.nchunks) {
    const PackChunk pc = pa.ch[c];
    const int half_rows = pc.nrows / 2;
    // one thread per 16-byte unit: nrows x 8 units
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < pc.nrows * 8; idx += gridDim.x * blockDim.x) {
      const int n = idx >> 3, u = idx & 7;
      const int half = n / half_rows, r = n % half_rows;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = 8 * u + j;
        f[j] = (k < pc.ncols) ? __ldg(pc.w + (size_t)n * pc.ld + pc.col0 + k) : 0.f;
      }
      uint4 v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                           pack_bf16x2(f[6], f[7]));
      *reinterpret_cast<uint4*>(blob + pc.off + (size_t)half * half_rows * 128 + sw128_offset(r, u)) = v;
    }
  } else if (c < pa.nchunks + pa.nbias) {
    // bias tile of a layer: [N x 16] bf16, column 0 = bf16(bias), column 1 = bf16(bias - column 0), rest zero; two CTA
    // halves of N/2 rows in 4 KB each (rows beyond N/2 zero), no-swizzle K-major layout (tc05::nosw16_offset)
    const int tix = c - pa.nchunks;
    const int half_rows = pa.bias_rows[tix] / 2;
    uint8_t* dst = blob + pa.bias_area + (size_t)tix * kBiasTileBytes;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * 128 * 16; idx += gridDim.x * blockDim.x) {
      const int half = idx >> 11, r = (idx >> 4) & 127, k = idx & 15;
      float v = 0.f;
      if (r < half_rows && k < 2) {
        const float b = __ldg(pa.bias_tile[tix] + half * half_rows + r);
        const float hi = __bfloat162float(__float2bfloat16(b));
        v = (k == 0) ? hi : (b - hi);
      }
      *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)half * kBiasTileHalf + nosw16_offset(r, k)) = __float2bfloat16(v);
    }
  } else if (blockIdx.x == 0) {
    fill_param_block(pa, reinterpret_cast<float*>(blob + chunk_area));
  }
}

// Transposed weight images for the data-gradient chain: chunk = B operand [n = 256 rows (two CTA
// halves), k = 64 columns] with B[n][k] = W[(row0 + k) * ld + col0 + n]  (i.e. W^T).
struct PackChunkT { const float* w; int ld, row0, col0; };
struct PackArgsT { PackChunkT ch[kNumBwdChunks]; };

__global__ void pack_weights_bwd_kernel(const __grid_constant__ PackArgsT pt, const __grid_constant__ PackArgs pa,
                                        uint8_t* __restrict__ blob) {
  const int c = blockIdx.y;
  if (c < kNumBwdChunks) {
    const PackChunkT pc = pt.ch[c];
    if (pc.w == nullptr) return;                   // chunk not used by this architecture (no view directions)
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 256 * 8; idx += gridDim.x * blockDim.x) {
      const int n = idx & 255, u = idx >> 8;       // n fastest: coalesced reads along a weight row
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __ldg(pc.w + (size_t)(pc.row0 + 8 * u + j) * pc.ld + pc.col0 + n);
      const uint4 v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                 pack_bf16x2(f[6], f[7]));
      *reinterpret_cast<uint4*>(blob + (size_t)c * kChunkFull + (size_t)(n >> 7) * kChunkBytes +
                                sw128_offset(n & 127, u)) = v;
    }
  } else if (blockIdx.x == 0) {
    fill_param_block(pa, reinterpret_cast<float*>(blob + kChunkAreaBwd));
  }
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_set_trace(long long* dev_buf) {
  g_trace = dev_buf;
  return NERFB_OK;
}

extern "C" size_t nerfb_packed_weights_bytes(int use_viewdirs) {
  return (size_t)(use_viewdirs ? kParamOffVD : kParamOffNoVD) + kParamBytes;
}

extern "C" int nerfb_pack_weights(const float* const* h_params, int use_viewdirs, int out_ch, void* packed,
                                  void* stream) {
  NERFB_REQUIRE(h_params && packed, "nerfb_pack_weights: null argument");
  for (int i = 0; i < 16; ++i) NERFB_REQUIRE(h_params[i], "nerfb_pack_weights: pts_linears parameter %d missing", i);
  PackArgs pa{};
  pa.use_viewdirs = use_viewdirs;
  pa.out_ch = out_ch;
  int n = 0;
  auto add = [&](const float* w, int ld, int col0, int ncols, int nrows) {
    const uint32_t off = n < kNumFullChunks ? (uint32_t)n * kChunkFull
                                            : (uint32_t)kNumFullChunks * kChunkFull + (uint32_t)(n - kNumFullChunks) * kChunkViews;
    pa.ch[n++] = PackChunk{w, ld, col0, ncols, nrows, off};
  };
  auto W = [&](int layer) { return h_params[2 * layer]; };
  for (int l = 0; l < 8; ++l) { pa.bias_tile[pa.nbias] = h_params[2 * l + 1]; pa.bias_rows[pa.nbias++] = 256; }   // bias tiles
  add(W(0), 63, 0, 63, 256);                                         // chunk 0
  for (int l = 1; l <= 4; ++l)
    for (int k = 0; k < 4; ++k) add(W(l), 256, 64 * k, 64, 256);     // 1..16
  add(W(5), 319, 0, 63, 256);                                        // 17: gamma(x) part of the skip layer
  for (int k = 0; k < 4; ++k) add(W(5), 319, 63 + 64 * k, 64, 256);  // 18..21
  for (int l = 6; l <= 7; ++l)
    for (int k = 0; k < 4; ++k) add(W(l), 256, 64 * k, 64, 256);     // 22..29
  for (int l = 0; l < 8; ++l) pa.bias[l] = h_params[2 * l + 1];
  if (use_viewdirs) {
    const float* vw = h_params[NERFB_P_VIEWS_W];
    const float* fw = h_params[NERFB_P_FEATURE_W];
    NERFB_REQUIRE(vw && fw && h_params[NERFB_P_VIEWS_B] && h_params[NERFB_P_FEATURE_B] && h_params[NERFB_P_ALPHA_W] &&
                      h_params[NERFB_P_ALPHA_B] && h_params[NERFB_P_RGB_W] && h_params[NERFB_P_RGB_B],
                  "nerfb_pack_weights: view-dependent head parameters missing");
    for (int k = 0; k < 4; ++k) add(fw, 256, 64 * k, 64, 256);       // 30..33 feature_linear
    add(vw, 283, 256, 27, 128);                                      // 34: gamma(d) part [feature | gamma(d)]
    pa.bias_tile[pa.nbias] = h_params[NERFB_P_FEATURE_B]; pa.bias_rows[pa.nbias++] = 256;
    pa.bias_tile[pa.nbias] = h_params[NERFB_P_VIEWS_B]; pa.bias_rows[pa.nbias++] = 128;
    for (int k = 0; k < 4; ++k) add(vw, 283, 64 * k, 64, 128);       // 35..38
    pa.bias[8] = h_params[NERFB_P_FEATURE_B];
    pa.bias[9] = h_params[NERFB_P_VIEWS_B];
    pa.src[0] = h_params[NERFB_P_ALPHA_W];
    pa.src[1] = h_params[NERFB_P_ALPHA_B];
    pa.src[2] = h_params[NERFB_P_RGB_W];
    pa.src[3] = h_params[NERFB_P_RGB_B];
  } else {
    NERFB_REQUIRE(out_ch >= 1 && out_ch <= 5 && h_params[NERFB_P_OUTPUT_W] && h_params[NERFB_P_OUTPUT_B],
                  "nerfb_pack_weights: output_linear missing or out_ch=%d unsupported", out_ch);
    pa.src[0] = h_params[NERFB_P_OUTPUT_W];
    pa.src[1] = h_params[NERFB_P_OUTPUT_B];
  }
  pa.nchunks = n;
  pa.bias_area = use_viewdirs ? kChunkAreaVD : kChunkAreaNoVD;
  dim3 grid(2, n + pa.nbias + 1);
  pack_weights_kernel<<<grid, 256, 0, as_stream(stream)>>>(pa, (uint8_t*)packed,
                                                          use_viewdirs ? kParamOffVD : kParamOffNoVD);
  return check_launch("pack_weights_kernel");
}

// Tensor maps over a packed weight blob viewed as [rows][128 B] uint8 (images are pre-swizzled, so SWIZZLE_NONE):
// box = 128 rows (one 16 KB chunk half) or 64 rows (views layer).  Cached per blob address.
struct BlobMaps { const void* ptr; size_t bytes; CUtensorMap tm128, tm64, tm32; };
static int get_blob_maps(const void* blob, size_t bytes, BlobMaps* out) {   // returns a COPY (the cache may evict)
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static BlobMaps cache[8];
  static int ncache = 0, victim = 0;
  static std::mutex mu;                       // the cache is shared by all host threads / streams
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < ncache; ++i)
    if (cache[i].ptr == blob && cache[i].bytes == bytes) { *out = cache[i]; return NERFB_OK; }
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NERFB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    NERFB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available");
    encode = (EncodeFn)fn;
  }
  BlobMaps m{};
  m.ptr = blob; m.bytes = bytes;
  const cuuint64_t dims[2] = {128, (cuuint64_t)(bytes / 128)};
  const cuuint64_t strides[1] = {128};
  const cuuint32_t estr[2] = {1, 1};
  for (int k = 0; k < 3; ++k) {
    const cuuint32_t box[2] = {128, k == 0 ? 128u : (k == 1 ? 64u : 32u)};      // 32 rows = one 4 KB bias-tile half
    const CUresult r = encode(k == 0 ? &m.tm128 : (k == 1 ? &m.tm64 : &m.tm32), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(blob), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    NERFB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for blob %p, %zu bytes", (int)r, blob, bytes);
  }
  int slot;
  if (ncache < 8) slot = ncache++;
  else { slot = victim; victim = (victim + 1) % 8; }
  cache[slot] = m;
  *out = m;
  return NERFB_OK;
}

constexpr int MODE_BWD_FUSED = 6;      // launch_mlp only: nerf_bwd_fused_kernel (chain CTAs + weight-gradient CTAs)

static int launch_mlp(int mode, MlpArgs& a, bool vd, void* stream, WArgs* wargs = nullptr, int chain_ctas = 0) {
  static Sched sched_vd, sched_novd, sched_bwd, sched_bwd_novd;
  static bool init[64] = {false};           // function attributes are per device
  static std::mutex mu;
  int dev = 0;
  NERFB_CUDA(cudaGetDevice(&dev));
  NERFB_REQUIRE(dev >= 0 && dev < 64, "launch_mlp: device ordinal %d out of range", dev);
  std::unique_lock<std::mutex> lock(mu);
  if (!init[dev]) {
    build_schedule(true, sched_vd);
    build_schedule(false, sched_novd);
    build_schedule_bwd(true, sched_bwd);
    build_schedule_bwd(false, sched_bwd_novd);
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_FWD_COMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_FWD_NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_FWD_COMP_NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_FWD_SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_FWD_SAVE_NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_mlp_kernel<MODE_BWD_NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    NERFB_CUDA(cudaFuncSetAttribute(nerf_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kSmemBytes > kWgSmemBytes ? kSmemBytes : kWgSmemBytes));
    init[dev] = true;
  }
  lock.unlock();
  a.nst = (int)((a.nrows + kSuperRows - 1) / kSuperRows);
  a.div_mul = 0u; a.div_shift = 0;
  if (a.S >= 2 && a.nrows < (1ll << 32)) {
    int l = 0;
    while ((1ll << l) < (long long)a.S) ++l;                                      // ceil(log2 S) >= 1
    a.div_mul = (uint32_t)((((1ull << l) - (unsigned long long)a.S) << 32) / (unsigned long long)a.S + 1ull);
    a.div_shift = l - 1;
  }
  a.nclusters = sm_count() / 2;
  if (const char* e = getenv("NERFB_MAX_CLUSTERS")) {      // experiments only (contention studies)
    const int m = atoi(e);
    if (m > 0 && m < a.nclusters) a.nclusters = m;
  }
  if (mode == MODE_BWD_FUSED) a.nclusters = chain_ctas / 2;
  if (a.nclusters > a.nst) a.nclusters = a.nst;
  a.watch = get_watch();
  a.trace = g_trace;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * a.nclusters);
  cudaLaunchAttribute coop[1];
  if (mode == MODE_BWD_FUSED) {
    NERFB_REQUIRE(wargs != nullptr && a.nclusters == chain_ctas / 2, "fused backward: too few samples for %d chain CTAs", chain_ctas);
    cfg.gridDim = dim3(sm_count());                       // chain CTAs + weight-gradient CTAs = every SM, all co-resident
    coop[0].id = cudaLaunchAttributeCooperative;          // ... which the cooperative launch verifies
    coop[0].val.cooperative = 1;
    cfg.attrs = coop;
    cfg.numAttrs = 1;
  }
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = as_stream(stream);
  {  // parameter block -> constant bank, in stream order; launches on different streams are chained by an event
    // (c_par, the event and the last stream are per device: every device has its own instance of the symbol)
    static cudaEvent_t last_done_d[64] = {nullptr};
    static cudaStream_t last_stream_d[64] = {nullptr};
    static std::mutex cmu;
    std::lock_guard<std::mutex> clock_(cmu);
    cudaEvent_t& last_done = last_done_d[dev];
    cudaStream_t& last_stream = last_stream_d[dev];
    const cudaStream_t st = as_stream(stream);
    if (last_done == nullptr) NERFB_CUDA(cudaEventCreateWithFlags(&last_done, cudaEventDisableTiming));
    else if (last_stream != st) NERFB_CUDA(cudaStreamWaitEvent(st, last_done, 0));
    NERFB_CUDA(cudaMemcpyToSymbolAsync(c_par, a.packed + a.param_off, kParamBytes, 0, cudaMemcpyDeviceToDevice, st));
    last_stream = st;
    g_const_event = last_done;
  }
  BlobMaps maps_v{};
  const BlobMaps* maps = &maps_v;
  const size_t blob_bytes = (size_t)a.param_off + kParamBytes;
  if (int rc = get_blob_maps(a.packed, blob_bytes, &maps_v)) return rc;
  if (mode == MODE_FWD) {
    if (vd) NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_FWD>, a, sched_vd, maps->tm128, maps->tm64, maps->tm32));
    else NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_FWD_NV>, a, sched_novd, maps->tm128, maps->tm64, maps->tm32));
    NERFB_CUDA(cudaEventRecord(g_const_event, cfg.stream));
    return check_launch("nerf_mlp_kernel<FWD>");
  } else if (mode == MODE_FWD_COMP) {
    if (vd) NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_FWD_COMP>, a, sched_vd, maps->tm128, maps->tm64, maps->tm32));
    else NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_FWD_COMP_NV>, a, sched_novd, maps->tm128, maps->tm64, maps->tm32));
    NERFB_CUDA(cudaEventRecord(g_const_event, cfg.stream));
    return check_launch("nerf_mlp_kernel<FWD_COMP>");
  } else if (mode == MODE_FWD_SAVE) {
    if (vd) NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_FWD_SAVE>, a, sched_vd, maps->tm128, maps->tm64, maps->tm32));
    else NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_FWD_SAVE_NV>, a, sched_novd, maps->tm128, maps->tm64, maps->tm32));
    NERFB_CUDA(cudaEventRecord(g_const_event, cfg.stream));
    return check_launch("nerf_mlp_kernel<FWD_SAVE>");
  }
  if (mode == MODE_BWD_FUSED) {
    cfg.dynamicSmemBytes = kSmemBytes > kWgSmemBytes ? kSmemBytes : kWgSmemBytes;
    NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_bwd_fused_kernel, a, sched_bwd, maps->tm128, maps->tm64, maps->tm32, *wargs, chain_ctas));
    NERFB_CUDA(cudaEventRecord(g_const_event, cfg.stream));
    return check_launch("nerf_bwd_fused_kernel");
  }
  if (vd) NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_BWD>, a, sched_bwd, maps->tm128, maps->tm64, maps->tm32));
  else NERFB_CUDA(cudaLaunchKernelEx(&cfg, nerf_mlp_kernel<MODE_BWD_NV>, a, sched_bwd_novd, maps->tm128, maps->tm64, maps->tm32));
  NERFB_CUDA(cudaEventRecord(g_const_event, cfg.stream));
  return check_launch("nerf_mlp_kernel<BWD>");
}

extern "C" int nerfb_mlp_fwd_bf16(int R, int S, const float* rays, int ray_width, const float* z, const float* pts,
                                  const float* viewdirs, int use_viewdirs, int out_ch, const void* packed, float* raw,
                                  void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && packed && raw, "nerfb_mlp_fwd_bf16: bad arguments");
  NERFB_REQUIRE(aligned16(packed) && aligned16(raw), "nerfb_mlp_fwd_bf16: packed / raw must be 16-byte aligned");
  NERFB_REQUIRE(use_viewdirs || (out_ch >= 1 && out_ch <= 5), "nerfb_mlp_fwd_bf16: out_ch=%d unsupported", out_ch);
  NERFB_REQUIRE(pts || (rays && z && ray_width >= 8), "nerfb_mlp_fwd_bf16: need pts or rays+z");
  NERFB_REQUIRE(!use_viewdirs || viewdirs || (rays && ray_width >= 11),
                "nerfb_mlp_fwd_bf16: view directions requested but not provided");
  if (R == 0) return NERFB_OK;
  MlpArgs a{};
  a.rays = rays; a.rw = ray_width; a.z = z; a.pts = pts; a.viewdirs = viewdirs;
  a.S = S; a.nrows = (long long)R * S; a.use_viewdirs = use_viewdirs;
  a.packed = (const uint8_t*)packed; a.raw = raw;
  a.out_ch = use_viewdirs ? 4 : out_ch;
  a.param_off = use_viewdirs ? kParamOffVD : kParamOffNoVD;
  a.bias_off = use_viewdirs ? kChunkAreaVD : kChunkAreaNoVD;
  return launch_mlp(MODE_FWD, a, use_viewdirs != 0, stream);
}

extern "C" int nerfb_mlp_fwd_composite_bf16(int R, int S, const float* rays, int ray_width, const float* z, int lindisp,
                                            int use_viewdirs, int out_ch, const void* packed, float* rec, float* wpart,
                                            void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && packed && rays && rec, "nerfb_mlp_fwd_composite_bf16: bad arguments");
  NERFB_REQUIRE(S % 32 == 0 && S <= 1024, "nerfb_mlp_fwd_composite_bf16: S=%d must be a multiple of 32, at most 1024", S);
  NERFB_REQUIRE((long long)R * S < 0x7fffffffll, "nerfb_mlp_fwd_composite_bf16: at most 2^31 samples per launch");
  NERFB_REQUIRE(ray_width >= (use_viewdirs ? 11 : 8), "nerfb_mlp_fwd_composite_bf16: ray_width=%d too small", ray_width);
  NERFB_REQUIRE(use_viewdirs || (out_ch >= 4 && out_ch <= 5), "nerfb_mlp_fwd_composite_bf16: out_ch=%d unsupported", out_ch);
  NERFB_REQUIRE(aligned16(packed) && aligned16(rec), "nerfb_mlp_fwd_composite_bf16: packed / rec must be 16-byte aligned");
  if (R == 0) return NERFB_OK;
  MlpArgs a{};
  a.rays = rays; a.rw = ray_width; a.z = z; a.pts = nullptr; a.viewdirs = nullptr;
  a.S = S; a.nrows = (long long)R * S; a.use_viewdirs = use_viewdirs;
  a.packed = (const uint8_t*)packed; a.raw = nullptr;
  a.out_ch = use_viewdirs ? 4 : out_ch;
  a.param_off = use_viewdirs ? kParamOffVD : kParamOffNoVD;
  a.bias_off = use_viewdirs ? kChunkAreaVD : kChunkAreaNoVD;
  a.lindisp = lindisp;
  a.zstep = S > 1 ? 1.0f / (float)(S - 1) : 0.f;
  a.comp_rec = rec; a.comp_wpart = wpart;
  return launch_mlp(MODE_FWD_COMP, a, use_viewdirs != 0, stream);
}

extern "C" size_t nerfb_train_store_bytes(long long nrows, int which) {
  const long long ntiles = 4 * ((nrows + kSuperRows - 1) / kSuperRows);
  return (size_t)ntiles * (size_t)(which == 0 ? kFwdTileBytes : (which == 1 ? kBwdTileBytes : kMaskTileBytes));
}

extern "C" int nerfb_mlp_fwd_save_bf16(int R, int S, const float* rays, int ray_width, const float* z,
                                       const float* pts, const float* viewdirs, const void* packed, float* raw,
                                       void* acts, void* masks, void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && packed && raw && acts && masks, "nerfb_mlp_fwd_save_bf16: bad arguments");
  NERFB_REQUIRE(aligned16(packed) && aligned16(raw) && aligned16(acts) && aligned16(masks),
                "nerfb_mlp_fwd_save_bf16: packed / raw / acts / masks must be 16-byte aligned");
  NERFB_REQUIRE(pts || (rays && z && ray_width >= 8), "nerfb_mlp_fwd_save_bf16: need pts or rays+z");
  NERFB_REQUIRE(viewdirs || (rays && ray_width >= 11), "nerfb_mlp_fwd_save_bf16: view directions not provided");
  if (R == 0) return NERFB_OK;
  MlpArgs a{};
  a.rays = rays; a.rw = ray_width; a.z = z; a.pts = pts; a.viewdirs = viewdirs;
  a.S = S; a.nrows = (long long)R * S; a.use_viewdirs = 1;
  a.packed = (const uint8_t*)packed; a.raw = raw; a.out_ch = 4;
  a.param_off = kParamOffVD;
  a.bias_off = kChunkAreaVD;
  a.acts = (uint8_t*)acts;
  a.masks = (uint8_t*)masks;
  return launch_mlp(MODE_FWD_SAVE, a, true, stream);
}

extern "C" size_t nerfb_packed_weights_bwd_bytes(void) { return (size_t)kChunkAreaBwd + kParamBytes; }

extern "C" int nerfb_pack_weights_bwd(const float* const* h_params, void* packed_bwd, void* stream) {
  NERFB_REQUIRE(h_params && packed_bwd, "nerfb_pack_weights_bwd: null argument");
  for (int i = 0; i < 16; ++i) NERFB_REQUIRE(h_params[i], "nerfb_pack_weights_bwd: pts_linears parameter %d missing", i);
  NERFB_REQUIRE(h_params[NERFB_P_VIEWS_W] && h_params[NERFB_P_VIEWS_B] && h_params[NERFB_P_FEATURE_W] &&
                    h_params[NERFB_P_FEATURE_B] && h_params[NERFB_P_ALPHA_W] && h_params[NERFB_P_ALPHA_B] &&
                    h_params[NERFB_P_RGB_W] && h_params[NERFB_P_RGB_B],
                "nerfb_pack_weights_bwd: view-dependent head parameters missing (the tensor-core backward "
                "supports use_viewdirs networks only)");
  PackArgsT pt{};
  int n = 0;
  for (int kc = 0; kc < 2; ++kc) pt.ch[n++] = PackChunkT{h_params[NERFB_P_VIEWS_W], 283, 64 * kc, 0};
  for (int kc = 0; kc < 4; ++kc) pt.ch[n++] = PackChunkT{h_params[NERFB_P_FEATURE_W], 256, 64 * kc, 0};
  for (int l = 7; l >= 1; --l)
    for (int kc = 0; kc < 4; ++kc) pt.ch[n++] = PackChunkT{h_params[2 * l], l == 5 ? 319 : 256, 64 * kc, l == 5 ? 63 : 0};
  PackArgs pa{};
  pa.use_viewdirs = 1;
  pa.out_ch = 4;
  for (int l = 0; l < 8; ++l) pa.bias[l] = h_params[2 * l + 1];
  pa.bias[8] = h_params[NERFB_P_FEATURE_B];
  pa.bias[9] = h_params[NERFB_P_VIEWS_B];
  pa.src[0] = h_params[NERFB_P_ALPHA_W];
  pa.src[1] = h_params[NERFB_P_ALPHA_B];
  pa.src[2] = h_params[NERFB_P_RGB_W];
  pa.src[3] = h_params[NERFB_P_RGB_B];
  dim3 grid(2, kNumBwdChunks + 1);
  pack_weights_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(pt, pa, (uint8_t*)packed_bwd);
  return check_launch("pack_weights_bwd_kernel");
}

extern "C" int nerfb_mlp_bwd_chain_bf16(long long nrows, const float* d_raw, const void* packed_bwd, const void* masks,
                                        void* gstore, void* stream) {
  NERFB_REQUIRE(nrows >= 0 && d_raw && packed_bwd && masks && gstore, "nerfb_mlp_bwd_chain_bf16: bad arguments");
  NERFB_REQUIRE(aligned16(d_raw) && aligned16(packed_bwd) && aligned16(masks) && aligned16(gstore),
                "nerfb_mlp_bwd_chain_bf16: all buffers must be 16-byte aligned");
  if (nrows == 0) return NERFB_OK;
  MlpArgs a{};
  a.nrows = nrows; a.S = 1; a.use_viewdirs = 1; a.out_ch = 4;
  a.packed = (const uint8_t*)packed_bwd;
  a.param_off = kChunkAreaBwd;
  a.masks = (uint8_t*)masks;
  a.gstore = (uint8_t*)gstore;
  a.d_raw = d_raw;
  return launch_mlp(MODE_BWD, a, true, stream);
}

extern "C" int nerfb_mlp_bwd_fused_bf16(long long nrows, const float* d_raw_padded, const void* packed_bwd, const void* masks,
                                        const void* acts, void* gstore, float* const* h_grads, int* flags, int epoch,
                                        int chain_sms, void* stream) {
  NERFB_REQUIRE(nrows > 0 && d_raw_padded && packed_bwd && masks && acts && gstore && h_grads && flags,
                "nerfb_mlp_bwd_fused_bf16: bad arguments");
  NERFB_REQUIRE(aligned16(d_raw_padded) && aligned16(packed_bwd) && aligned16(masks) && aligned16(gstore) && aligned16(acts),
                "nerfb_mlp_bwd_fused_bf16: all buffers must be 16-byte aligned");
  const int sms = sm_count();
  NERFB_REQUIRE(chain_sms >= 2 && chain_sms % 2 == 0 && sms % 2 == 0 && sms - chain_sms >= kMaxJobs,
                "nerfb_mlp_bwd_fused_bf16: chain_sms=%d does not split %d SMs", chain_sms, sms);
  MlpArgs a{};
  a.nrows = nrows; a.S = 1; a.use_viewdirs = 1; a.out_ch = 4;
  a.packed = (const uint8_t*)packed_bwd;
  a.param_off = kChunkAreaBwd;
  a.masks = (uint8_t*)masks;
  a.gstore = (uint8_t*)gstore;
  a.d_raw = d_raw_padded;
  a.flags = flags;
  a.epoch = epoch;
  WArgs w;
  if (int rc = build_wgrad_args("nerfb_mlp_bwd_fused_bf16", nrows, 1, 4, acts, gstore, d_raw_padded, nullptr, h_grads,
                                sms - chain_sms, w))
    return rc;
  w.flags = flags;
  w.epoch = epoch;
  return launch_mlp(MODE_BWD_FUSED, a, true, stream, &w, chain_sms);
}

// ---- the same three steps for the network WITHOUT view directions (output_linear head, nerf.py:50-55, 108-110)

extern "C" int nerfb_mlp_fwd_save_nv_bf16(int R, int S, const float* rays, int ray_width, const float* z,
                                          const float* pts, int out_ch, const void* packed, float* raw, void* acts,
                                          void* masks, void* stream) {
  NERFB_REQUIRE(R >= 0 && S > 0 && packed && raw && acts && masks, "nerfb_mlp_fwd_save_nv_bf16: bad arguments");
  NERFB_REQUIRE(aligned16(packed) && aligned16(raw) && aligned16(acts) && aligned16(masks),
                "nerfb_mlp_fwd_save_nv_bf16: packed / raw / acts / masks must be 16-byte aligned");
  NERFB_REQUIRE(out_ch >= 1 && out_ch <= 5, "nerfb_mlp_fwd_save_nv_bf16: out_ch=%d unsupported", out_ch);
  NERFB_REQUIRE(pts || (rays && z && ray_width >= 8), "nerfb_mlp_fwd_save_nv_bf16: need pts or rays+z");
  if (R == 0) return NERFB_OK;
  MlpArgs a{};
  a.rays = rays; a.rw = ray_width; a.z = z; a.pts = pts; a.viewdirs = nullptr;
  a.S = S; a.nrows = (long long)R * S; a.use_viewdirs = 0;
  a.packed = (const uint8_t*)packed; a.raw = raw; a.out_ch = out_ch;
  a.param_off = kParamOffNoVD;
  a.bias_off = kChunkAreaNoVD;
  a.acts = (uint8_t*)acts;
  a.masks = (uint8_t*)masks;
  return launch_mlp(MODE_FWD_SAVE, a, false, stream);
}

extern "C" int nerfb_pack_weights_bwd_nv(const float* const* h_params, int out_ch, void* packed_bwd, void* stream) {
  NERFB_REQUIRE(h_params && packed_bwd, "nerfb_pack_weights_bwd_nv: null argument");
  for (int i = 0; i < 16; ++i) NERFB_REQUIRE(h_params[i], "nerfb_pack_weights_bwd_nv: pts_linears parameter %d missing", i);
  NERFB_REQUIRE(out_ch >= 1 && out_ch <= 5 && h_params[NERFB_P_OUTPUT_W] && h_params[NERFB_P_OUTPUT_B],
                "nerfb_pack_weights_bwd_nv: output_linear missing or out_ch=%d unsupported", out_ch);
  PackArgsT pt{};
  int n = 6;                                       // chunks 0..5 (views^T, feature^T) stay unused
  for (int l = 7; l >= 1; --l)
    for (int kc = 0; kc < 4; ++kc) pt.ch[n++] = PackChunkT{h_params[2 * l], l == 5 ? 319 : 256, 64 * kc, l == 5 ? 63 : 0};
  PackArgs pa{};
  pa.use_viewdirs = 0;
  pa.out_ch = out_ch;                              // head block: output_linear rows 0..out_ch-1, zero above
  for (int l = 0; l < 8; ++l) pa.bias[l] = h_params[2 * l + 1];
  pa.src[0] = h_params[NERFB_P_OUTPUT_W];
  pa.src[1] = h_params[NERFB_P_OUTPUT_B];
  dim3 grid(2, kNumBwdChunks + 1);
  pack_weights_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(pt, pa, (uint8_t*)packed_bwd);
  return check_launch("pack_weights_bwd_kernel");
}

extern "C" int nerfb_mlp_bwd_chain_nv_bf16(long long nrows, const float* d_raw, const float* d_raw4, const void* packed_bwd,
                                           const void* masks, void* gstore, void* stream) {
  NERFB_REQUIRE(nrows >= 0 && d_raw && packed_bwd && masks && gstore, "nerfb_mlp_bwd_chain_nv_bf16: bad arguments");
  NERFB_REQUIRE(aligned16(d_raw) && aligned16(packed_bwd) && aligned16(masks) && aligned16(gstore),
                "nerfb_mlp_bwd_chain_nv_bf16: all buffers must be 16-byte aligned");
  if (nrows == 0) return NERFB_OK;
  MlpArgs a{};
  a.nrows = nrows; a.S = 1; a.use_viewdirs = 0; a.out_ch = 4;
  a.packed = (const uint8_t*)packed_bwd;
  a.param_off = kChunkAreaBwd;
  a.masks = (uint8_t*)masks;
  a.gstore = (uint8_t*)gstore;
  a.d_raw = d_raw;
  a.d_raw4 = d_raw4;
  return launch_mlp(MODE_BWD, a, false, stream);
}

/* Debug / test hook: the static schedule of a kernel mode (0 forward with view directions, 1 forward without,
 * 2 data-gradient chain, 3 data-gradient chain without view directions) as flat int arrays -- loads: {blob offset, bytes per CTA}; ops: {load, tile, asrc, flags,
 * block length (non-zero on the first op of a block)}.  Host only; lets the CPU test-suite check the ring invariants. */
extern "C" int nerfb_debug_schedule(int which, int* h_nloads, int* h_nops, int* h_nlayers, int* h_loads, int* h_ops,
                                    int* h_ring_slots) {
  NERFB_REQUIRE(which >= 0 && which <= 3 && h_nloads && h_nops && h_nlayers && h_loads && h_ops && h_ring_slots,
                "nerfb_debug_schedule: bad arguments");
  static Sched s;
  if (which >= 2) build_schedule_bwd(which == 2, s); else build_schedule(which == 0, s);
  *h_nloads = s.nloads; *h_nops = s.nops; *h_nlayers = s.nlayers; *h_ring_slots = kNS;
  for (int i = 0; i < s.nloads; ++i) { h_loads[2 * i] = (int)s.loads[i].off; h_loads[2 * i + 1] = (int)(s.loads[i].bytes & ~63u); }
  for (int i = 0; i < s.nops; ++i) {
    h_ops[5 * i + 0] = s.ops[i].load; h_ops[5 * i + 1] = s.ops[i].tile & 0xF; h_ops[5 * i + 2] = s.ops[i].asrc;
    h_ops[5 * i + 3] = s.ops[i].flags; h_ops[5 * i + 4] = s.ops[i].tile >> 4;
  }
  return NERFB_OK;
}
