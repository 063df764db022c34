// Weight-gradient kernel body shared by the stand-alone kernel (wgrad_tc.cu) and the fused backward launch
// (mlp_tc.cu: nerf_bwd_fused_kernel).  See wgrad_tc.cu for the design.
#pragma once
#include <cstdlib>
#include "common.cuh"
#include "tc05.cuh"
#include "mlp_layout.cuh"

namespace nerfb {
using namespace tc05;

tc05::Watch get_watch();  // umma_selftest.cu

// completion flags of the fused backward launch: per 128-sample tile, kFlagGroups gradient panel groups (d hv, d feature,
// dY_0..dY_7) x 2 column halves; a flag holds the launch's epoch once the chain's bulk stores of that half have completed
constexpr int kFlagGroups = 10;

constexpr int kWgThreads = 192;             // warp 0 loader, warp 1 mma, warps 2..5 simt + epilogue
constexpr int kStageRows = 64;              // samples per stage
constexpr int kHalfPanel = kPanel / 2;      // 64 rows x 128 B
constexpr int kMaxStagePanels = 9;
constexpr int kStageDraw = kMaxStagePanels * kHalfPanel;   // d_raw slice [64][4] fp32
constexpr int kStageStride = kStageDraw + 1024;            // 74752, multiple of 1024
constexpr int kWgStages = 3;
constexpr int kWgSmBars = kWgStages * kStageStride;
constexpr int kWgSmDrain = kWgSmBars + 128;           // TMEM drain staging: 4 warps x [32 rows][8 columns] fp32
constexpr int kWgSmemBytes = kWgSmDrain + 4 * 1024;
static_assert(kWgSmemBytes <= 232448, "shared memory budget exceeded");
static_assert(kStageStride % 1024 == 0, "stage buffers must keep the swizzle phase");

enum { SIMT_NONE = 0, SIMT_ALPHA = 1, SIMT_RGB = 2, SIMT_OUT = 3 };

struct WJob {
  uint32_t a_off;          // gradient panels inside a bwd tile (gstore)
  int a_panels;            // 4 (256 outputs: two M halves) or 2 (128 outputs)
  uint32_t b_off[2];       // activation panel groups inside a fwd tile (acts)
  int b_panels[2];         // 4, 1 or 0
  uint32_t x_off;          // extra activation panels staged for the SIMT reduction (hv)
  int x_panels;
  float* dW; int ld;       // destination [outputs, ld]
  int col_off[2], ncols[2];// group g lands at dW[:, col_off[g] : col_off[g]+ncols[g]]
  float* dbias;            // += colsum(dY) or null
  int simt;                // SIMT_*
  float* head_w;           // alpha_linear.weight [1,256] / rgb_linear.weight [3,128] gradient
  float* head_b;
  int weight;              // relative cost of one stage (bytes / 1024)
  int flag_grp;            // fused launch: index of the job's dY panel group (0 = d hv, 1 = d feature, 2 + l = dY_l)
};
constexpr int kMaxJobs = 12;
struct WArgs {
  WJob jobs[kMaxJobs];
  int njobs;
  const uint8_t* acts;
  const uint8_t* gstore;
  const float* d_raw;      // [ntiles*128, 4], zero beyond the real rows
  const float* d_raw4;     // SIMT_OUT with out_ch = 5: gradient of output channel 4, [ntiles*128] (zero padded) or null
  int out_ch;              // SIMT_OUT: rows of output_linear
  long long nstages;       // 2 per 128-sample tile
  int cta_begin[kMaxJobs + 1];   // fused launch: CTAs [cta_begin[j], cta_begin[j+1]) take job j's tiles round-robin
  // stand-alone kernel: the jobs laid end to end on a COST timeline (job j = nstages * weight_j units starting at
  // cost_begin[j]); CTA c of ncta takes the stages whose start falls into [c, c + 1) * cost_total / ncta -- one work item
  // for most CTAs, two (the tail of one job and the head of the next) for the ~njobs CTAs that straddle a job boundary
  long long cost_begin[kMaxJobs + 1];
  int ncta;
  Watch watch;
  const int* flags;        // fused launch: [ntiles][kFlagGroups][2]
  int epoch;
};

// Body of the weight-gradient kernel.  FUSED = false: the stand-alone kernel (nerf_wgrad_kernel): the CTAs of a job split
// its stages into contiguous ranges.  FUSED = true: the CTAs run beside the data-gradient chain inside ONE launch
// (nerf_bwd_fused_kernel, mlp_tc.cu) and consume its output gradients while they are still in L2: the CTAs of a job take
// the 128-sample tiles round-robin (tile = isplit, isplit + nsplit, ...: the order in which the chain produces them) and
// the loader waits for the chain's per-tile completion flags before it issues the bulk loads of a tile's dY panels.
// `vblock` = index of this CTA among the weight-gradient CTAs; blockDim may exceed kWgThreads (extra warps idle).
template <bool FUSED>
__device__ __forceinline__ void wgrad_body(const WArgs& a, const int vblock) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(smem + kWgSmBars + 64);   // (dynamic: the fused launch has no room for static shared memory)
  auto bar = [&](int i) -> uint32_t { return sbase + kWgSmBars + 8 * i; };
  const int B_FULL = 0, B_EMPTY = kWgStages, B_ACC = 2 * kWgStages, B_TFREE = 2 * kWgStages + 1;

  if ((sbase & 1023u) != 0) watchdog_fire(a.watch, 0xB11, sbase, 0);
  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(bar(B_FULL + s), 1);
      mbar_init(bar(B_EMPTY + s), 1 + 4);     // tcgen05.commit + one arrive per SIMT warp
    }
    mbar_init(bar(B_ACC), 1);
    mbar_init(bar(B_TFREE), 4);               // the four draining warps, between two work items of a CTA
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<1>(smem_u32(&tmem_slot), 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // ---- my work items: (job, stages).  Stand-alone kernel: my slice of the cost timeline (1 or 2 items); fused launch: ONE
  // item, job jmine's tiles isplit, isplit + nsplit, ... (2 stages = the two 64-row halves of a tile)
  int item_job[2] = {0, 0};
  long long item_s0[2] = {0, 0}, item_n[2] = {0, 0};
  int nitems = 0;
  int nsplit = 1, isplit = 0;
  if (FUSED) {
    int jmine = 0;
    while (jmine + 1 < a.njobs && vblock >= a.cta_begin[jmine + 1]) ++jmine;
    nsplit = a.cta_begin[jmine + 1] - a.cta_begin[jmine];
    isplit = vblock - a.cta_begin[jmine];
    const long long ntiles = a.nstages >> 1;
    item_job[0] = jmine;
    item_n[0] = isplit < ntiles ? 2 * ((ntiles - 1 - isplit) / nsplit + 1) : 0;
    nitems = item_n[0] > 0 ? 1 : 0;
  } else if (a.ncta == 0) {
    // small launches: the CTAs of a job split its stages evenly, ONE item (one TMEM drain) per CTA
    int jmine = 0;
    while (jmine + 1 < a.njobs && vblock >= a.cta_begin[jmine + 1]) ++jmine;
    const int ns = a.cta_begin[jmine + 1] - a.cta_begin[jmine], is = vblock - a.cta_begin[jmine];
    const long long s0 = a.nstages * is / ns, s1 = a.nstages * (is + 1) / ns;
    item_job[0] = jmine; item_s0[0] = s0; item_n[0] = s1 - s0;
    nitems = s1 > s0 ? 1 : 0;
  } else {
    const long long T = a.cost_begin[a.njobs];
    const long long lo = T * vblock / a.ncta, hi = T * (vblock + 1) / a.ncta;
    for (int j = 0; j < a.njobs && nitems < 2; ++j) {
      const long long c0 = a.cost_begin[j], wj = a.jobs[j].weight;
      long long b0 = lo <= c0 ? 0 : (lo - c0 + wj - 1) / wj, b1 = hi <= c0 ? 0 : (hi - c0 + wj - 1) / wj;   // first stage starting at or after lo / hi
      if (b0 > a.nstages) b0 = a.nstages;
      if (b1 > a.nstages) b1 = a.nstages;
      if (b1 > b0) { item_job[nitems] = j; item_s0[nitems] = b0; item_n[nitems] = b1 - b0; ++nitems; }
    }
  }
  // global stage index of stage g of item it
  auto stage_of = [&](int it, long long g) -> long long {
    return FUSED ? 2 * ((long long)isplit + (g >> 1) * nsplit) + (g & 1) : item_s0[it] + g;
  };
  const bool has_work = nitems > 0;

  if (warp == 0) {
    // ===================== loader (gg = position in the stage ring, running across my items)
    if (lane == 0 && has_work) {
      long long gg = 0;
      for (int it = 0; it < nitems; ++it) {
        const int jmine = item_job[it];
        const WJob& jb = a.jobs[jmine];
        const int npan = jb.a_panels + jb.b_panels[0] + jb.b_panels[1] + jb.x_panels;
        for (long long g = 0; g < item_n[it]; ++g, ++gg) {
          const long long s = stage_of(it, g);
          const int slot = (int)(gg % kWgStages);
          if (FUSED && (g & 1) == 0 && jb.a_panels > 0) {
            // the chain's bulk stores of this tile's dY panels (both column halves) have completed: acquire the flags,
            // then order the async-proxy reads below after them
            const int* f = a.flags + ((size_t)(s >> 1) * kFlagGroups + jb.flag_grp) * 2;
            const long long t0 = clock64();
            for (int h = 0; h < 2; ++h) {
              int v;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f + h) : "memory");
                if (v != a.epoch && clock64() - t0 > kWatchdogCycles)
                  watchdog_fire(a.watch, 0x1200 + jmine, (uint32_t)(s >> 1), (uint32_t)v);
              } while (v != a.epoch);
            }
            asm volatile("fence.proxy.async;" ::: "memory");
          }
          if (gg >= kWgStages) mbar_wait(bar(B_EMPTY + slot), (uint32_t)(((gg / kWgStages) + 1) & 1), a.watch, 0x1100 + jmine);
          const uint32_t dst = sbase + slot * kStageStride;
          const uint32_t fb = bar(B_FULL + slot);
          mbar_arrive_expect_tx(fb, (uint32_t)npan * kHalfPanel + 1024u +
                                        ((jb.simt == SIMT_OUT && a.d_raw4 != nullptr) ? (uint32_t)kStageRows * 4u : 0u));
          const size_t tile = (size_t)(s >> 1), half = (size_t)(s & 1) * kHalfPanel;
          const uint8_t* gt = a.gstore + tile * kBwdTileBytes + jb.a_off + half;
          const uint8_t* ft = a.acts + tile * kFwdTileBytes + half;
          int p = 0;
          for (int i = 0; i < jb.a_panels; ++i, ++p) bulk_g2s(dst + p * kHalfPanel, gt + (size_t)i * kPanel, kHalfPanel, fb);
          for (int grp = 0; grp < 2; ++grp)
            for (int i = 0; i < jb.b_panels[grp]; ++i, ++p)
              bulk_g2s(dst + p * kHalfPanel, ft + jb.b_off[grp] + (size_t)i * kPanel, kHalfPanel, fb);
          for (int i = 0; i < jb.x_panels; ++i, ++p)
            bulk_g2s(dst + p * kHalfPanel, ft + jb.x_off + (size_t)i * kPanel, kHalfPanel, fb);
          bulk_g2s(dst + kStageDraw, a.d_raw + (size_t)s * kStageRows * 4, 1024, fb);
          if (jb.simt == SIMT_OUT && a.d_raw4 != nullptr)      // channel 4 slice [64] fp32 behind the 4 staged h7 panels
            bulk_g2s(dst + 4 * kHalfPanel, a.d_raw4 + (size_t)s * kStageRows, kStageRows * 4, fb);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane)
    if (has_work && elect_one()) {
      long long gg = 0;
      for (int it = 0; it < nitems; ++it) {
        const int jmine = item_job[it];
        const WJob& jb = a.jobs[jmine];
        const int nmh = jb.a_panels / 2;
        const uint32_t idesc0 = make_idesc_bf16_mn(128, 64 * jb.b_panels[0]);
        const uint32_t idesc1 = make_idesc_bf16_mn(128, jb.b_panels[1] ? 64 * jb.b_panels[1] : 64);
        if (it > 0) {     // the previous item's accumulators have been drained out of TMEM
          mbar_wait(bar(B_TFREE), (uint32_t)((it - 1) & 1), a.watch, 0x1600 + jmine);
          tc_fence_after();
        }
        for (long long g = 0; g < item_n[it]; ++g, ++gg) {
          const int slot = (int)(gg % kWgStages);
          mbar_wait(bar(B_FULL + slot), (uint32_t)((gg / kWgStages) & 1), a.watch, 0x1300 + jmine);
          tc_fence_after();
          const uint32_t st = sbase + slot * kStageStride;
          for (int mh = 0; mh < nmh; ++mh) {
            const uint64_t da = make_sdesc_sw128_mn(st + mh * 2 * kHalfPanel, kHalfPanel);
            const uint64_t db0 = make_sdesc_sw128_mn(st + jb.a_panels * kHalfPanel, kHalfPanel);
            const uint32_t acc0 = (g > 0) ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < kStageRows / 16; ++k)
              umma_bf16<1>(tmem_base + mh * 256, da + (uint64_t)(k * 128), db0 + (uint64_t)(k * 128), idesc0,
                           (k > 0) ? 1u : acc0);
            if (jb.b_panels[1]) {
              const uint64_t db1 = make_sdesc_sw128_mn(st + (jb.a_panels + jb.b_panels[0]) * kHalfPanel, kHalfPanel);
              const uint32_t d1 = tmem_base + mh * 256 + 64 * jb.b_panels[0];   // only used with one M half (views)
#pragma unroll
              for (int k = 0; k < kStageRows / 16; ++k)
                umma_bf16<1>(d1, da + (uint64_t)(k * 128), db1 + (uint64_t)(k * 128), idesc1, (k > 0) ? 1u : acc0);
            }
          }
          umma_commit<1>(bar(B_EMPTY + slot));
          if (g + 1 == item_n[it]) umma_commit<1>(bar(B_ACC));
        }
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===================== SIMT reductions + epilogue (warps 2..5, TMEM lane quarter = warp % 4)
    const int st_id = tid - 64;                 // 0..127
    const int q = warp & 3;
    const int row_in_half = 32 * q + lane;      // accumulator lane = output feature within the M half
    long long gg = 0;
    for (int it = 0; it < nitems; ++it) {
      const int j = item_job[it];
      const WJob& jb = a.jobs[j];
      // per-thread SIMT accumulators.  Mapping for 4-panel operands: 16-byte unit u = st_id & 7 (8 columns)
      // of panel (st_id >> 3) & 3, rows r = (st_id >> 5) mod 4; a warp reads 4 full 128-byte rows per LDS.128.
      float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // bias: 8 columns of dY
      float hs[5][8];                                              // head weight gradients
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) hs[k][i] = 0.f;
      float hb[5] = {0.f, 0.f, 0.f, 0.f, 0.f};                     // sums of d_raw components (thread 0)
      const int su = st_id & 7;
      // dY has 4 panels (rows r = rg4 mod 4) or 2 panels (rows r = rg8 mod 8)
      const int bpan = jb.a_panels == 4 ? (st_id >> 3) & 3 : (st_id >> 3) & 1;
      const int brow0 = jb.a_panels == 4 ? (st_id >> 5) : (st_id >> 4);
      const int bstep = jb.a_panels == 4 ? 4 : 8;
      const bool do_bias = jb.dbias != nullptr;
      auto acc8 = [](float (&acc)[8], const uint4 v, const float w) {
        acc[0] = fmaf(w, __uint_as_float(v.x << 16), acc[0]); acc[1] = fmaf(w, __uint_as_float(v.x & 0xFFFF0000u), acc[1]);
        acc[2] = fmaf(w, __uint_as_float(v.y << 16), acc[2]); acc[3] = fmaf(w, __uint_as_float(v.y & 0xFFFF0000u), acc[3]);
        acc[4] = fmaf(w, __uint_as_float(v.z << 16), acc[4]); acc[5] = fmaf(w, __uint_as_float(v.z & 0xFFFF0000u), acc[5]);
        acc[6] = fmaf(w, __uint_as_float(v.w << 16), acc[6]); acc[7] = fmaf(w, __uint_as_float(v.w & 0xFFFF0000u), acc[7]);
      };
      for (long long g = 0; g < item_n[it]; ++g, ++gg) {
        const int slot = (int)(gg % kWgStages);
        mbar_wait(bar(B_FULL + slot), (uint32_t)((gg / kWgStages) & 1), a.watch, 0x1400 + j);
        const uint8_t* stp = smem + slot * kStageStride;
        const float4* draw = reinterpret_cast<const float4*>(stp + kStageDraw);
        if (do_bias) {
          const uint8_t* pa = stp + bpan * kHalfPanel;
#pragma unroll 4
          for (int r = brow0; r < kStageRows; r += bstep) {
            const uint4 v = *reinterpret_cast<const uint4*>(pa + r * 128 + ((su ^ (r & 7)) << 4));
            bs[0] += __uint_as_float(v.x << 16); bs[1] += __uint_as_float(v.x & 0xFFFF0000u);
            bs[2] += __uint_as_float(v.y << 16); bs[3] += __uint_as_float(v.y & 0xFFFF0000u);
            bs[4] += __uint_as_float(v.z << 16); bs[5] += __uint_as_float(v.z & 0xFFFF0000u);
            bs[6] += __uint_as_float(v.w << 16); bs[7] += __uint_as_float(v.w & 0xFFFF0000u);
          }
        }
        if (jb.simt == SIMT_ALPHA) {
          // d alpha_linear.weight[c] += sum_r dsigma[r] * h7[r][c]   (h7 = B group 0, 4 panels)
          const uint8_t* pb = stp + (jb.a_panels + ((st_id >> 3) & 3)) * kHalfPanel;
#pragma unroll 4
          for (int r = st_id >> 5; r < kStageRows; r += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(pb + r * 128 + ((su ^ (r & 7)) << 4));
            acc8(hs[0], v, draw[r].w);
          }
          if (st_id == 0)
            for (int r = 0; r < kStageRows; ++r) hb[3] += draw[r].w;
        } else if (jb.simt == SIMT_RGB) {
          // d rgb_linear.weight[k][c] += sum_r d_rgb[r][k] * hv[r][c]   (hv = extra panels, 2 x 64 columns)
          const uint8_t* px = stp + (jb.a_panels + jb.b_panels[0] + jb.b_panels[1] + ((st_id >> 3) & 1)) * kHalfPanel;
#pragma unroll 2
          for (int r = st_id >> 4; r < kStageRows; r += 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(px + r * 128 + ((su ^ (r & 7)) << 4));
            const float4 d = draw[r];
            acc8(hs[0], v, d.x); acc8(hs[1], v, d.y); acc8(hs[2], v, d.z);
          }
          if (st_id == 0)
            for (int r = 0; r < kStageRows; ++r) { hb[0] += draw[r].x; hb[1] += draw[r].y; hb[2] += draw[r].z; }
        } else if (jb.simt == SIMT_OUT) {
          // d output_linear.weight[k][c] += sum_r d_raw[r][k] * h7[r][c]   (h7 = the 4 staged panels; no MMA in this job)
          const uint8_t* px = stp + ((st_id >> 3) & 3) * kHalfPanel;
          const float* d4 = a.d_raw4 != nullptr ? reinterpret_cast<const float*>(stp + 4 * kHalfPanel) : nullptr;
#pragma unroll 2
          for (int r = st_id >> 5; r < kStageRows; r += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(px + r * 128 + ((su ^ (r & 7)) << 4));
            const float4 d = draw[r];
            acc8(hs[0], v, d.x); acc8(hs[1], v, d.y); acc8(hs[2], v, d.z); acc8(hs[3], v, d.w);
            if (d4 != nullptr) acc8(hs[4], v, d4[r]);
          }
          if (st_id == 0)
            for (int r = 0; r < kStageRows; ++r) {
              hb[0] += draw[r].x; hb[1] += draw[r].y; hb[2] += draw[r].z; hb[3] += draw[r].w;
              if (d4 != nullptr) hb[4] += d4[r];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_EMPTY + slot));
      }
      // ---- flush the SIMT sums (several row groups add into the same columns)
      if (do_bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(jb.dbias + 64 * bpan + 8 * su + i, bs[i]);
      }
      if (jb.simt == SIMT_ALPHA) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(jb.head_w + 64 * ((st_id >> 3) & 3) + 8 * su + i, hs[0][i]);
        if (st_id == 0) atomicAdd(jb.head_b, hb[3]);
      } else if (jb.simt == SIMT_RGB) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int i = 0; i < 8; ++i) atomicAdd(jb.head_w + k * 128 + 64 * ((st_id >> 3) & 1) + 8 * su + i, hs[k][i]);
        if (st_id == 0)
          for (int k = 0; k < 3; ++k) atomicAdd(jb.head_b + k, hb[k]);
      } else if (jb.simt == SIMT_OUT) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          if (k < a.out_ch) {
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(jb.head_w + k * 256 + 64 * ((st_id >> 3) & 3) + 8 * su + i, hs[k][i]);
            if (st_id == 0) atomicAdd(jb.head_b + k, hb[k]);
          }
        }
      }
      // ---- drain the accumulators: dW[out, col_off + c] += D[out, c]
      mbar_wait(bar(B_ACC), (uint32_t)(it & 1), a.watch, 0x1500 + j);
      tc_fence_after();
      // A thread holds one accumulator ROW (output feature), so reductions straight from the registers hit 32 different
      // rows per warp instruction -- 32 L2 sector operations carrying one float (or one float4) each; with all CTAs
      // draining at once that was 0.11-0.15 ms at the end of every launch (what-if without the drain: 0.54 -> 0.43 ms on
      // 262 k samples).  Instead every 32 x 8 block goes through a 1 KB shared-memory tile and is added with 8 lanes per
      // row: 4 rows x 32 contiguous bytes per instruction, whole sectors.
      const int nmh = jb.a_panels / 2;
      float* stg = reinterpret_cast<float*>(smem + kWgSmDrain) + q * 256;
      const int scol = lane & 7, srow = lane >> 3;
      for (int mh = 0; mh < nmh; ++mh) {
        int cbase = 0;
        for (int grp = 0; grp < 2; ++grp) {
          const int ncol_g = 64 * jb.b_panels[grp];
          float* dst = jb.dW + (size_t)(128 * mh + 32 * q) * jb.ld + jb.col_off[grp];     // first row of my lane quarter
          for (int c0 = 0; c0 < ncol_g; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + mh * 256 + cbase + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
              *reinterpret_cast<uint4*>(stg + lane * 8) = make_uint4(v[8 * sub], v[8 * sub + 1], v[8 * sub + 2], v[8 * sub + 3]);
              *reinterpret_cast<uint4*>(stg + lane * 8 + 4) =
                  make_uint4(v[8 * sub + 4], v[8 * sub + 5], v[8 * sub + 6], v[8 * sub + 7]);
              __syncwarp();
              const int cc = c0 + 8 * sub + scol;
              if (cc < jb.ncols[grp]) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const int rr = srow + 4 * k;
                  atomicAdd(dst + (size_t)rr * jb.ld + cc, stg[rr * 8 + scol]);
                }
              }
              __syncwarp();
            }
          }
          cbase += ncol_g;
        }
      }
      if (it + 1 < nitems) {      // hand TMEM back to the MMA issuer for my next item
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_TFREE));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<1>(tmem_base, 512);
}


// fills `a` (jobs, tile stores, stage count, CTA ranges for `grid` weight-gradient CTAs); returns an error code
inline int build_wgrad_args(const char* who, long long nrows, int use_viewdirs, int out_ch, const void* acts, const void* gstore,
                            const float* d_raw_padded, const float* d_raw4_padded, float* const* h_grads, int grid, WArgs& a) {
  NERFB_REQUIRE(nrows >= 0 && acts && gstore && d_raw_padded && h_grads, "%s: bad arguments", who);
  NERFB_REQUIRE(aligned16(acts) && aligned16(gstore) && aligned16(d_raw_padded) && aligned16(d_raw4_padded),
                "%s: acts / gstore / d_raw_padded must be 16-byte aligned", who);
  for (int i = 0; i < NERFB_NUM_PARAMS; ++i) {
    const bool needed = i < 16 || (use_viewdirs ? i < NERFB_P_OUTPUT_W : i >= NERFB_P_OUTPUT_W);
    if (!needed) continue;
    NERFB_REQUIRE(h_grads[i], "%s: gradient tensor %d missing", who, i);
    NERFB_REQUIRE(aligned16(h_grads[i]), "%s: gradient tensor %d is not 16-byte aligned", who, i);
  }
  a = WArgs{};
  int n = 0;
  auto G = [&](int i) { return h_grads[i]; };
  // relative cost of one 64-sample stage of each job kind (bytes/1024 for the HBM-bound plain layers; the narrow
  // gamma(x) jobs move 41 KB but pay the same CUDA-core bias reduction, so they weigh more than their bytes)
  int narrow_w = 50, views_w = 9 * 8 + 1;   // (measured sweep: profiles/README.md)
  if (const char* e = getenv("NERFB_WG_NARROW_W")) narrow_w = atoi(e);
  if (const char* e = getenv("NERFB_WG_VIEWS_W")) views_w = atoi(e);
  int out_w = 110;                          // output_linear job: 33 KB per stage but 5 CUDA-core FMAs per staged element (sweep: profiles/README.md)
  if (const char* e = getenv("NERFB_WG_OUT_W")) out_w = atoi(e);
  auto plain = [&](uint32_t a_off, uint32_t b_off, float* dW, int ld, int col_off, float* dbias) {
    WJob j{};
    j.a_off = a_off; j.a_panels = 4;
    j.b_off[0] = b_off; j.b_panels[0] = 4;
    j.dW = dW; j.ld = ld; j.col_off[0] = col_off; j.ncols[0] = 256;
    j.dbias = dbias;
    j.weight = 8 * 8 + 1;
    return j;
  };
  auto narrow = [&](uint32_t a_off, float* dW, int ld, float* dbias) {     // gamma(x) block: 63 valid columns
    WJob j{};
    j.a_off = a_off; j.a_panels = 4;
    j.b_off[0] = kFXg; j.b_panels[0] = 1;
    j.dW = dW; j.ld = ld; j.col_off[0] = 0; j.ncols[0] = 63;
    j.dbias = dbias;
    j.weight = narrow_w;
    return j;
  };
  const uint32_t dY = kGY, hL = kFH;
  a.jobs[n++] = narrow(dY + 0 * 4 * kPanel, G(0), 63, G(1));                                          // pts_linears.0
  for (int l = 1; l <= 7; ++l) {
    if (l == 5) {
      a.jobs[n++] = plain(dY + 5 * 4 * kPanel, hL + 4 * 4 * kPanel, G(10), 319, 63, G(11));           // [. | h4] part
      a.jobs[n++] = narrow(dY + 5 * 4 * kPanel, G(10), 319, nullptr);                                 // gamma(x) part
    } else {
      a.jobs[n++] = plain(dY + l * 4 * kPanel, hL + (l - 1) * 4 * kPanel, G(2 * l), 256, 0, G(2 * l + 1));
    }
  }
  if (!use_viewdirs) {  // output_linear on CUDA cores from the h7 tiles (staged as "extra" panels; the job has no MMA)
    WJob j{};
    j.x_off = hL + 7 * 4 * kPanel; j.x_panels = 4;
    j.simt = SIMT_OUT;
    j.head_w = G(NERFB_P_OUTPUT_W); j.head_b = G(NERFB_P_OUTPUT_B);
    j.weight = out_w;
    a.jobs[n++] = j;
  }
  if (use_viewdirs) {  // feature_linear (+ alpha_linear on CUDA cores from the same h7 tiles)
    WJob j = plain(kGFeat, hL + 7 * 4 * kPanel, G(NERFB_P_FEATURE_W), 256, 0, G(NERFB_P_FEATURE_B));
    j.simt = SIMT_ALPHA;
    j.head_w = G(NERFB_P_ALPHA_W); j.head_b = G(NERFB_P_ALPHA_B);
    a.jobs[n++] = j;
  }
  if (use_viewdirs) {  // views_linears.0: dhv^T [feature | gamma(d)]  (+ rgb_linear on CUDA cores from hv)
    WJob j{};
    j.a_off = kGHv; j.a_panels = 2;
    j.b_off[0] = kFFeat; j.b_panels[0] = 4; j.col_off[0] = 0; j.ncols[0] = 256;
    j.b_off[1] = kFDg; j.b_panels[1] = 1; j.col_off[1] = 256; j.ncols[1] = 27;
    j.x_off = kFHv; j.x_panels = 2;
    j.dW = G(NERFB_P_VIEWS_W); j.ld = 283;
    j.dbias = G(NERFB_P_VIEWS_B);
    j.simt = SIMT_RGB;
    j.head_w = G(NERFB_P_RGB_W); j.head_b = G(NERFB_P_RGB_B);
    j.weight = views_w;
    a.jobs[n++] = j;
  }
  a.njobs = n;
  a.acts = (const uint8_t*)acts;
  a.gstore = (const uint8_t*)gstore;
  a.d_raw = d_raw_padded;
  a.d_raw4 = d_raw4_padded;
  a.out_ch = out_ch;
  a.nstages = 8 * ((nrows + kSuperRows - 1) / kSuperRows);    // 4 tiles x 2 stages per 512-row super-tile
  a.watch = get_watch();
  // deal the SMs to the jobs in proportion to their per-stage bytes (largest remainder), at least one each
  NERFB_REQUIRE(grid >= a.njobs, "%s: %d CTAs for %d jobs", who, grid, a.njobs);
  {
    int total_w = 0;
    for (int j = 0; j < a.njobs; ++j) total_w += a.jobs[j].weight;
    int cnt[kMaxJobs], used = 0;
    double frac[kMaxJobs];
    for (int j = 0; j < a.njobs; ++j) {
      const double ideal = (double)grid * a.jobs[j].weight / total_w;
      cnt[j] = (int)ideal < 1 ? 1 : (int)ideal;
      frac[j] = ideal - cnt[j];
      used += cnt[j];
    }
    while (used < grid) {
      int best = 0;
      for (int j = 1; j < a.njobs; ++j) if (frac[j] > frac[best]) best = j;
      ++cnt[best]; frac[best] -= 1.0; ++used;
    }
    while (used > grid) {
      int best = -1;
      for (int j = 0; j < a.njobs; ++j) if (cnt[j] > 1 && (best < 0 || frac[j] < frac[best])) best = j;
      --cnt[best]; frac[best] += 1.0; --used;
    }
    a.cta_begin[0] = 0;
    for (int j = 0; j < a.njobs; ++j) a.cta_begin[j + 1] = a.cta_begin[j] + cnt[j];
  }
  // cost timeline of the stand-alone kernel (one CTA's slice must not span more than two jobs).  The ~njobs CTAs that
  // straddle a job boundary drain TMEM twice (35-150 us each), so the timeline pays only when a CTA's share of the stages
  // is long against a drain: measured break-even ~650 k samples (profiles/README.md); below it the jobs get whole CTAs
  long long timeline_min = 10000;                      // stages (64 samples each)
  if (const char* e = getenv("NERFB_WG_TIMELINE_MIN")) timeline_min = atoll(e);
  a.ncta = a.nstages >= timeline_min ? grid : 0;
  a.cost_begin[0] = 0;
  {
    long long total_w = 0;
    for (int j = 0; j < a.njobs; ++j) total_w += a.jobs[j].weight;
    for (int j = 0; j < a.njobs; ++j) {
      NERFB_REQUIRE(a.jobs[j].weight >= 1 && (long long)a.jobs[j].weight * grid >= total_w,
                    "%s: job %d (weight %d) is smaller than one CTA's share of %d CTAs", who, j, a.jobs[j].weight, grid);
      a.cost_begin[j + 1] = a.cost_begin[j] + (long long)a.jobs[j].weight * a.nstages;
    }
  }
  for (int j = 0; j < a.njobs; ++j) {
    const uint32_t off = a.jobs[j].a_off;
    a.jobs[j].flag_grp = off == (uint32_t)kGHv ? 0 : (off == (uint32_t)kGFeat ? 1 : 2 + (int)((off - kGY) / (4 * kPanel)));
  }
  return NERFB_OK;
}

}  // namespace nerfb
