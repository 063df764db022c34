// Self-test of the tcgen05 building blocks in tc05.cuh: one (cta_group 1) or
// two paired (cta_group 2) CTAs compute C = A B^T from operands written to
// shared memory by ordinary threads in the UMMA K-major/128B-swizzle layout --
// exactly what the fused MLP kernel's epilogue does for its activations.
#include "common.cuh"
#include "tc05.cuh"

namespace nerfb {
using namespace tc05;

static unsigned long long* g_watch_host = nullptr;
static unsigned long long* g_watch_dev = nullptr;

tc05::Watch get_watch() {
  if (g_watch_host == nullptr) {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 8 * sizeof(unsigned long long), cudaHostAllocMapped) == cudaSuccess) {
      g_watch_host = (unsigned long long*)h;
      for (int i = 0; i < 8; ++i) g_watch_host[i] = 0;
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) g_watch_dev = (unsigned long long*)d;
    }
    cudaGetLastError();
  }
  tc05::Watch w;
  w.rec = g_watch_dev;
  return w;
}

// writes a [rows x K] fp32 matrix (row stride ld) as K/64 swizzled bf16 chunks
__device__ void stage_operand(uint8_t* smem, const float* __restrict__ src, int rows, int K, int ld, int tid,
                              int nthreads) {
  const int units_per_row = K / 8;  // 16-byte units (8 bf16)
  for (int idx = tid; idx < rows * units_per_row; idx += nthreads) {
    const int r = idx / units_per_row, u = idx % units_per_row;
    const float* s = src + (size_t)r * ld + u * 8;
    uint4 v;
    v.x = pack_bf16x2(s[0], s[1]); v.y = pack_bf16x2(s[2], s[3]);
    v.z = pack_bf16x2(s[4], s[5]); v.w = pack_bf16x2(s[6], s[7]);
    const int chunk = u / 8, unit = u % 8;
    *reinterpret_cast<uint4*>(smem + (size_t)chunk * rows * 128 + sw128_offset(r, unit)) = v;
  }
}

template <int CG>
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(int N, int K, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                     Watch watch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  const int nb_rows = N / CG;                     // rows of B held by this CTA
  uint8_t* sA = smem;                             // K/64 chunks x 16 KB
  uint8_t* sB = smem + (size_t)(K / 64) * 128 * 128;

  stage_operand(sA, A + (size_t)rank * 128 * K, 128, K, K, tid, 128);
  stage_operand(sB, B + (size_t)rank * nb_rows * K, nb_rows, K, K, tid, 128);
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) {
    tmem_alloc<CG>(smem_u32(&tmem_base_slot), ncols);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128 * CG, N);
    for (int kc = 0; kc < K / 64; ++kc) {
      const uint64_t da = make_sdesc_sw128(smem_u32(sA + (size_t)kc * 128 * 128));
      const uint64_t db = make_sdesc_sw128(smem_u32(sB + (size_t)kc * nb_rows * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16<CG>(tmem_base, sdesc_advance_k(da, k), sdesc_advance_k(db, k), idesc, (kc | k) ? 1u : 0u);
    }
    umma_commit<CG>(smem_u32(&bar_done));
  }
  mbar_wait(smem_u32(&bar_done), 0, watch, 1);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  float* crow = C + ((size_t)rank * 128 + row) * N;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) crow[c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) tmem_dealloc<CG>(tmem_base, ncols);
}

// Bias-by-MMA self-test (the form the fused MLP uses to add a layer's bias on the tensor cores): 2-CTA pair,
// C[256,N] = A[:, 48:64] Bb^T + A Bw^T with A [256,64] and Bw [N,64] staged as 128B-swizzled K-major chunks and the bias
// tile Bb [N,16] staged in the NO-swizzle K-major layout (nosw16_offset); the first MMA reads K slice 3 of the A chunk.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma_selftest_bias_kernel(int N, const float* __restrict__ A, const float* __restrict__ Bw, const float* __restrict__ Bb,
                          const float* __restrict__ A8, float* __restrict__ C, Watch watch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  const int nb_rows = N / 2;
  uint8_t* sA = smem;                             // 16 KB
  uint8_t* sB = smem + 16384;                     // nb_rows x 128 B
  uint8_t* sBias = smem + 16384 + 16384;          // nb_rows x 32 B, no swizzle
  uint8_t* sA8 = sBias + 8192;                    // A8 != null: [8 rows x 16] bf16, no swizzle, every 8-row group aliases it (SBO = 0)
  if (A8 != nullptr)
    for (int idx = tid; idx < 8 * 16; idx += 128)
      *reinterpret_cast<__nv_bfloat16*>(sA8 + nosw16_offset(idx / 16, idx % 16)) = __float2bfloat16(A8[idx]);
  stage_operand(sA, A + (size_t)rank * 128 * 64, 128, 64, 64, tid, 128);
  stage_operand(sB, Bw + (size_t)rank * nb_rows * 64, nb_rows, 64, 64, tid, 128);
  for (int idx = tid; idx < nb_rows * 16; idx += 128) {
    const int n = idx / 16, k = idx % 16;
    *reinterpret_cast<__nv_bfloat16*>(sBias + nosw16_offset(n, k)) = __float2bfloat16(Bb[((size_t)rank * nb_rows + n) * 16 + k]);
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar_done), 1); fence_barrier_init(); }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) { tmem_alloc<2>(smem_u32(&tmem_base_slot), ncols); tmem_relinquish<2>(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_bf16(256, N);
    const uint64_t da = make_sdesc_sw128(smem_u32(sA));
    const uint64_t db = make_sdesc_sw128(smem_u32(sB));
    const uint64_t dbias = make_sdesc_nosw(smem_u32(sBias), 128, 256);
    if (A8 != nullptr) umma_bf16<2>(tmem_base, make_sdesc_nosw(smem_u32(sA8), 128, 0), dbias, idesc, 0u);
    else umma_bf16<2>(tmem_base, sdesc_advance_k(da, 3), dbias, idesc, 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16<2>(tmem_base, sdesc_advance_k(da, k), sdesc_advance_k(db, k), idesc, 1u);
    umma_commit<2>(smem_u32(&bar_done));
  }
  mbar_wait(smem_u32(&bar_done), 0, watch, 3);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  float* crow = C + ((size_t)rank * 128 + row) * N;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) crow[c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc<2>(tmem_base, ncols);
}

// MN-major self-test (the wgrad operand form): C[128,N] = At^T Bt with At [K,128] and Bt [K,N]
// row-major, staged as [K rows x 64-col panels] swizzled images, K multiple of 16 (<= 128).
__global__ void __launch_bounds__(128, 1)
umma_selftest_mn_kernel(int N, int K, const float* __restrict__ At, const float* __restrict__ Bt,
                        float* __restrict__ C, Watch watch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int panel_bytes = 128 * 128;              // panels are laid out as 128-row tiles (like the act images)
  uint8_t* sA = smem;                             // 2 panels (M = 128)
  uint8_t* sB = smem + 2 * panel_bytes;           // N/64 panels
  // stage: element (k, m) of At -> panel m/64, row k, col m%64
  for (int idx = tid; idx < 128 * (128 / 8); idx += 128) {
    const int k = idx / 16, u = idx % 16;         // u: 16-byte unit along M (8 elements)
    uint4 v = make_uint4(0, 0, 0, 0);
    if (k < K) {
      const float* s = At + (size_t)k * 128 + u * 8;
      v = make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
    }
    *reinterpret_cast<uint4*>(sA + (u / 8) * panel_bytes + sw128_offset(k, u % 8)) = v;
  }
  for (int idx = tid; idx < 128 * (N / 8); idx += 128) {
    const int k = idx / (N / 8), u = idx % (N / 8);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (k < K) {
      const float* s = Bt + (size_t)k * N + u * 8;
      v = make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]), pack_bf16x2(s[6], s[7]));
    }
    *reinterpret_cast<uint4*>(sB + (u / 8) * panel_bytes + sw128_offset(k, u % 8)) = v;
  }
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(smem_u32(&bar_done), 1); fence_barrier_init(); }
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;
  if (warp == 0) { tmem_alloc<1>(smem_u32(&tmem_base_slot), ncols); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16_mn(128, N);
    const uint64_t da = make_sdesc_sw128_mn(smem_u32(sA), panel_bytes);
    const uint64_t db = make_sdesc_sw128_mn(smem_u32(sB), panel_bytes);
    if (elect_one()) {
      for (int k = 0; k < K / 16; ++k)
        umma_bf16<1>(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, k ? 1u : 0u);
      umma_commit<1>(smem_u32(&bar_done));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar_done), 0, watch, 2);
  tc_fence_after();
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_base, ncols);
}


// Probe of the 1-D bulk-copy path (global/L2 -> shared): every CTA streams `nchunks` chunks of
// `chunk_bytes` (split in `pieces` copies) from the same buffer through a ring of `depth` slots and
// reports cycles per chunk in steady state.  Used to size the weight ring of the fused MLP kernel.
__global__ void __launch_bounds__(32, 1)
bulk_probe_kernel(const uint8_t* __restrict__ src, int src_bytes, int nchunks, int chunk_bytes, int pieces, int depth,
                  long long* __restrict__ out, Watch watch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    const uint32_t base = smem_u32(smem_raw);
    const int piece = chunk_bytes / pieces;
    long long t0 = 0;
    const int warm = depth * 4;
    for (int i = 0; i < nchunks + depth; ++i) {
      const int slot = i % depth;
      if (i >= depth) mbar_wait(smem_u32(&bars[slot]), (uint32_t)(((i / depth) - 1) & 1), watch, 0x77);
      if (i == warm) t0 = clock64();
      if (i < nchunks) {
        const size_t off = ((size_t)i * chunk_bytes + (size_t)blockIdx.x * 4096) % (size_t)(src_bytes - chunk_bytes);
        mbar_arrive_expect_tx(smem_u32(&bars[slot]), chunk_bytes);
        for (int pc = 0; pc < pieces; ++pc)
          bulk_g2s(base + slot * chunk_bytes + pc * piece, src + (off & ~(size_t)15) + pc * piece, piece,
                   smem_u32(&bars[slot]));
      }
    }
    out[blockIdx.x] = (clock64() - t0) / (nchunks + depth - warm);
  }
}


// Probe 2: the bulk-copy stream of bulk_probe_kernel running CONCURRENTLY with back-to-back UMMAs on static
// operands (mode 0: no MMA, 1: SS form -- A and B from shared memory, 2: TS form -- A from tensor memory) and,
// optionally, 8 warps of 16-byte shared-memory stores (the epilogue's store traffic).  Reports cycles per
// 16 KB chunk and cycles per MMA (M=128, N=256, K=16, cta_group::1).  Answers: who competes for shared memory?
__global__ void __launch_bounds__(320, 1)
contention_probe_kernel(const uint8_t* __restrict__ src, int src_bytes, int nchunks, int mode, int sts_warps,
                        long long* __restrict__ out, Watch watch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[8];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kChunk = 16384, kDepth = 4;
  // layout: [A 16 KB][B 32 KB][ring 4 x 16 KB][sts scratch 32 KB]
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  uint8_t* ring = smem + 49152;
  uint8_t* scratch = ring + kDepth * kChunk;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
    fence_barrier_init();
    stop_flag = 0;
  }
  for (int i = tid; i < 49152 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0x3c003c00u, 0, 0x3c003c00u, 0);
  fence_proxy_async_smem();
  if (warp == 1) { tmem_alloc<1>(smem_u32(&tmem_slot), 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t base = smem_u32(ring);
      long long t0 = 0, lat_sum = 0;
      const int warm = 16;
      for (int i = 0; i < nchunks + kDepth; ++i) {
        const int slot = i % kDepth;
        if (i >= kDepth) mbar_wait(smem_u32(&bars[slot]), (uint32_t)(((i / kDepth) - 1) & 1), watch, 0x78);
        if (i == warm) t0 = clock64();
        if (i < nchunks) {
          const uint32_t off = ((uint32_t)i * kChunk + blockIdx.x * 4096u) % (uint32_t)(src_bytes - kChunk);
          mbar_arrive_expect_tx(smem_u32(&bars[slot]), kChunk);
          bulk_g2s(base + slot * kChunk, src + (off & ~15u), kChunk, smem_u32(&bars[slot]));
        }
      }
      out[blockIdx.x * 4 + 0] = (clock64() - t0) / (nchunks + kDepth - warm);
      (void)lat_sum;
      stop_flag = 1;
    }
  } else if (warp == 1) {
    if (mode != 0) {
      const uint32_t idesc = make_idesc_bf16(128, 256);
      const uint64_t da = make_sdesc_sw128(smem_u32(sA)), db = make_sdesc_sw128(smem_u32(sB));
      long long n = 0;
      const long long t0 = clock64();
      uint32_t phase = 0;
      while (!stop_flag) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (mode == 1) umma_bf16<1>(tmem_base, da + (uint64_t)((k & 3) * 2), db + (uint64_t)((k & 3) * 2), idesc, 1u);
            else umma_bf16_ts<1>(tmem_base, tmem_base + 256 + (k & 3) * 8, db + (uint64_t)((k & 3) * 2), idesc, 1u);
          }
          umma_commit<1>(smem_u32(&bars[4]));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bars[4]), phase, watch, 0x79);
        phase ^= 1;
        n += 16;
      }
      if (lane == 0) { out[blockIdx.x * 4 + 1] = (clock64() - t0) / (n > 0 ? n : 1); out[blockIdx.x * 4 + 2] = n; }
    }
  } else if (warp - 2 < sts_warps) {
    // epilogue-like store traffic: every thread writes 16 B per iteration, conflict-free rows
    uint4 v = make_uint4(tid, tid, tid, tid);
    uint8_t* p = scratch + (size_t)(tid & 255) * 128;
    long long n = 0;
    while (!stop_flag) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p + (((i ^ tid) & 7) << 4))), "r"(v.x),
                     "r"(v.y), "r"(v.z), "r"(v.w)
                     : "memory");
        v.x += 1;
      }
      n += 8;
    }
    if (tid == 64) out[blockIdx.x * 4 + 3] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_umma_selftest_mn(int N, int K, const float* At, const float* Bt, float* C, void* stream) {
  NERFB_REQUIRE(N >= 64 && N <= 256 && N % 64 == 0 && K >= 16 && K <= 128 && K % 16 == 0 && At && Bt && C,
                "nerfb_umma_selftest_mn: bad arguments (N=%d K=%d)", N, K);
  const size_t smem = (size_t)(2 + N / 64) * 128 * 128 + 1024;
  tc05::Watch w = get_watch();
  NERFB_CUDA(cudaFuncSetAttribute(umma_selftest_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_mn_kernel<<<1, 128, smem, as_stream(stream)>>>(N, K, At, Bt, C, w);
  return check_launch("umma_selftest_mn_kernel");
}

extern "C" int nerfb_umma_selftest_bias(int N, const float* A, const float* Bw, const float* Bb, const float* A8, float* C,
                                        void* stream) {
  NERFB_REQUIRE((N == 256 || N == 128) && A && Bw && Bb && C, "nerfb_umma_selftest_bias: bad arguments (N=%d)", N);
  const size_t smem = 16384 + 16384 + 8192 + 256 + 1024;
  tc05::Watch w = get_watch();
  NERFB_CUDA(cudaFuncSetAttribute(umma_selftest_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_bias_kernel<<<2, 128, smem, as_stream(stream)>>>(N, A, Bw, Bb, A8, C, w);
  return check_launch("umma_selftest_bias_kernel");
}

extern "C" const unsigned long long* nerfb_watchdog_record(void) {
  get_watch();
  return g_watch_host;
}

extern "C" int nerfb_umma_selftest(int cta_group, int N, int K, const float* A, const float* B, float* C,
                                   void* stream) {
  NERFB_REQUIRE((cta_group == 1 || cta_group == 2) && N >= 32 && N <= 256 && N % 32 == 0 && K >= 64 && K <= 256 &&
                    K % 64 == 0 && A && B && C,
                "nerfb_umma_selftest: bad arguments (cta_group=%d N=%d K=%d)", cta_group, N, K);
  const size_t smem = (size_t)(K / 64) * 128 * 128 + (size_t)(K / 64) * (N / cta_group) * 128 + 1024;
  tc05::Watch w = get_watch();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cta_group);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cta_group == 1) {
    NERFB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NERFB_CUDA(cudaLaunchKernelEx(&cfg, umma_selftest_kernel<1>, N, K, A, B, C, w));
  } else {
    NERFB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NERFB_CUDA(cudaLaunchKernelEx(&cfg, umma_selftest_kernel<2>, N, K, A, B, C, w));
  }
  return check_launch("umma_selftest_kernel");
}

extern "C" int nerfb_bulk_probe(const void* src, int src_bytes, int nchunks, int chunk_bytes, int pieces, int depth,
                                int nblocks, long long* out_cycles, void* stream) {
  NERFB_REQUIRE(src && out_cycles && depth >= 1 && depth <= 16 && chunk_bytes % (16 * pieces) == 0 &&
                    depth * chunk_bytes <= 200 * 1024 && src_bytes > 2 * chunk_bytes && nchunks > 8 * depth,
                "nerfb_bulk_probe: bad arguments");
  const size_t smem = (size_t)depth * chunk_bytes + 1024;
  tc05::Watch w = get_watch();
  NERFB_CUDA(cudaFuncSetAttribute(bulk_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bulk_probe_kernel<<<nblocks, 32, smem, as_stream(stream)>>>((const uint8_t*)src, src_bytes, nchunks, chunk_bytes,
                                                             pieces, depth, out_cycles, w);
  return check_launch("bulk_probe_kernel");
}

extern "C" int nerfb_contention_probe(const void* src, int src_bytes, int nchunks, int mode, int sts_warps, int nblocks,
                                      long long* out, void* stream) {
  NERFB_REQUIRE(src && out && src_bytes > 65536 && nchunks > 64 && mode >= 0 && mode <= 2 && sts_warps >= 0 &&
                    sts_warps <= 8, "nerfb_contention_probe: bad arguments");
  const size_t smem = 49152 + 4 * 16384 + 32768 + 1024;
  tc05::Watch w = get_watch();
  NERFB_CUDA(cudaFuncSetAttribute(contention_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  contention_probe_kernel<<<nblocks, 320, smem, as_stream(stream)>>>((const uint8_t*)src, src_bytes, nchunks, mode,
                                                                     sts_warps, out, w);
  return check_launch("contention_probe_kernel");
}
