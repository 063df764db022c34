// Blackwell (sm_100a) primitives used by the fused MLP kernel: mbarrier,
// bulk async copy (TMA engine, 1-D), tcgen05 MMA / TMEM, cluster helpers and
// the UMMA shared-memory layout (K-major, 128-byte swizzle).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace tc05 {

// ---------------------------------------------------------------- addresses
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// one lane of a fully converged warp (the pattern tcgen05 issue wants: the surrounding
// control flow stays warp-uniform so operands can live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- watchdog
// Every wait is bounded: on timeout a record is written to host-mapped memory
// (survives the trap) and the kernel traps, so a protocol bug is an error with
// a location, never a hung GPU.
struct Watch {
  volatile unsigned long long* rec;   // host-mapped, 8 x u64, may be null
};

__device__ __forceinline__ void watchdog_fire(Watch w, uint32_t tag, uint32_t a, uint32_t b) {
  if (w.rec != nullptr) {
    w.rec[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
    w.rec[2] = tag;
    w.rec[3] = ((unsigned long long)a << 32) | b;
    __threadfence_system();
    w.rec[0] = 0xDEADull;
    __threadfence_system();
  }
  __trap();
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that lives in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  const uint32_t remote = mapa(bar, rank);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// same with the default (CTA-scope release) semantics, as cutlass::arch::ClusterBarrier::arrive(cta_id):
// used where the guarded data is consumed by the async/tensor proxy of the SM that wrote it
// (after fence.proxy.async + a CTA barrier), so no cluster-scope memory fence is needed.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  const uint32_t remote = mapa(bar, rank);
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// cluster-scope acquire variant (observes remote arrivals / multicast commits)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
constexpr long long kWatchdogCycles = 4000000000ll;  // ~2 s at 2 GHz

// CTA-scope acquire (what CUTLASS's ClusterBarrier::wait uses, also for barriers with
// remote arrivals): a cluster-scope acquire makes ptxas emit CCTL.IVALL (L1 invalidate)
// after every wait, which was the top stall of the epilogue warps (profiles/r01).
// The data guarded by these barriers is read through the async/tensor proxy or TMEM.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, Watch w, uint32_t tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) watchdog_fire(w, tag, bar, parity);
  }
}

// ------------------------------------------------------- async proxy / TMA
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 1-D bulk copy global -> shared (TMA engine; SASS UBLKCP), completes `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}


// 2-D tiled TMA load global -> shared of the executing CTA, completing its bytes on an mbarrier that may live
// in the PEER CTA of the pair (.cta_group::2): both CTAs of a 2-CTA MMA signal the leader's barrier directly.
// `mbar_cluster` is a shared::cluster address (mapa).  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t mbar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst), "l"(tmap), "r"(mbar_cluster), "r"(c0), "r"(c1)
      : "memory");
}

// 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
  // The training tile stores are write-once streams of gigabytes: marked evict_first in L2 so that they leave the cache
  // promptly instead of displacing each other as dirty normal-priority lines (measured on 786 432 samples:
  // forward-with-save 1.095 -> 0.994 ms, data-gradient chain 0.890 -> 0.855 ms; profiles/README.md, round 2)
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// all but the 4 most recently committed bulk groups of this thread have COMPLETED (their global writes are done)
__device__ __forceinline__ void bulk_wait4() { asm volatile("cp.async.bulk.wait_group 4;" ::: "memory"); }

// bulk prefetch of a global range into L2 (no destination; fire and forget)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- tcgen05
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32; issued by ONE thread
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// One scheduled "op" of the fused MLP kernel as a single instruction sequence, executed by a
// fully converged warp: (optional) non-blocking probes of the barriers the NEXT op needs, then
// four K=16 MMAs (cta_group::2) and up to three commits issued by one elected lane, and only
// then the probe results are read -- so the ~90-cycle try_wait latency overlaps the MMA issue
// instead of idling the tensor core (whose queue is ~1 MMA deep).
__device__ __forceinline__ void umma_op4_cg2(uint32_t d_tmem, uint64_t da0, uint64_t da1, uint64_t da2, uint64_t da3,
                                             uint64_t db0, uint64_t db1, uint64_t db2, uint64_t db3, uint32_t idesc,
                                             uint32_t acc0, uint32_t bar_rel, uint32_t f_rel, uint32_t bar_enc,
                                             uint32_t f_enc, uint32_t bar_done, uint32_t f_done, uint32_t need_w,
                                             uint32_t bar_w, uint32_t par_w, uint32_t need_a, uint32_t bar_a,
                                             uint32_t par_a, uint32_t& ok_w, uint32_t& ok_a) {
  asm volatile(
      "{\n\t"
      ".reg .pred Pe, Pacc, Pw, Pa, Pnw, Pna, Pf, Pc, Pt;\n\t"
      ".reg .b16 mask;\n\t"
      "mov.u16 mask, 3;\n\t"
      "setp.eq.b32 Pt, 0, 0;\n\t"
      "setp.ne.b32 Pnw, %19, 0;\n\t"
      "setp.eq.b32 Pw, %19, 0;\n\t"
      "setp.ne.b32 Pna, %22, 0;\n\t"
      "setp.eq.b32 Pa, %22, 0;\n\t"
      "@Pnw mbarrier.try_wait.parity.shared::cta.b64 Pw, [%20], %21;\n\t"
      "@Pna mbarrier.try_wait.parity.shared::cta.b64 Pa, [%23], %24;\n\t"
      "elect.sync _|Pe, 0xffffffff;\n\t"
      "setp.ne.b32 Pacc, %12, 0;\n\t"
      "@Pe tcgen05.mma.cta_group::2.kind::f16 [%2], %3, %7, %11, Pacc;\n\t"
      "@Pe tcgen05.mma.cta_group::2.kind::f16 [%2], %4, %8, %11, Pt;\n\t"
      "@Pe tcgen05.mma.cta_group::2.kind::f16 [%2], %5, %9, %11, Pt;\n\t"
      "@Pe tcgen05.mma.cta_group::2.kind::f16 [%2], %6, %10, %11, Pt;\n\t"
      "setp.ne.b32 Pf, %14, 0;\n\t"
      "and.pred Pc, Pe, Pf;\n\t"
      "@Pc tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%13], mask;\n\t"
      "setp.ne.b32 Pf, %16, 0;\n\t"
      "and.pred Pc, Pe, Pf;\n\t"
      "@Pc tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%15], mask;\n\t"
      "setp.ne.b32 Pf, %18, 0;\n\t"
      "and.pred Pc, Pe, Pf;\n\t"
      "@Pc tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%17], mask;\n\t"
      "selp.u32 %0, 1, 0, Pw;\n\t"
      "selp.u32 %1, 1, 0, Pa;\n\t"
      "}"
      : "=r"(ok_w), "=r"(ok_a)
      : "r"(d_tmem), "l"(da0), "l"(da1), "l"(da2), "l"(da3), "l"(db0), "l"(db1), "l"(db2), "l"(db3), "r"(idesc),
        "r"(acc0), "r"(bar_rel), "r"(f_rel), "r"(bar_enc), "r"(f_enc), "r"(bar_done), "r"(f_done), "r"(need_w),
        "r"(bar_w), "r"(par_w), "r"(need_a), "r"(bar_a), "r"(par_a)
      : "memory");
}


// D[tmem] (+)= A[tmem] * B[smem]^T (TS form: the A operand is read from tensor memory)
template <int CG>
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// registers -> TMEM: 32 lanes x 32 consecutive 32-bit columns (one row per thread)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// all previously issued MMAs of this thread complete -> arrive(1) on `bar`
// (CG==2: on the barrier at the same offset in every CTA of `mask`)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint16_t mask = 3) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns (one row per thread)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------ UMMA smem layout
// K-major operand tile of [rows x 64] bf16 ("K chunk"), SWIZZLE_128B: row r is
// 128 contiguous bytes at r*128, 8-row groups 1024 B apart, the 16-byte unit
// index (address bits 4..6) XOR-ed with r%8.  Tile base must be 1024-B aligned.
constexpr int kChunkK = 64;                       // bf16 elements per row of a chunk
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t unit16) {
  return row * 128u + ((unit16 ^ (row & 7u)) << 4);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1)
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}
// K-major operand WITHOUT swizzle (canonical "interleave" layout): core matrices of 8 rows x 16 bytes (8 bf16 along K)
// stored as 128 contiguous bytes; LBO = byte distance between the core matrices of consecutive K groups, SBO = byte
// distance between consecutive 8-row groups.  Used for the [rows x 16] bias tiles of the fused MLP (one K=16 MMA).
__device__ __forceinline__ uint64_t make_sdesc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell); layout type 0 = no swizzle
  return d;
}
// byte offset of element (row, k) of a [rows x 16] bf16 tile in that layout with LBO = 128, SBO = 256
__host__ __device__ __forceinline__ uint32_t nosw16_offset(uint32_t row, uint32_t k) {
  return (row >> 3) * 256u + (k >> 3) * 128u + (row & 7u) * 16u + (k & 7u) * 2u;
}
// advance along K inside the 128-byte swizzle atom: +32 B per UMMA_K=16 bf16
__device__ __forceinline__ uint64_t sdesc_advance_k(uint64_t d, int k16) { return d + (uint64_t)(k16 * 2); }

// MN-major operand (the contraction index K runs over the ROWS of the stored tile): the same
// physical image -- rows of 128 B (64 bf16 along M/N), 8-row groups 1024 B apart, 128B swizzle --
// read "transposed".  LBO = byte stride between 64-element M/N panels, SBO = 1024 (8 K-rows).
// A K=16 step advances the start address by 16 rows = 2048 B.
__device__ __forceinline__ uint64_t make_sdesc_sw128_mn(uint32_t smem_addr, uint32_t panel_stride_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((panel_stride_bytes >> 4) & 0x3FFF) << 16;   // leading byte offset
  d |= (uint64_t)(1024 >> 4) << 32;                            // stride byte offset
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor with both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// instruction descriptor: kind::f16, A/B = bf16 K-major, D = fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// max(x,0) fused into the conversion
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// 32 sign bits (column j of a 32-column batch at bit 31 - j) -> zero the bf16 halves of w[16] whose sign bit is set.
// prmt's sign-replicate mode turns the top bit of a selected byte into 0x00 / 0xFF, so after shifting the wanted bit to a
// byte's msb one prmt builds the AND mask of a packed pair: 7 shifts + 16 prmt + 16 lop3 per 32 columns.
__device__ __forceinline__ void apply_sign_mask32(uint32_t (&w)[16], uint32_t m) {
  uint32_t ms[9];
#pragma unroll
  for (int s = 0; s < 9; ++s) ms[s] = m << s;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int j = 2 * i, s = j & 7, a = j >> 3;                 // column j sits at the msb of byte 3-a of (m << s)
    const uint32_t sel = (uint32_t)((8 | (3 - a)) * 0x11) | ((uint32_t)((8 | (7 - a)) * 0x11) << 8);
    uint32_t neg;                                               // 0xFFFF per half whose unit was inactive
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(neg) : "r"(ms[s]), "r"(ms[s + 1]), "r"(sel));
    w[i] &= ~neg;
  }
}

// per bf16 half: 0xFFFF where the half is non-zero (ReLU mask from a stored post-ReLU activation pair)
__device__ __forceinline__ uint32_t bf16x2_nonzero_mask(uint32_t h2) {
  uint32_t d;
  const uint32_t zero = 0;
  asm("set.neu.u32.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(h2), "r"(zero));
  return d;
}

}  // namespace tc05
