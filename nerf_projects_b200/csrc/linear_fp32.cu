// fp32 SIMT linear layers: the "FP32 validation mode" of the MLP
// (nn.Linear/F.relu of nerf.py:70-109 and their autograd backward).
// One 128x128x16 register-tiled SGEMM kernel, three operand layouts:
//   MODE 0  fwd   : Y[M,N]  = act(X[M,K] W[N,K]^T + b)
//   MODE 1  dgrad : dX[M,K] = (dX +) dY[M,N] W[N,K]  (* relu mask)
//   MODE 2  wgrad : dW[N,K] += dY[M,N]^T X[M,K]      (split over M, atomics)
// Generic strides, any M/N/K; correctness first -- the tensor-core path in
// mlp_tc.cu is the fast one.
#include "common.cuh"

namespace nerfb {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

struct GemmP {
  int Mo, No, Kr;          // output rows, output cols, reduction length
  const float* A; int lda; // operand giving output rows
  const float* B; int ldb; // operand giving output cols
  float* C; int ldc;
  const float* bias;       // MODE 0
  int relu;                // MODE 0
  const float* mask; int ldmask; int accumulate;  // MODE 1
  int k_per_split;         // MODE 2
};

// A_KCONTIG: element (i,k) at A[i*lda+k]; else at A[k*lda+i]
template <bool KCONTIG>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int ld, int r0, int rmax, int k0,
                                          int kmax, float (*dst)[BM + 4], int tid) {
  if (KCONTIG) {
    const int row = tid >> 1, ks = (tid & 1) * 8;
    const int gr = r0 + row;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int gk = k0 + ks + q;
      dst[ks + q][row] = (gr < rmax && gk < kmax) ? __ldg(P + (size_t)gr * ld + gk) : 0.f;
    }
  } else {
    const int k = tid >> 4, js = (tid & 15) * 8;
    const int gk = k0 + k;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int gr = r0 + js + q;
      dst[k][js + q] = (gr < rmax && gk < kmax) ? __ldg(P + (size_t)gk * ld + gr) : 0.f;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT) sgemm_kernel(GemmP p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  int kbeg = 0, kend = p.Kr;
  if (MODE == 2) {
    kbeg = blockIdx.z * p.k_per_split;
    kend = min(p.Kr, kbeg + p.k_per_split);
  }
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    load_tile<MODE != 2>(p.A, p.lda, i0, p.Mo, k0, kend, As, tid);
    load_tile<MODE == 0>(p.B, p.ldb, j0, p.No, k0, kend, Bs, tid);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int gi = i0 + (a < 4 ? ty * 4 + a : 64 + ty * 4 + (a - 4));
    if (gi >= p.Mo) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int gj = j0 + (b < 4 ? tx * 4 + b : 64 + tx * 4 + (b - 4));
      if (gj >= p.No) continue;
      float v = acc[a][b];
      float* dst = p.C + (size_t)gi * p.ldc + gj;
      if (MODE == 0) {
        if (p.bias) v += __ldg(p.bias + gj);
        if (p.relu) v = fmaxf(v, 0.f);
        *dst = v;
      } else if (MODE == 1) {
        if (p.accumulate) v += *dst;
        if (p.mask) v = (__ldg(p.mask + (size_t)gi * p.ldmask + gj) > 0.f) ? v : 0.f;
        *dst = v;
      } else {
        atomicAdd(dst, v);
      }
    }
  }
}

// db[n] += sum_m dY[m,n]
__global__ void colsum_kernel(int M, int N, const float* __restrict__ dY, int ld, float* __restrict__ db,
                              int rows_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  for (int m = m0; m < m1; ++m) s += __ldg(dY + (size_t)m * ld + n);
  atomicAdd(db + n, s);
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_linear_fwd(int M, int N, int K, const float* X, int ldx, const float* W, int ldw,
                                const float* b, float* Y, int ldy, int relu, void* stream) {
  NERFB_REQUIRE(M >= 0 && N > 0 && K > 0 && X && W && Y && ldx >= K && ldw >= K && ldy >= N,
                "nerfb_linear_fwd: bad arguments (M=%d N=%d K=%d)", M, N, K);
  if (M == 0) return NERFB_OK;
  GemmP p{};
  p.Mo = M; p.No = N; p.Kr = K; p.A = X; p.lda = ldx; p.B = W; p.ldb = ldw; p.C = Y; p.ldc = ldy;
  p.bias = b; p.relu = relu;
  dim3 grid(cdiv(N, BN), cdiv(M, BM), 1);
  sgemm_kernel<0><<<grid, NT, 0, as_stream(stream)>>>(p);
  return check_launch("sgemm_kernel<fwd>");
}

extern "C" int nerfb_linear_bwd_data(int M, int N, int K, const float* dY, int lddy, const float* W, int ldw,
                                     float* dX, int lddx, const float* relu_src, int ld_relu, int accumulate,
                                     void* stream) {
  NERFB_REQUIRE(M >= 0 && N > 0 && K > 0 && dY && W && dX && lddy >= N && ldw >= K && lddx >= K,
                "nerfb_linear_bwd_data: bad arguments (M=%d N=%d K=%d)", M, N, K);
  if (M == 0) return NERFB_OK;
  GemmP p{};
  p.Mo = M; p.No = K; p.Kr = N; p.A = dY; p.lda = lddy; p.B = W; p.ldb = ldw; p.C = dX; p.ldc = lddx;
  p.mask = relu_src; p.ldmask = ld_relu; p.accumulate = accumulate;
  dim3 grid(cdiv(K, BN), cdiv(M, BM), 1);
  sgemm_kernel<1><<<grid, NT, 0, as_stream(stream)>>>(p);
  return check_launch("sgemm_kernel<dgrad>");
}

extern "C" int nerfb_linear_bwd_weight(int M, int N, int K, const float* dY, int lddy, const float* X, int ldx,
                                       float* dW, int lddw, float* db, void* stream) {
  NERFB_REQUIRE(M >= 0 && N > 0 && K > 0 && dY && X && dW && lddy >= N && ldx >= K && lddw >= K,
                "nerfb_linear_bwd_weight: bad arguments (M=%d N=%d K=%d)", M, N, K);
  if (M == 0) return NERFB_OK;
  GemmP p{};
  p.Mo = N; p.No = K; p.Kr = M; p.A = dY; p.lda = lddy; p.B = X; p.ldb = ldx; p.C = dW; p.ldc = lddw;
  const int tiles = cdiv(N, BM) * cdiv(K, BN);
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  const int max_splits = cdiv(M, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int kps = cdiv(M, splits);
  kps = cdiv(kps, BK) * BK;
  splits = cdiv(M, kps);
  p.k_per_split = kps;
  dim3 grid(cdiv(K, BN), cdiv(N, BM), splits);
  sgemm_kernel<2><<<grid, NT, 0, as_stream(stream)>>>(p);
  int rc = check_launch("sgemm_kernel<wgrad>");
  if (rc != NERFB_OK) return rc;
  if (db) {
    const int rpb = 2048;
    dim3 g2(cdiv(N, 128), cdiv(M, rpb));
    colsum_kernel<<<g2, 128, 0, as_stream(stream)>>>(M, N, dY, lddy, db, rpb);
    rc = check_launch("colsum_kernel");
  }
  return rc;
}
