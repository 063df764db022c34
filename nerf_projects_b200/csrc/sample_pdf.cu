// Hierarchical sampling: sample_pdf (nerf_helpers.py:372-439) and the
// z_mid / sample / sort / z_std block of render_rays (ipynb:459-467, 486).
// One warp per ray.  The arithmetic follows oracle/sample_pdf_np.py operation
// by operation so that indices AND samples are bit-identical to the
// reference's CPU fp32 result:
//   * S = torch.sum(w) in ATen's vectorised order (8 lanes x 4 ILP slots,
//     4-level cascade, sequential scalar tail, lanes added in order);
//   * pdf = w / S with IEEE division; cdf = fp64 running sum rounded to fp32
//     (all partial sums are exact in fp64, so a parallel scan is bit-safe);
//   * u = torch.linspace formula (single-rounded fma in the upper half);
//   * searchsorted(right=True) by binary search; un-fused fp32 lerp.
#include "composite.cuh"
#include <math_constants.h>

namespace nerfb {

constexpr int PW = 4;  // warps per block

__device__ __forceinline__ int ceil_log2_aten(int x) {  // c10::utils::CeilLog2
  if (x <= 2) return 1;
  return 32 - __clz(x - 1);
}

// ATen multi_row_sum for one (ILP slot, vector lane): items w[(4*i+k)*stride_v + l]
__device__ __forceinline__ float cascade_sum(const float* __restrict__ w, int size_ilp, int k, int lanes, int l) {
  const int level_power = max(4, ceil_log2_aten(size_ilp) / 4);
  const int level_step = 1 << level_power;
  const int level_mask = level_step - 1;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int i = 0;
  while (i + level_step <= size_ilp) {
    for (int j = 0; j < level_step; ++j, ++i) acc[0] = __fadd_rn(acc[0], w[(4 * i + k) * lanes + l]);
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      acc[j] = __fadd_rn(acc[j], acc[j - 1]);
      acc[j - 1] = 0.f;
      if ((i & (level_mask << (j * level_power))) != 0) break;
    }
  }
  for (; i < size_ilp; ++i) acc[0] = __fadd_rn(acc[0], w[(4 * i + k) * lanes + l]);
#pragma unroll
  for (int j = 1; j < 4; ++j) acc[0] = __fadd_rn(acc[0], acc[j]);
  return acc[0];
}

// fp32 torch.sum over w[0..n) (contiguous) in ATen CPU order; result on all lanes
__device__ float aten_sum_f32(const float* __restrict__ w, int n, int lane) {
  float result = 0.f;
  if (n < 8) {
    if (lane == 0) {
      const int size_ilp = n / 4;
      float part[4];
      for (int k = 0; k < 4; ++k) part[k] = cascade_sum(w, size_ilp, k, 1, 0);
      for (int i = size_ilp * 4; i < n; ++i) part[0] = __fadd_rn(part[0], w[i]);
      for (int k = 1; k < 4; ++k) part[0] = __fadd_rn(part[0], part[k]);
      result = part[0];
    }
    return __shfl_sync(0xffffffffu, result, 0);
  }
  const int nvec = n / 8;
  const int size_ilp = nvec / 4;
  const int k = lane >> 3, l = lane & 7;
  float part = cascade_sum(w, size_ilp, k, 8, l);
  if (k == 0)
    for (int v = size_ilp * 4; v < nvec; ++v) part = __fadd_rn(part, w[v * 8 + l]);
#pragma unroll
  for (int kk = 1; kk < 4; ++kk) {
    const float other = __shfl_sync(0xffffffffu, part, kk * 8 + l);
    if (k == 0) part = __fadd_rn(part, other);
  }
  float tail = 0.f;
  if (lane == 0)
    for (int j = nvec * 8; j < n; ++j) tail = __fadd_rn(tail, w[j]);
#pragma unroll
  for (int ll = 0; ll < 8; ++ll) {
    const float v = __shfl_sync(0xffffffffu, part, ll);
    if (lane == 0) tail = __fadd_rn(tail, v);
  }
  return __shfl_sync(0xffffffffu, tail, 0);
}

// Builds the CDF for one ray.  s_w: n = M-1 weights already offset by +1e-5;
// writes s_cdf[0..M).
__device__ void build_cdf(float* __restrict__ s_w, float* __restrict__ s_cdf, int M, int lane) {
  const int n = M - 1;
  const float S = aten_sum_f32(s_w, n, lane);
  const int seg = (n + 31) / 32;
  const int j0 = min(n, lane * seg), j1 = min(n, j0 + seg);
  double local = 0.0;
  for (int j = j0; j < j1; ++j) local += (double)__fdiv_rn(s_w[j], S);
  double incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double q = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += q;
  }
  double run = incl - local;
  if (lane == 0) s_cdf[0] = 0.f;
  for (int j = j0; j < j1; ++j) {
    run += (double)__fdiv_rn(s_w[j], S);
    s_cdf[j + 1] = (float)run;
  }
  __syncwarp();
}

__device__ __forceinline__ float draw_sample(const float* __restrict__ s_cdf, const float* __restrict__ s_bins,
                                             int M, float u, int* ind_out) {
  int lo = 0, hi = M;
  while (lo < hi) {  // searchsorted(right=True): #{j : cdf[j] <= u}
    const int mid = (lo + hi) >> 1;
    if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  *ind_out = lo;
  const int below = max(lo - 1, 0), above = min(lo, M - 1);
  const float c0 = s_cdf[below], c1 = s_cdf[above];
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  const float b0 = s_bins[below], b1 = s_bins[above];
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

__global__ void __launch_bounds__(32 * PW)
sample_pdf_kernel(int R, int M, int N, const float* __restrict__ bins, int ld_bins,
                  const float* __restrict__ weights, int ld_w, int det, const float* __restrict__ u_in,
                  float* __restrict__ samples, long long* __restrict__ inds) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = blockIdx.x * PW + wib;
  if (r >= R) return;
  float* s_w = smem + (size_t)wib * 3 * M;
  float* s_cdf = s_w + M;
  float* s_bins = s_cdf + M;
  for (int j = lane; j < M - 1; j += 32) s_w[j] = __fadd_rn(weights[(size_t)r * ld_w + j], 1e-5f);
  for (int j = lane; j < M; j += 32) s_bins[j] = bins[(size_t)r * ld_bins + j];
  __syncwarp();
  build_cdf(s_w, s_cdf, M, lane);
  for (int i = lane; i < N; i += 32) {
    const float u = det ? linspace01(i, N) : u_in[(size_t)r * N + i];
    int ind;
    samples[(size_t)r * N + i] = draw_sample(s_cdf, s_bins, M, u, &ind);
    if (inds) inds[(size_t)r * N + i] = ind;
  }
}

// The z_mid / sample_pdf / sort / z_std block of render_rays for ONE ray (one warp).  `wr` = the ray's S compositing
// weights (global OR shared memory: the fused kernel below keeps them on-chip), `s_base` = 3*(S-1) + P floats of scratch.
__device__ void hierarchical_z_ray(int S, int N, int P, const float* zr, const float* wr, int det, const float* u_row,
                                   float* z_out_row, float* z_samples_row, long long* inds_row, float* z_std,
                                   float* s_base, int lane) {
  const int M = S - 1;  // bins = z_mid (S-1 edges), weights[1:-1] (S-2 values)
  float* s_w = s_base;
  float* s_cdf = s_w + M;
  float* s_bins = s_cdf + M;
  float* s_sort = s_bins + M;
  for (int j = lane; j < M - 1; j += 32) s_w[j] = __fadd_rn(wr[j + 1], 1e-5f);
  for (int j = lane; j < M; j += 32) s_bins[j] = __fmul_rn(0.5f, __fadd_rn(zr[j + 1], zr[j]));
  for (int j = lane; j < S; j += 32) s_sort[j] = zr[j];
  for (int j = S + N + lane; j < P; j += 32) s_sort[j] = CUDART_INF_F;
  __syncwarp();
  build_cdf(s_w, s_cdf, M, lane);
  double sum = 0.0;
  for (int i = lane; i < N; i += 32) {
    const float u = det ? linspace01(i, N) : u_row[i];
    int ind;
    const float s = draw_sample(s_cdf, s_bins, M, u, &ind);
    s_sort[S + i] = s;
    sum += (double)s;
    if (z_samples_row) z_samples_row[i] = s;
    if (inds_row) inds_row[i] = ind;
  }
  __syncwarp();
  if (z_std) {  // torch.std(unbiased=False): population std of the N new samples
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double mean = sum / N;
    double var = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double d = (double)s_sort[S + i] - mean;
      var += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    if (lane == 0) *z_std = (float)sqrt(var / N);
  }
  // torch.sort(cat([z_vals, z_samples])) returns values only, so any correct sort is bit-identical.  When both lists
  // are already non-decreasing (always for the coarse depths; for the new samples whenever u is sorted, i.e. the
  // deterministic test-time case, up to last-ulp effects at bin edges) a merge-path merge does it in ~100
  // instructions per lane; the check below keeps the general bitonic sort as the fallback, so the result never
  // depends on that assumption.
  bool a_sorted = true, b_sorted = true;
  for (int i = lane; i < S + N; i += 32)
    if (i != 0 && i != S && s_sort[i] < s_sort[i - 1]) { if (i < S) a_sorted = false; else b_sorted = false; }
  a_sorted = __all_sync(0xffffffffu, a_sorted);
  b_sorted = __all_sync(0xffffffffu, b_sorted);
  // random u (training): only the N new samples are out of order -> bitonic sort of those (padded with the +inf already
  // behind them to PN = pow2 >= N, when that fits the buffer), then the same merge: 112 instead of 288 compare-exchanges
  // per lane at 64 + 128
  if (a_sorted && !b_sorted) {
    int PN = 1;
    while (PN < N) PN <<= 1;
    if (S + PN <= P) {
      float* Bs = s_sort + S;
      for (int k = 2; k <= PN; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = lane; t < PN; t += 32) {
            const int partner = t ^ j;
            if (partner > t) {
              const float a = Bs[t], b = Bs[partner];
              const bool up = ((t & k) == 0);
              if ((a > b) == up) { Bs[t] = b; Bs[partner] = a; }
            }
          }
          __syncwarp();
        }
      }
      b_sorted = true;
    }
  }
  if (a_sorted && b_sorted) {
    const float* A = s_sort;            // S coarse depths
    const float* B = s_sort + S;        // N new samples
    const int T = S + N, per = (T + 31) / 32;
    const int p0 = min(T, lane * per), p1 = min(T, p0 + per);
    int lo = max(0, p0 - N), hi = min(p0, S);
    while (lo < hi) {                   // number of A elements among the first p0 outputs (A first on ties)
      const int mid = (lo + hi) >> 1;
      if (A[mid] <= B[p0 - mid - 1]) lo = mid + 1; else hi = mid;
    }
    int i = lo, j = p0 - lo;
    float* out = z_out_row;
    for (int p = p0; p < p1; ++p) {
      const bool take_a = (i < S) && (j >= N || A[i] <= B[j]);
      out[p] = take_a ? A[i] : B[j];
      if (take_a) ++i; else ++j;
    }
    return;
  }
  // general case: warp bitonic sort over P = pow2 >= S+N
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < P; t += 32) {
        const int partner = t ^ j;
        if (partner > t) {
          const float a = s_sort[t], b = s_sort[partner];
          const bool up = ((t & k) == 0);
          if ((a > b) == up) { s_sort[t] = b; s_sort[partner] = a; }
        }
      }
      __syncwarp();
    }
  }
  for (int j = lane; j < S + N; j += 32) z_out_row[j] = s_sort[j];
}

__global__ void __launch_bounds__(32 * PW)
hierarchical_z_kernel(int R, int S, int N, int P, const float* __restrict__ z, const float* __restrict__ weights,
                      int det, const float* __restrict__ u_in, float* __restrict__ z_out,
                      float* __restrict__ z_samples, long long* __restrict__ inds, float* __restrict__ z_std) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = blockIdx.x * PW + wib;
  if (r >= R) return;
  hierarchical_z_ray(S, N, P, z + (size_t)r * S, weights + (size_t)r * S, det, u_in ? u_in + (size_t)r * N : nullptr,
                     z_out + (size_t)r * (S + N), z_samples ? z_samples + (size_t)r * N : nullptr,
                     inds ? inds + (size_t)r * N : nullptr, z_std ? z_std + r : nullptr,
                     smem + (size_t)wib * (3 * (S - 1) + P), lane);
}

// Coarse compositing + hierarchical resampling of a ray in ONE kernel (ipynb:451-467): raw2outputs of the coarse pass
// (rgb0 / disp0 / acc0 / depth0), then z_mid, sample_pdf(weights[1:-1]), sort(cat) and z_std with the S coarse weights
// kept in shared memory -- they never go to HBM unless the caller asks for them (`weights` != NULL).  Same device
// functions as the two separate kernels, so every output is bit-identical to raw2outputs_fwd + hierarchical_z.
__global__ void __launch_bounds__(32 * PW)
composite_resample_kernel(int R, int S, int N, int P, const float* __restrict__ raw, int ch, const float* __restrict__ z,
                          const float* __restrict__ rays_d, int ld_d, const float* __restrict__ noise, int white, int det,
                          const float* __restrict__ u_in, float* __restrict__ rgb, float* __restrict__ disp,
                          float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ weights,
                          float* __restrict__ z_out, float* __restrict__ z_samples, float* __restrict__ z_std) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = blockIdx.x * PW + wib;
  if (r >= R) return;
  float* s_wts = smem + (size_t)wib * (S + 3 * (S - 1) + P);
  const float* rr = raw + (size_t)r * S * ch;
  const float* zr = z + (size_t)r * S;
  const float* nr = noise ? noise + (size_t)r * S : nullptr;
  const float nrm = ray_norm(rays_d + (size_t)r * ld_d);
  float c0 = 0.f, c1 = 0.f, c2 = 0.f, dsum = 0.f, asum = 0.f;
  float* wr = weights ? weights + (size_t)r * S : nullptr;
  scan_ray(S, rr, ch, zr, nrm, nr, lane, [&](int i, float alpha, float x, float T) {
    const float w = __fmul_rn(alpha, T);
    s_wts[i] = w;
    if (wr) wr[i] = w;
    c0 += w * sigmoidf_(rr[(size_t)i * ch + 0]);
    c1 += w * sigmoidf_(rr[(size_t)i * ch + 1]);
    c2 += w * sigmoidf_(rr[(size_t)i * ch + 2]);
    dsum += w * zr[i];
    asum += w;
  });
  c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); dsum = warp_sum(dsum); asum = warp_sum(asum);
  if (lane == 0) {
    if (white) {
      const float bg = __fsub_rn(1.0f, asum);
      c0 += bg; c1 += bg; c2 += bg;
    }
    rgb[3 * r + 0] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2;
    if (depth) depth[r] = dsum;
    if (acc) acc[r] = asum;
    if (disp) disp[r] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(dsum, fmaxf(1e-10f, asum)), 1e-10f));
  }
  __syncwarp();
  hierarchical_z_ray(S, N, P, zr, s_wts, det, u_in ? u_in + (size_t)r * N : nullptr, z_out + (size_t)r * (S + N),
                     z_samples ? z_samples + (size_t)r * N : nullptr, nullptr, z_std ? z_std + r : nullptr, s_wts + S, lane);
}

// The coarse pass of the fused inference path: nerf_mlp_kernel<FWD_COMP> left, per ray, the segment records and the
// per-sample partial weights w'_i = alpha_i * (exclusive product inside the 32-sample segment).  This kernel (one warp per
// ray) multiplies in the transmittance in front of each segment (product of the earlier segments' Q, fp64), writes the
// coarse maps rgb0 / disp0 / acc0 / depth0, and runs the z_mid / sample_pdf / sort / z_std block on the finished weights,
// which stay in shared memory.  Coarse depths come from `z` or, when NULL, are recomputed from the rays' near / far.
__global__ void __launch_bounds__(32 * PW)
resample_records_kernel(int R, int S, int N, int P, const float* __restrict__ rays, int rw, const float* __restrict__ z,
                        int lindisp, const float* __restrict__ wpart, const float* __restrict__ rec, int white, int det,
                        const float* __restrict__ u_in, float* __restrict__ rgb, float* __restrict__ disp,
                        float* __restrict__ acc, float* __restrict__ depth, float* __restrict__ z_out,
                        float* __restrict__ z_samples, float* __restrict__ z_std) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int r = blockIdx.x * PW + wib;
  if (r >= R) return;
  float* s_wts = smem + (size_t)wib * (2 * S + 3 * (S - 1) + P);
  float* s_z = s_wts + S;
  const int nseg = S >> 5;
  const float* ray = rays + (size_t)r * rw;
  float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
  if (lane < nseg) {
    const float4* q = reinterpret_cast<const float4*>(rec + ((size_t)r * nseg + lane) * 8);
    ra = __ldg(q);
    rb = __ldg(q + 1);
  }
  double T = 1.0;
  float asum = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, dsum = 0.f;
  const float near_v = __ldg(ray + 6), far_v = __ldg(ray + 7);
  for (int s = 0; s < nseg; ++s) {
    const float qlo = __shfl_sync(0xffffffffu, ra.x, s), qhi = __shfl_sync(0xffffffffu, ra.y, s);
    const float A = __shfl_sync(0xffffffffu, ra.z, s), C0 = __shfl_sync(0xffffffffu, ra.w, s);
    const float C1 = __shfl_sync(0xffffffffu, rb.x, s), C2 = __shfl_sync(0xffffffffu, rb.y, s);
    const float D = __shfl_sync(0xffffffffu, rb.z, s);
    const float Tf = (float)T;
    asum = fmaf(Tf, A, asum); c0 = fmaf(Tf, C0, c0); c1 = fmaf(Tf, C1, c1); c2 = fmaf(Tf, C2, c2); dsum = fmaf(Tf, D, dsum);
    const int i = 32 * s + lane;
    s_wts[i] = __fmul_rn(Tf, __ldg(wpart + (size_t)r * S + i));
    s_z[i] = z ? __ldg(z + (size_t)r * S + i) : z_at(i, S, near_v, far_v, lindisp);
    T *= (double)qlo;
  }
  if (lane == 0) {
    if (white) {
      const float bg = __fsub_rn(1.0f, asum);
      c0 += bg; c1 += bg; c2 += bg;
    }
    rgb[3 * r + 0] = c0; rgb[3 * r + 1] = c1; rgb[3 * r + 2] = c2;
    if (depth) depth[r] = dsum;
    if (acc) acc[r] = asum;
    if (disp) disp[r] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(dsum, fmaxf(1e-10f, asum)), 1e-10f));
  }
  __syncwarp();
  hierarchical_z_ray(S, N, P, s_z, s_wts, det, u_in ? u_in + (size_t)r * N : nullptr, z_out + (size_t)r * (S + N),
                     z_samples ? z_samples + (size_t)r * N : nullptr, nullptr, z_std ? z_std + r : nullptr, s_z + S, lane);
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_resample_records(int R, int S, int N, const float* rays, int ray_width, const float* z, int lindisp,
                                      const float* wpart, const float* rec, int white_bkgd, int det, const float* u,
                                      float* rgb, float* disp, float* acc, float* depth, float* z_out, float* z_samples,
                                      float* z_std, void* stream) {
  NERFB_REQUIRE(R >= 0 && S >= 32 && S % 32 == 0 && S <= 1024 && N >= 1 && rays && ray_width >= 8 && wpart && rec && rgb && z_out,
                "nerfb_resample_records: bad arguments (R=%d S=%d N=%d)", R, S, N);
  NERFB_REQUIRE(det || u, "nerfb_resample_records: det=0 needs u");
  NERFB_REQUIRE(aligned16(rec), "nerfb_resample_records: rec must be 16-byte aligned");
  int P = 1;
  while (P < S + N) P <<= 1;
  const size_t smem = (size_t)PW * (2 * S + 3 * (S - 1) + P) * sizeof(float);
  NERFB_REQUIRE(smem <= 48 * 1024, "nerfb_resample_records: S+N too large (%d+%d)", S, N);
  if (R == 0) return NERFB_OK;
  resample_records_kernel<<<cdiv(R, PW), 32 * PW, smem, as_stream(stream)>>>(R, S, N, P, rays, ray_width, z, lindisp, wpart,
                                                                            rec, white_bkgd, det, u, rgb, disp, acc, depth,
                                                                            z_out, z_samples, z_std);
  return check_launch("resample_records_kernel");
}

extern "C" int nerfb_sample_pdf(int R, int M, int N, const float* bins, int ld_bins, const float* weights,
                                int ld_w, int det, const float* u, float* samples, long long* inds,
                                void* stream) {
  NERFB_REQUIRE(R >= 0 && M >= 2 && N >= 1 && bins && weights && samples && ld_bins >= M && ld_w >= M - 1,
                "nerfb_sample_pdf: bad arguments (R=%d M=%d N=%d)", R, M, N);
  NERFB_REQUIRE(det || u, "nerfb_sample_pdf: det=0 needs u");
  const size_t smem = (size_t)PW * 3 * M * sizeof(float);
  NERFB_REQUIRE(smem <= 48 * 1024, "nerfb_sample_pdf: too many bins (%d)", M);
  if (R == 0) return NERFB_OK;
  sample_pdf_kernel<<<cdiv(R, PW), 32 * PW, smem, as_stream(stream)>>>(R, M, N, bins, ld_bins, weights, ld_w, det, u,
                                                                      samples, inds);
  return check_launch("sample_pdf_kernel");
}

extern "C" int nerfb_hierarchical_z(int R, int S, int N, const float* z, const float* weights, int det,
                                    const float* u, float* z_out, float* z_samples, long long* inds,
                                    float* z_std, void* stream) {
  NERFB_REQUIRE(R >= 0 && S >= 3 && N >= 1 && z && weights && z_out,
                "nerfb_hierarchical_z: bad arguments (R=%d S=%d N=%d)", R, S, N);
  NERFB_REQUIRE(det || u, "nerfb_hierarchical_z: det=0 needs u");
  int P = 1;
  while (P < S + N) P <<= 1;
  const size_t smem = (size_t)PW * (3 * (S - 1) + P) * sizeof(float);
  NERFB_REQUIRE(smem <= 48 * 1024, "nerfb_hierarchical_z: S+N too large (%d+%d)", S, N);
  if (R == 0) return NERFB_OK;
  hierarchical_z_kernel<<<cdiv(R, PW), 32 * PW, smem, as_stream(stream)>>>(R, S, N, P, z, weights, det, u, z_out,
                                                                          z_samples, inds, z_std);
  return check_launch("hierarchical_z_kernel");
}

extern "C" int nerfb_composite_resample(int R, int S, int N, const float* raw, int raw_ch, const float* z,
                                        const float* rays_d, int ld_d, const float* noise, int white_bkgd, int det,
                                        const float* u, float* rgb, float* disp, float* acc, float* depth, float* weights,
                                        float* z_out, float* z_samples, float* z_std, void* stream) {
  NERFB_REQUIRE(R >= 0 && S >= 3 && N >= 1 && raw && z && rays_d && rgb && z_out && raw_ch >= 4 && ld_d >= 3,
                "nerfb_composite_resample: bad arguments (R=%d S=%d N=%d)", R, S, N);
  NERFB_REQUIRE(det || u, "nerfb_composite_resample: det=0 needs u");
  int P = 1;
  while (P < S + N) P <<= 1;
  const size_t smem = (size_t)PW * (S + 3 * (S - 1) + P) * sizeof(float);
  NERFB_REQUIRE(smem <= 48 * 1024, "nerfb_composite_resample: S+N too large (%d+%d)", S, N);
  if (R == 0) return NERFB_OK;
  composite_resample_kernel<<<cdiv(R, PW), 32 * PW, smem, as_stream(stream)>>>(
      R, S, N, P, raw, raw_ch, z, rays_d, ld_d, noise, white_bkgd, det, u, rgb, disp, acc, depth, weights, z_out, z_samples,
      z_std);
  return check_launch("composite_resample_kernel");
}
