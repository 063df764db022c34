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
  float pdf[4];                       // a lane's pdf terms (IEEE divisions: computed once when the lane owns <= 4)
  if (seg <= 4) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      pdf[t] = (j0 + t < j1) ? __fdiv_rn(s_w[j0 + t], S) : 0.f;
      if (j0 + t < j1) local += (double)pdf[t];
    }
  } else {
    for (int j = j0; j < j1; ++j) local += (double)__fdiv_rn(s_w[j], S);
  }
  double incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double q = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += q;
  }
  double run = incl - local;
  if (lane == 0) s_cdf[0] = 0.f;
  if (seg <= 4) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (j0 + t < j1) {
        run += (double)pdf[t];
        s_cdf[j0 + t + 1] = (float)run;
      }
  } else {
    for (int j = j0; j < j1; ++j) {
      run += (double)__fdiv_rn(s_w[j], S);
      s_cdf[j + 1] = (float)run;
    }
  }
  __syncwarp();
}

// shared-memory accesses by 32-bit shared-window address: with generic `float*` scratch pointers the compiler
// rematerialises the window base (S2R SR_CgaCtaId + LEA) around every access inside the search loops
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// searchsorted(right=True): #{j : cdf[j] <= u}.  The cdf is non-decreasing (fp64 running sum of non-negative terms,
// rounded monotonically), so the count is found by descending power-of-two steps -- one load, one compare, one predicated
// add per step, no data-dependent branch (the lo/hi loop cost 12-14 instructions per step and was 28 % of the resampling
// kernel's instructions).  TOP = largest power of two <= M when M + 1 is a power of two (every step stays in range, fully
// unrolled: M = 63 for 64 coarse samples), else 0: `top` at run time with a range check.
template <int TOP>
__device__ __forceinline__ int count_le(uint32_t cdf_sa, int M, int top, float u) {
  if (TOP > 0) {
    uint32_t p = cdf_sa;                       // address of cdf[lo]
#pragma unroll
    for (int step = TOP; step > 0; step >>= 1)
      if (lds_f32(p + 4u * (uint32_t)(step - 1)) <= u) p += 4u * (uint32_t)step;
    return (int)((p - cdf_sa) >> 2);
  }
  int lo = 0;
  for (int step = top; step > 0; step >>= 1) {
    const int nxt = lo + step;
    const float c = (nxt <= M) ? lds_f32(cdf_sa + 4u * (uint32_t)(nxt - 1)) : CUDART_INF_F;
    if (c <= u) lo = nxt;
  }
  return lo;
}

template <int TOP>
__device__ __forceinline__ float draw_sample(uint32_t cdf_sa, uint32_t bins_sa, int M, int top, float u, int* ind_out) {
  const int lo = count_le<TOP>(cdf_sa, M, top, u);
  *ind_out = lo;
  const int below = max(lo - 1, 0), above = min(lo, M - 1);
  const float c0 = lds_f32(cdf_sa + 4u * (uint32_t)below), c1 = lds_f32(cdf_sa + 4u * (uint32_t)above);
  float denom = __fsub_rn(c1, c0);
  if (denom < 1e-5f) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, c0), denom);
  const float b0 = lds_f32(bins_sa + 4u * (uint32_t)below), b1 = lds_f32(bins_sa + 4u * (uint32_t)above);
  return __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
}

// draws samples i = lane, lane + 32, ... < N of one ray; f(i, sample, ind) consumes them
template <typename F>
__device__ __forceinline__ void draw_all(const float* s_cdf, const float* s_bins, int M, int N, int det, const float* u_row,
                                         int lane, F&& f) {
  const uint32_t cdf_sa = (uint32_t)__cvta_generic_to_shared(s_cdf), bins_sa = (uint32_t)__cvta_generic_to_shared(s_bins);
  const int top = 1 << (31 - __clz(M));
  // torch.linspace(0, 1, N)[i] (common.cuh: linspace01) with the IEEE division of its step done once per ray
  const float lstep = (N > 1) ? __fdiv_rn(1.0f, (float)(N - 1)) : 0.f;
  auto u_at = [&](int i) -> float {
    if (!det) return u_row[i];
    if (N <= 1) return 0.f;
    return (i < N / 2) ? __fmul_rn(lstep, (float)i) : __fmaf_rn(-lstep, (float)(N - 1 - i), 1.0f);
  };
  if (M == 63) {
    for (int i = lane; i < N; i += 32) {
      int ind;
      const float s = draw_sample<32>(cdf_sa, bins_sa, M, top, u_at(i), &ind);
      f(i, s, ind);
    }
  } else {
    for (int i = lane; i < N; i += 32) {
      int ind;
      const float s = draw_sample<0>(cdf_sa, bins_sa, M, top, u_at(i), &ind);
      f(i, s, ind);
    }
  }
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
  draw_all(s_cdf, s_bins, M, N, det, det ? nullptr : u_in + (size_t)r * N, lane, [&](int i, float s, int ind) {
    samples[(size_t)r * N + i] = s;
    if (inds) inds[(size_t)r * N + i] = ind;
  });
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
  {
    const uint32_t w_sa = (uint32_t)__cvta_generic_to_shared(s_w), bins_sa = (uint32_t)__cvta_generic_to_shared(s_bins);
    const uint32_t sort_sa = (uint32_t)__cvta_generic_to_shared(s_sort);
    for (int j = lane; j < S; j += 32) {
      const float zj = zr[j];
      sts_f32(sort_sa + 4u * (uint32_t)j, zj);
      if (j < M) sts_f32(bins_sa + 4u * (uint32_t)j, __fmul_rn(0.5f, __fadd_rn(zr[j + 1], zj)));
      if (j < M - 1) sts_f32(w_sa + 4u * (uint32_t)j, __fadd_rn(wr[j + 1], 1e-5f));
    }
  }
  __syncwarp();
  build_cdf(s_w, s_cdf, M, lane);
  double sum = 0.0;
  draw_all(s_cdf, s_bins, M, N, det, u_row, lane, [&](int i, float s, int ind) {
    s_sort[S + i] = s;
    sum += (double)s;
    if (z_samples_row) z_samples_row[i] = s;
    if (inds_row) inds_row[i] = ind;
  });
  __syncwarp();
  if (z_std) {  // torch.std(unbiased=False): population std of the N new samples
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    // N a power of two (128 by default): the two fp64 divisions are exact scalings (same bits, ~50 instructions fewer)
    const bool pow2 = (N & (N - 1)) == 0;
    const double invN = __longlong_as_double((long long)(1023 - (31 - __clz(N))) << 52);
    const double mean = pow2 ? sum * invN : sum / N;
    double var = 0.0;
    for (int i = lane; i < N; i += 32) {
      const double d = (double)s_sort[S + i] - mean;
      var += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    if (lane == 0) *z_std = (float)sqrt(pow2 ? var * invN : var / N);
  }
  // torch.sort(cat([z_vals, z_samples])) returns values only, so any correct sort is bit-identical.  The coarse depths
  // are non-decreasing, and so are the new samples whenever u is (the deterministic test-time case, up to last-ulp
  // effects at bin edges); with random u (training) only the N new samples are out of order and are bitonic-sorted first
  // (112 instead of 288 compare-exchanges per lane at 64 + 128).  A merge-path merge of the two lists then does it in ~100
  // instructions per lane.  Nothing is ASSUMED: the merge is verified afterwards -- every lane's outputs non-decreasing
  // and continuing the previous lane's, and every lane ending exactly where the next one started (so the output is a
  // permutation of the input) -- and any violation falls through to the general bitonic sort.
  const uint32_t A_sa = (uint32_t)__cvta_generic_to_shared(s_sort), B_sa = A_sa + 4u * (uint32_t)S;
  bool try_merge = true;
  if (!det) {
    int PN = 1;
    while (PN < N) PN <<= 1;
    if (S + PN <= P) {
      for (int j = S + N + lane; j < S + PN; j += 32) s_sort[j] = CUDART_INF_F;
      __syncwarp();
      for (int k = 2; k <= PN; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int t = lane; t < PN; t += 32) {
            const int partner = t ^ j;
            if (partner > t) {
              const float a = lds_f32(B_sa + 4u * (uint32_t)t), b = lds_f32(B_sa + 4u * (uint32_t)partner);
              const bool up = ((t & k) == 0);
              if ((a > b) == up) { sts_f32(B_sa + 4u * (uint32_t)t, b); sts_f32(B_sa + 4u * (uint32_t)partner, a); }
            }
          }
          __syncwarp();
        }
      }
    } else {
      try_merge = false;
    }
  }
  if (try_merge) {
    const int T = S + N, per = (T + 31) / 32;
    const int p0 = min(T, lane * per), p1 = min(T, p0 + per);
    int lo = max(0, p0 - N), hi = min(p0, S);
    while (lo < hi) {                   // number of A elements among the first p0 outputs (A first on ties)
      const int mid = (lo + hi) >> 1;
      if (lds_f32(A_sa + 4u * (uint32_t)mid) <= lds_f32(B_sa + 4u * (uint32_t)(p0 - mid - 1))) lo = mid + 1; else hi = mid;
    }
    const int i0 = lo, j0 = p0 - lo;
    // the heads of the two lists stay in registers; one shared-memory load per output
    uint32_t pa = A_sa + 4u * (uint32_t)i0, pb = B_sa + 4u * (uint32_t)j0;
    const uint32_t ea = B_sa, eb = B_sa + 4u * (uint32_t)N;
    float a = (pa < ea) ? lds_f32(pa) : CUDART_INF_F, b = (pb < eb) ? lds_f32(pb) : CUDART_INF_F;
    float prev = -CUDART_INF_F, first = CUDART_INF_F;
    bool ok = true;
    float* out = z_out_row;
    for (int p = p0; p < p1; ++p) {
      const bool take_a = (pa < ea) && (pb >= eb || a <= b);
      const float v = take_a ? a : b;
      out[p] = v;
      ok = ok && !(v < prev);
      if (p == p0) first = v;
      prev = v;
      if (take_a) { pa += 4u; a = (pa < ea) ? lds_f32(pa) : CUDART_INF_F; } else { pb += 4u; b = (pb < eb) ? lds_f32(pb) : CUDART_INF_F; }
    }
    // seams: my end = the next lane's start (in both lists), and my first output >= the previous lane's last
    const int iend = (int)((pa - A_sa) >> 2), jend = (int)((pb - B_sa) >> 2);
    const int inext = __shfl_down_sync(0xffffffffu, i0, 1), jnext = __shfl_down_sync(0xffffffffu, j0, 1);
    const float last_prev = __shfl_up_sync(0xffffffffu, prev, 1);
    if (lane < 31) ok = ok && iend == inext && jend == jnext;
    else ok = ok && iend == S && jend == N;
    if (lane > 0 && p1 > p0) ok = ok && !(first < last_prev);
    if (__all_sync(0xffffffffu, ok)) return;
  }
  // general case: warp bitonic sort over P = pow2 >= S+N
  for (int j = S + N + lane; j < P; j += 32) s_sort[j] = CUDART_INF_F;
  __syncwarp();
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
