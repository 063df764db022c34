// Image-quality metrics of render_path / calculate_metrics (nerf_helpers.py:21-203): SSIM with a separable
// Gaussian window (zero padding, sigma clamps as in calculate_ssim, nerf_helpers.py:21-111) and the clipped MSE
// behind PSNR (nerf_helpers.py:179-189), for [H,W,C] fp32 images, in ONE pass over the two images:
// a CTA owns a 16x32 pixel tile, stages tile+halo of both images in shared memory (clamped), blurs the five
// moments (x, y, x^2, y^2, xy) along W into shared memory, then along H in registers, evaluates the SSIM formula,
// and writes one (ssim, squared-error) partial per CTA; a second one-CTA kernel adds the partials in a fixed
// order (deterministic means).  HBM-bound: 8 B read per pixel-channel (+4 B when the map is requested).
#include <cstdlib>
#include "common.cuh"

namespace nerfb {

constexpr int kTH = 16, kTW = 32, kMaxF = 33, kThreads = 256;

struct SsimArgs {
  float filt[kMaxF];
  int fs, hw;
  float max_val, c1, c2;
};

__global__ void __launch_bounds__(kThreads) ssim_mse_kernel(int H, int W, int C, const float* __restrict__ a,
                                                            const float* __restrict__ b, SsimArgs p,
                                                            float* __restrict__ ssim_map, double* __restrict__ partials) {
  extern __shared__ float smem[];
  const int hw = p.hw, IH = kTH + 2 * hw, IW = kTW + 2 * hw;
  float* sa = smem;                 // [IH][IW] clamped image 1 (0 outside the image)
  float* sb = sa + IH * IW;         // [IH][IW] image 2
  float* sh = sb + IH * IW;         // [5][IH][kTW] row-blurred moments
  const int y0 = blockIdx.y * kTH, x0 = blockIdx.x * kTW, tid = threadIdx.x;
  double acc_ssim = 0.0, acc_se = 0.0;
  for (int c = 0; c < C; ++c) {
    for (int i = tid; i < IH * IW; i += kThreads) {
      const int r = i / IW, q = i - r * IW, y = y0 + r - hw, x = x0 + q - hw;
      float va = 0.f, vb = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        const size_t o = ((size_t)y * W + x) * C + c;
        va = a[o];
        vb = b[o];
        if (r >= hw && r < hw + kTH && q >= hw && q < hw + kTW) {  // own pixel: squared error of the [0,1]-clipped images
          const float d = fminf(fmaxf(va, 0.f), 1.f) - fminf(fmaxf(vb, 0.f), 1.f);
          acc_se += (double)(d * d);
        }
        va = fminf(fmaxf(va, 0.f), p.max_val);
        vb = fminf(fmaxf(vb, 0.f), p.max_val);
      }
      sa[i] = va;
      sb[i] = vb;
    }
    __syncthreads();
    for (int i = tid; i < IH * kTW; i += kThreads) {
      const int r = i / kTW, q = i - r * kTW;
      const float* ra = sa + r * IW + q;
      const float* rb = sb + r * IW + q;
      float m0 = 0.f, m1 = 0.f, m00 = 0.f, m11 = 0.f, m01 = 0.f;
      for (int k = 0; k < p.fs; ++k) {
        const float w = p.filt[k], u = ra[k], v = rb[k];
        m0 = fmaf(w, u, m0);
        m1 = fmaf(w, v, m1);
        m00 = fmaf(w, u * u, m00);
        m11 = fmaf(w, v * v, m11);
        m01 = fmaf(w, u * v, m01);
      }
      sh[0 * IH * kTW + i] = m0;
      sh[1 * IH * kTW + i] = m1;
      sh[2 * IH * kTW + i] = m00;
      sh[3 * IH * kTW + i] = m11;
      sh[4 * IH * kTW + i] = m01;
    }
    __syncthreads();
    for (int i = tid; i < kTH * kTW; i += kThreads) {
      const int r = i / kTW, q = i - r * kTW, y = y0 + r, x = x0 + q;
      if (y < H && x < W) {
        float m[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const float* col = sh + j * IH * kTW + r * kTW + q;
          float s = 0.f;
          for (int k = 0; k < p.fs; ++k) s = fmaf(p.filt[k], col[k * kTW], s);
          m[j] = s;
        }
        const float mu00 = m[0] * m[0], mu11 = m[1] * m[1], mu01 = m[0] * m[1];
        const float s00 = fmaxf(m[2] - mu00, 0.f), s11 = fmaxf(m[3] - mu11, 0.f);
        float s01 = m[4] - mu01;
        const float lim = fminf(sqrtf(s00 * s11), fabsf(s01));
        s01 = s01 > 0.f ? lim : (s01 < 0.f ? -lim : 0.f);
        const float numer = (2.f * mu01 + p.c1) * (2.f * s01 + p.c2);
        const float denom = (mu00 + mu11 + p.c1) * (s00 + s11 + p.c2);
        const float v = numer / denom;
        acc_ssim += (double)v;
        if (ssim_map) ssim_map[((size_t)c * H + y) * W + x] = v;   // [C,H,W] like the reference's map
      }
    }
    __syncthreads();
  }
  // block reduction in a fixed order: warp shuffles, then warp 0 over the 8 warp sums
  __shared__ double red[2][kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) {
    acc_ssim += __shfl_down_sync(0xffffffffu, acc_ssim, o);
    acc_se += __shfl_down_sync(0xffffffffu, acc_se, o);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = acc_ssim;
    red[1][tid >> 5] = acc_se;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0, e = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      s += red[0][w];
      e += red[1][w];
    }
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    partials[2 * bid] = s;
    partials[2 * bid + 1] = e;
  }
}

// Specialisation for a compile-time window (the reference's default 11 taps): 32x32 pixel tile, every thread produces FOUR
// adjacent outputs per pass from a register window -- the row pass reads 4 aligned float4 per image instead of 4 x FS
// scalars, the column pass FS+3 values per moment instead of 4 x FS -- about 4x fewer shared-memory instructions than the
// generic kernel above (68 -> see profiles/README.md).  Same summation order per output, so the values are identical.
template <int FS>
__global__ void __launch_bounds__(kThreads) ssim_mse_fixed_kernel(int H, int W, int C, const float* __restrict__ a,
                                                                  const float* __restrict__ b, SsimArgs p,
                                                                  float* __restrict__ ssim_map,
                                                                  double* __restrict__ partials) {
  constexpr int TH = 32, HW = FS / 2, IH = TH + 2 * HW, IW = kTW + 2 * HW, IWP = (IW + 3) / 4 * 4, WIN = FS + 3;
  static_assert(4 * (kTW / 4 - 1) + 16 <= IWP, "row window must stay inside the padded row");
  extern __shared__ float smem[];
  float* sa = smem;                   // [IH][IWP]
  float* sb = sa + IH * IWP;          // [IH][IWP]
  float* sh = sb + IH * IWP;          // [5][IH][kTW]
  const int y0 = blockIdx.y * TH, x0 = blockIdx.x * kTW, tid = threadIdx.x;
  double acc_ssim = 0.0, acc_se = 0.0;
  for (int c = 0; c < C; ++c) {
    for (int i = tid; i < IH * IWP; i += kThreads) {
      const int r = i / IWP, q = i - r * IWP, y = y0 + r - HW, x = x0 + q - HW;
      float va = 0.f, vb = 0.f;
      if (q < IW && y >= 0 && y < H && x >= 0 && x < W) {
        const size_t o = ((size_t)y * W + x) * C + c;
        va = a[o];
        vb = b[o];
        if (r >= HW && r < HW + TH && q >= HW && q < HW + kTW) {
          const float d = fminf(fmaxf(va, 0.f), 1.f) - fminf(fmaxf(vb, 0.f), 1.f);
          acc_se += (double)(d * d);
        }
        va = fminf(fmaxf(va, 0.f), p.max_val);
        vb = fminf(fmaxf(vb, 0.f), p.max_val);
      }
      sa[i] = va;
      sb[i] = vb;
    }
    __syncthreads();
    // ---- blur along W: item = (row r, group g of 4 output columns)
    for (int it = tid; it < IH * (kTW / 4); it += kThreads) {
      const int r = it / (kTW / 4), g = it - r * (kTW / 4);
      float u[16], v[16];
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const float4 fa = *reinterpret_cast<const float4*>(sa + r * IWP + 4 * g + 4 * k4);
        const float4 fb = *reinterpret_cast<const float4*>(sb + r * IWP + 4 * g + 4 * k4);
        u[4 * k4] = fa.x; u[4 * k4 + 1] = fa.y; u[4 * k4 + 2] = fa.z; u[4 * k4 + 3] = fa.w;
        v[4 * k4] = fb.x; v[4 * k4 + 1] = fb.y; v[4 * k4 + 2] = fb.z; v[4 * k4 + 3] = fb.w;
      }
      float uu[WIN], vv[WIN], uv[WIN];
#pragma unroll
      for (int k = 0; k < WIN; ++k) { uu[k] = u[k] * u[k]; vv[k] = v[k] * v[k]; uv[k] = u[k] * v[k]; }
      float m[5][4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float m0 = 0.f, m1 = 0.f, m00 = 0.f, m11 = 0.f, m01 = 0.f;
#pragma unroll
        for (int k = 0; k < FS; ++k) {
          const float w = p.filt[k];
          m0 = fmaf(w, u[o + k], m0);
          m1 = fmaf(w, v[o + k], m1);
          m00 = fmaf(w, uu[o + k], m00);
          m11 = fmaf(w, vv[o + k], m11);
          m01 = fmaf(w, uv[o + k], m01);
        }
        m[0][o] = m0; m[1][o] = m1; m[2][o] = m00; m[3][o] = m11; m[4][o] = m01;
      }
#pragma unroll
      for (int j = 0; j < 5; ++j)
        *reinterpret_cast<float4*>(sh + (j * IH + r) * kTW + 4 * g) = make_float4(m[j][0], m[j][1], m[j][2], m[j][3]);
    }
    __syncthreads();
    // ---- blur along H + SSIM: item = (group rg of 4 output rows, column q)
    for (int it = tid; it < (TH / 4) * kTW; it += kThreads) {
      const int rg = it / kTW, q = it - rg * kTW, x = x0 + q;
      float mo[5][4];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        float col[WIN];
#pragma unroll
        for (int k = 0; k < WIN; ++k) col[k] = sh[(j * IH + 4 * rg + k) * kTW + q];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < FS; ++k) s = fmaf(p.filt[k], col[o + k], s);
          mo[j][o] = s;
        }
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int y = y0 + 4 * rg + o;
        if (y < H && x < W) {
          const float mu00 = mo[0][o] * mo[0][o], mu11 = mo[1][o] * mo[1][o], mu01 = mo[0][o] * mo[1][o];
          const float s00 = fmaxf(mo[2][o] - mu00, 0.f), s11 = fmaxf(mo[3][o] - mu11, 0.f);
          float s01 = mo[4][o] - mu01;
          const float lim = fminf(sqrtf(s00 * s11), fabsf(s01));
          s01 = s01 > 0.f ? lim : (s01 < 0.f ? -lim : 0.f);
          const float numer = (2.f * mu01 + p.c1) * (2.f * s01 + p.c2);
          const float denom = (mu00 + mu11 + p.c1) * (s00 + s11 + p.c2);
          const float val = numer / denom;
          acc_ssim += (double)val;
          if (ssim_map) ssim_map[((size_t)c * H + y) * W + x] = val;
        }
      }
    }
    __syncthreads();
  }
  __shared__ double red[2][kThreads / 32];
  for (int o = 16; o > 0; o >>= 1) {
    acc_ssim += __shfl_down_sync(0xffffffffu, acc_ssim, o);
    acc_se += __shfl_down_sync(0xffffffffu, acc_se, o);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = acc_ssim;
    red[1][tid >> 5] = acc_se;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0, e = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      s += red[0][w];
      e += red[1][w];
    }
    const int bid = blockIdx.y * gridDim.x + blockIdx.x;
    partials[2 * bid] = s;
    partials[2 * bid + 1] = e;
  }
}

__global__ void __launch_bounds__(kThreads) metrics_finalize_kernel(int nblocks, double count,
                                                                   const double* __restrict__ partials,
                                                                   float* __restrict__ out) {
  __shared__ double red[2][kThreads];
  double s = 0.0, e = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += kThreads) {
    s += partials[2 * i];
    e += partials[2 * i + 1];
  }
  red[0][threadIdx.x] = s;
  red[1][threadIdx.x] = e;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float mse = (float)(red[1][0] / count);
    out[0] = mse;                                          // img2mse of the clipped images
    out[1] = -10.0f * logf(mse) / logf(10.0f);             // mse2psnr (nerf_helpers.py:12-14)
    out[2] = (float)(red[0][0] / count);                   // mean SSIM
  }
}

}  // namespace nerfb

extern "C" size_t nerfb_image_metrics_workspace_bytes(int H, int W) {
  using namespace nerfb;
  if (H <= 0 || W <= 0) return 0;
  return (size_t)cdiv(H, kTH) * cdiv(W, kTW) * 2 * sizeof(double);
}

extern "C" int nerfb_image_metrics(int H, int W, int C, const float* img1, const float* img2, float max_val,
                                   int filter_size, float filter_sigma, float k1, float k2, float* ssim_map,
                                   void* workspace, float* out3, void* stream) {
  using namespace nerfb;
  NERFB_REQUIRE(H > 0 && W > 0 && C > 0, "nerfb_image_metrics: bad shape %dx%dx%d", H, W, C);
  NERFB_REQUIRE(filter_size >= 1 && filter_size <= kMaxF && (filter_size & 1) == 1,
                "nerfb_image_metrics: filter_size must be odd and <= %d (got %d)", kMaxF, filter_size);
  NERFB_REQUIRE(img1 && img2 && workspace && out3, "nerfb_image_metrics: null pointer");
  SsimArgs p{};
  p.fs = filter_size;
  p.hw = filter_size / 2;
  // 1-D Gaussian of calculate_ssim (nerf_helpers.py:70-74), fp32 like the reference's tensor ops
  const float shift = (float)(2 * p.hw - filter_size + 1) / 2.0f;
  float sum = 0.f;
  for (int i = 0; i < filter_size; ++i) {
    const float t = ((float)(i - p.hw) + shift) / filter_sigma;
    p.filt[i] = expf(-0.5f * (t * t));
    sum += p.filt[i];
  }
  for (int i = 0; i < filter_size; ++i) p.filt[i] /= sum;
  p.max_val = max_val;
  p.c1 = (k1 * max_val) * (k1 * max_val);
  p.c2 = (k2 * max_val) * (k2 * max_val);
  const int IH = kTH + 2 * p.hw, IW = kTW + 2 * p.hw;
  const size_t smem = (size_t)(2 * IH * IW + 5 * IH * kTW) * sizeof(float);
  if (smem > 48 * 1024)   // only for windows wider than 27 taps
    NERFB_CUDA(cudaFuncSetAttribute(ssim_mse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(W, kTW), cdiv(H, kTH));
  static const bool force_generic = getenv("NERFB_SSIM_GENERIC") != nullptr;      // A/B measurements only
  if (filter_size == 11 && !force_generic) {     // the reference's default window: register-window kernel, 32x32 tiles
    constexpr int FS = 11, TH = 32, IHf = TH + 2 * (FS / 2), IWPf = (kTW + 2 * (FS / 2) + 3) / 4 * 4;
    const size_t smem_f = (size_t)(2 * IHf * IWPf + 5 * IHf * kTW) * sizeof(float);
    grid = dim3(cdiv(W, kTW), cdiv(H, TH));
    ssim_mse_fixed_kernel<FS><<<grid, kThreads, smem_f, as_stream(stream)>>>(H, W, C, img1, img2, p, ssim_map,
                                                                            reinterpret_cast<double*>(workspace));
  } else {
    ssim_mse_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(H, W, C, img1, img2, p, ssim_map,
                                                                 reinterpret_cast<double*>(workspace));
  }
  int rc = check_launch("ssim_mse_kernel");
  if (rc) return rc;
  metrics_finalize_kernel<<<1, kThreads, 0, as_stream(stream)>>>((int)(grid.x * grid.y), (double)H * W * C,
                                                                 reinterpret_cast<const double*>(workspace), out3);
  return check_launch("metrics_finalize_kernel");
}
