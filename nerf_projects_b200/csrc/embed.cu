// Positional encoding gamma(): Embedder.embed (embedder.py:36-80) and the input
// stage of run_network (ipynb:824-843) fused with pts = o + d*z (ipynb:447).
// fp32 validation path: accurate sincosf, one thread per (sample, coordinate).
#include "common.cuh"

namespace nerfb {

// columns: [x(3) | sin f0 (3) | cos f0 (3) | sin f1 (3) | ...], f_k = 2^k
__device__ __forceinline__ void encode_coord(float x, int L, float* __restrict__ row, int c) {
  row[c] = x;
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    float s, co;
    sincosf(__fmul_rn(x, f), &s, &co);
    row[3 + 6 * k + c] = s;
    row[3 + 6 * k + 3 + c] = co;
    f *= 2.0f;
  }
}

__global__ void embed_kernel(int n, const float* __restrict__ x, int ldx, int L, float* __restrict__ out,
                             int ld_out, int col0) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3LL * n) return;
  const int i = (int)(t / 3), c = (int)(t % 3);
  encode_coord(x[(size_t)i * ldx + c], L < 0 ? 0 : L, out + (size_t)i * ld_out + col0, c);
}

__global__ void embed_samples_kernel(int R, int S, const float* __restrict__ rays, int rw,
                                     const float* __restrict__ z, const float* __restrict__ pts,
                                     const float* __restrict__ viewdirs, int L, int Lv, int has_views,
                                     float* __restrict__ enc, int ld_enc, int col_pts,
                                     float* __restrict__ enc_dir, int ld_dir, int col_dir) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)R * S;
  if (t >= 3 * n) return;
  const long long row = t / 3;
  const int c = (int)(t % 3);
  const int r = (int)(row / S);
  float p;
  if (pts != nullptr) {
    p = pts[row * 3 + c];
  } else {
    const float* ray = rays + (size_t)r * rw;
    p = __fadd_rn(__ldg(ray + c), __fmul_rn(__ldg(ray + 3 + c), z[row]));
  }
  encode_coord(p, L < 0 ? 0 : L, enc + (size_t)row * ld_enc + col_pts, c);
  if (has_views) {
    const float v = viewdirs != nullptr ? __ldg(viewdirs + (size_t)r * 3 + c) : __ldg(rays + (size_t)r * rw + 8 + c);
    encode_coord(v, Lv < 0 ? 0 : Lv, enc_dir + (size_t)row * ld_dir + col_dir, c);
  }
}

}  // namespace nerfb

using namespace nerfb;

extern "C" int nerfb_embed(int n, const float* x, int ldx, int L, float* out, int ld_out, int col0, void* stream) {
  NERFB_REQUIRE(n >= 0 && x && out && ldx >= 3 && L <= 16, "nerfb_embed: bad arguments");
  if (n == 0) return NERFB_OK;
  embed_kernel<<<cdiv(3LL * n, 256), 256, 0, as_stream(stream)>>>(n, x, ldx, L, out, ld_out, col0);
  return check_launch("embed_kernel");
}

extern "C" int nerfb_embed_samples(int R, int S, const float* rays, int ray_width, const float* z,
                                   const float* pts, const float* viewdirs, int L, int Lv, float* enc,
                                   int ld_enc, int col_pts, float* enc_dir, int ld_dir, int col_dir,
                                   void* stream) {
  const int has_views = (Lv != NERFB_NO_VIEWS);
  NERFB_REQUIRE(R >= 0 && S > 0 && enc, "nerfb_embed_samples: bad arguments");
  NERFB_REQUIRE(pts || (rays && z && ray_width >= 8), "nerfb_embed_samples: need pts or rays+z");
  NERFB_REQUIRE(!has_views || (enc_dir && (viewdirs || (rays && ray_width >= 11))),
                "nerfb_embed_samples: view directions requested but not provided");
  if (R == 0) return NERFB_OK;
  embed_samples_kernel<<<cdiv(3LL * R * S, 256), 256, 0, as_stream(stream)>>>(
      R, S, rays, ray_width, z, pts, viewdirs, L, Lv, has_views, enc, ld_enc, col_pts, enc_dir, ld_dir, col_dir);
  return check_launch("embed_samples_kernel");
}
