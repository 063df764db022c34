// Counter-based random draws of render_rays keyed by (seed, GLOBAL ray index): SURVEY.md 8(e).
// The reference draws t_rand (ipynb:428-444), u (nerf_helpers.py:404) and the density noise (ipynb:312-325)
// from torch's global generator, so its draws depend on chunking and on how rays are sharded.  Here every
// value is a pure function of (seed, ray, sample, stream): Philox4x32-10 (Salmon et al., SC'11) with
//   counter = (ray_lo, ray_hi, sample/4, stream_id), key = (seed_lo, seed_hi), word = sample%4.
// A render of rays [a,b) on one GPU and of [a,c) + [c,b) on two produce bit-identical draws.
// One launch fills up to four tensors (all the draws of one render_rays call).
#include "common.cuh"

namespace nerfb {

struct RngFill {
  float* out;   // [R, S]
  int S;        // row length
  int groups;   // ceil(S/4)
  int stream;   // stream id (counter word 3)
  int normal;   // 0: uniform [0,1); 1: N(0,1) * scale
  float scale;
};
struct RngArgs {
  RngFill f[4];
  int n;
  long long group_end[4];  // prefix sums of R * groups
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float scale, float& n0, float& n1) {
  const float u1 = (float)((a >> 8) + 1u) * 5.9604644775390625e-8f;  // (0,1]
  const float r = sqrtf(-2.0f * logf(u1)) * scale;
  float s, c;
  sincospif(2.0f * u01(b), &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__global__ void __launch_bounds__(256) keyed_rng_kernel(RngArgs a, uint2 key, unsigned long long ray0, long long total) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int i = 0;
    while (i + 1 < a.n && t >= a.group_end[i]) ++i;
    const RngFill f = a.f[i];
    const long long local = t - (i ? a.group_end[i - 1] : 0);
    const long long ray = local / f.groups;
    const int g = (int)(local - ray * f.groups);
    const unsigned long long gray = ray0 + (unsigned long long)ray;
    const uint4 w = philox4x32_10(make_uint4((uint32_t)gray, (uint32_t)(gray >> 32), (uint32_t)g, (uint32_t)f.stream), key);
    float v[4];
    if (f.normal) {
      box_muller(w.x, w.y, f.scale, v[0], v[1]);
      box_muller(w.z, w.w, f.scale, v[2], v[3]);
    } else {
      v[0] = u01(w.x); v[1] = u01(w.y); v[2] = u01(w.z); v[3] = u01(w.w);
    }
    float* row = f.out + ray * f.S + 4 * g;
    const int left = f.S - 4 * g;
    if (left >= 4 && (f.S & 3) == 0) {
      *reinterpret_cast<float4*>(row) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < left) row[j] = v[j];
    }
  }
}

static int launch(RngArgs& a, unsigned long long seed, long long ray_offset, int R, void* stream) {
  long long total = 0;
  for (int i = 0; i < a.n; ++i) {
    NERFB_REQUIRE(aligned16(a.f[i].out), "keyed rng: output %d is not 16-byte aligned", i);
    a.f[i].groups = (a.f[i].S + 3) / 4;
    total += (long long)R * a.f[i].groups;
    a.group_end[i] = total;
  }
  if (total == 0) return NERFB_OK;
  const long long want = (total + 255) / 256;
  const int grid = (int)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
  keyed_rng_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                                                        (unsigned long long)ray_offset, total);
  return check_launch("keyed_rng_kernel");
}

}  // namespace nerfb

extern "C" int nerfb_keyed_rng(unsigned long long seed, long long ray_offset, int R, int S, int stream_id, int normal,
                               float scale, float* out, void* stream) {
  using namespace nerfb;
  NERFB_REQUIRE(R >= 0 && S >= 0 && ray_offset >= 0 && out != nullptr, "nerfb_keyed_rng: bad shape R=%d S=%d", R, S);
  RngArgs a{};
  a.n = 1;
  a.f[0] = RngFill{out, S, 0, stream_id, normal ? 1 : 0, scale};
  return launch(a, seed, ray_offset, R, stream);
}

extern "C" int nerfb_render_rays_rng(unsigned long long seed, long long ray_offset, int R, int N_samples, int N_importance,
                                     float noise_std, float* t_rand, float* u, float* noise_coarse, float* noise_fine,
                                     void* stream) {
  using namespace nerfb;
  NERFB_REQUIRE(R >= 0 && N_samples > 0 && N_importance >= 0 && ray_offset >= 0, "nerfb_render_rays_rng: bad shape");
  RngArgs a{};
  if (t_rand) a.f[a.n++] = RngFill{t_rand, N_samples, 0, NERFB_RNG_T_RAND, 0, 1.f};
  if (u) a.f[a.n++] = RngFill{u, N_importance, 0, NERFB_RNG_U, 0, 1.f};
  if (noise_coarse) a.f[a.n++] = RngFill{noise_coarse, N_samples, 0, NERFB_RNG_NOISE_COARSE, 1, noise_std};
  if (noise_fine) a.f[a.n++] = RngFill{noise_fine, N_samples + N_importance, 0, NERFB_RNG_NOISE_FINE, 1, noise_std};
  return launch(a, seed, ray_offset, R, stream);
}
