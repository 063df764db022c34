/*
 * nerf_b200.h -- C ABI of libnerf_b200.so: the B200 (sm_100a) implementation
 * of the original-NeRF render/train hot path of isaacchunn/nerf-projects.
 *
 * The reference has no FFI of its own for this path (it is eager PyTorch called
 * from notebook cells), so this ABI is the inner boundary that the Python
 * drop-in (`nerf_projects_b200`, same names/signatures as the notebook
 * functions) binds with ctypes.  Each entry point names the reference code it
 * replaces; "ipynb:N" is the raw JSON line number of
 * /root/reference/nerf/nerf.ipynb, other paths are relative to
 * /root/reference/nerf/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (caller-allocated, e.g. torch tensor
 *     storage) unless the parameter name starts with `h_`;
 *   - tensors are dense row-major fp32 unless a leading dimension `ld*`
 *     (in elements) is given;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it,
 *     no call synchronises the device;
 *   - return value: 0 on success, negative on error; the message of the last
 *     error on the calling thread is returned by nerfb_last_error();
 *   - no C++ exceptions cross this boundary, no torch types appear in it.
 */
#ifndef NERF_B200_H_
#define NERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERFB_ABI_VERSION 1

enum {
  NERFB_OK = 0,
  NERFB_ERR_ARG = -1,     /* invalid argument / unsupported shape */
  NERFB_ERR_CUDA = -2,    /* CUDA runtime error (launch, attribute, ...) */
  NERFB_ERR_ARCH = -3     /* device is not sm_100 */
};

int nerfb_abi_version(void);
const char* nerfb_last_error(void);
/* sm count / compute capability of the current device */
int nerfb_device_info(int* h_sm_count, int* h_cc_major, int* h_cc_minor);

/* ------------------------------------------------------------------ rays */

/* get_rays (nerf_helpers.py:222-296): one pinhole ray per pixel, integer pixel
 * coordinates, dirs=((i-cx)/fx, -(j-cy)/fy, -1) rotated by c2w[:3,:3].
 * c2w: [3,4] row-major.  rays_o, rays_d: [H*W,3]. */
int nerfb_get_rays(int H, int W, float fx, float fy, float cx, float cy,
                   const float* c2w, float* rays_o, float* rays_d, void* stream);

/* ndc_rays (nerf_helpers.py:311-369), n rays, in place allowed. */
int nerfb_ndc_rays(int n, int H, int W, double focal, double near_plane,
                   const float* rays_o, const float* rays_d,
                   float* out_o, float* out_d, void* stream);

/* Ray packing of render() (ipynb:598-629): viewdirs = d/|d| taken from
 * `view_d` (the pre-NDC directions; may equal rays_d), optional NDC warp with
 * near=1 and focal, then rows [o(3) d(3) near far (viewdir(3))].
 * out: [n, 8] or [n, 11] when use_viewdirs. */
int nerfb_pack_rays(int n, const float* rays_o, const float* rays_d, const float* view_d,
                    int use_viewdirs, int ndc, int H, int W, double focal,
                    float near_v, float far_v, float* out, void* stream);

/* Coarse depths of render_rays (ipynb:418-444): t=linspace(0,1,S);
 * z = near(1-t)+far t, or 1/(1/near(1-t)+1/far t) when lindisp; when t_rand
 * ([R,S], uniform [0,1)) is not NULL the stratified jitter is applied.
 * rays: packed rows of width ray_width (near/far at columns 6,7). z: [R,S]. */
int nerfb_coarse_z(int R, int S, const float* rays, int ray_width, int lindisp,
                   const float* t_rand, float* z, void* stream);

/* Ray-batch producer of train() (ipynb:1167-1255: `use_batching` shuffled rays_rgb, or per-image
 * get_rays + np.random.choice of pixel coordinates): rays of n SELECTED pixels generated from
 * (pose, intrinsics, pixel) with get_rays' exact arithmetic, and their target colours gathered.
 * sel[i] = image*H*W + row*W + col when img0 < 0, else the pixel id (row*W + col) inside image img0.
 * poses: [n_images] blocks of pose_stride floats holding a row-major [3,4] (12) or [4,4] (16) c2w.
 * images: [n_images,H,W,3] or NULL (then target is not written). */
int nerfb_gather_ray_batch(int n, int H, int W, float fx, float fy, float cx, float cy,
                           const float* poses, int pose_stride, int n_images, const float* images,
                           const long long* sel, int img0, float* rays_o, float* rays_d, float* target,
                           void* stream);

/* ------------------------------------------------- positional encoding */

/* Embedder.embed (embedder.py:36-80): out[i, col0 : col0+3+6L] =
 * [x, sin(2^0 x), cos(2^0 x), ..., cos(2^(L-1) x)]; L < 0 -> identity (3). */
int nerfb_embed(int n, const float* x, int ldx, int L, float* out, int ld_out, int col0,
                void* stream);

/* run_network's input stage (ipynb:824-843) fused with pts = o + d z
 * (ipynb:447): writes gamma(pts) for all R*S samples into
 * enc[:, col_pts : col_pts+3+6L] and, when Lv != NERFB_NO_VIEWS, gamma(viewdir)
 * broadcast to every sample of the ray into enc_dir[:, col_dir : ...].
 * If `pts` is not NULL it is used instead of o + d z ([R*S,3]); `viewdirs`
 * ([R,3]) overrides columns 8..10 of `rays` when not NULL. */
#define NERFB_NO_VIEWS (-1000)
int nerfb_embed_samples(int R, int S, const float* rays, int ray_width, const float* z,
                        const float* pts, const float* viewdirs, int L, int Lv,
                        float* enc, int ld_enc, int col_pts,
                        float* enc_dir, int ld_dir, int col_dir, void* stream);

/* ------------------------------------------- fp32 validation-mode layers */

/* nn.Linear (+ReLU) forward, nerf.py:70-109: Y[M,N] = act(X[M,K] W[N,K]^T + b). */
int nerfb_linear_fwd(int M, int N, int K, const float* X, int ldx, const float* W, int ldw,
                     const float* b, float* Y, int ldy, int relu, void* stream);
/* dX[M,K] (= or +=) dY[M,N] W[N,K]; then, if relu_src != NULL, dX *= (relu_src > 0)
 * (relu_src is the post-ReLU activation that fed this layer). */
int nerfb_linear_bwd_data(int M, int N, int K, const float* dY, int lddy, const float* W, int ldw,
                          float* dX, int lddx, const float* relu_src, int ld_relu,
                          int accumulate, void* stream);
/* dW[N,K] += dY[M,N]^T X[M,K];  db[N] += column sums of dY (db may be NULL). */
int nerfb_linear_bwd_weight(int M, int N, int K, const float* dY, int lddy, const float* X, int ldx,
                            float* dW, int lddw, float* db, void* stream);

/* ------------------------------------------------------------ compositing */

/* raw2outputs forward (ipynb:254-349).  raw: [R,S,raw_ch] (raw_ch >= 4, channels
 * 0..2 rgb logits, 3 sigma); z: [R,S]; rays_d: rows of stride ld_d (|d| scales
 * the intervals; last interval 1e10); noise: NULL or [R,S] added to sigma.
 * Outputs (any may be NULL except rgb): rgb[R,3] disp[R] acc[R] weights[R,S] depth[R]. */
int nerfb_raw2outputs_fwd(int R, int S, const float* raw, int raw_ch, const float* z,
                          const float* rays_d, int ld_d, const float* noise, int white_bkgd,
                          float* rgb, float* disp, float* acc, float* weights, float* depth,
                          void* stream);
/* Backward of the above w.r.t. raw.  Upstream grads may be NULL (treated as 0).
 * d_raw: [R,S,raw_ch] (channels >= 4 are written as 0). */
int nerfb_raw2outputs_bwd(int R, int S, const float* raw, int raw_ch, const float* z,
                          const float* rays_d, int ld_d, const float* noise, int white_bkgd,
                          const float* d_rgb, const float* d_disp, const float* d_acc,
                          const float* d_weights, const float* d_depth,
                          float* d_raw, void* stream);

/* --------------------------------------------------- hierarchical sampling */

/* sample_pdf (nerf_helpers.py:372-439), one warp per ray, reproducing the
 * reference's CPU fp32 arithmetic bit for bit (see oracle/sample_pdf_np.py):
 * bins[R,M] (stride ld_bins), weights[R,M-1] (stride ld_w).  det!=0 ->
 * u=linspace(0,1,N); else u must be given ([R,N]).  samples[R,N];
 * inds (int64 [R,N]) may be NULL. */
int nerfb_sample_pdf(int R, int M, int N, const float* bins, int ld_bins,
                     const float* weights, int ld_w, int det, const float* u,
                     float* samples, long long* inds, void* stream);

/* The hierarchical step of render_rays (ipynb:459-467, 486) in one kernel:
 * z_mid = .5(z[1:]+z[:-1]); z_samples = sample_pdf(z_mid, weights[1:-1], N);
 * z_out = sort(cat(z, z_samples)); z_std = std(z_samples, unbiased=False).
 * z, weights: [R,S]; z_out: [R,S+N]; z_samples/inds/z_std may be NULL. */
int nerfb_hierarchical_z(int R, int S, int N, const float* z, const float* weights,
                         int det, const float* u, float* z_out, float* z_samples,
                         long long* inds, float* z_std, void* stream);

/* Coarse compositing + hierarchical resampling in one launch (render_rays, ipynb:451-467): raw2outputs of the coarse
 * pass (rgb/disp/acc/depth = the reference's rgb0/disp0/acc0 + depth0), then nerfb_hierarchical_z on its weights, which
 * stay in shared memory (written to `weights` [R,S] only when not NULL).  Bit-identical to the two separate calls. */
int nerfb_composite_resample(int R, int S, int N, const float* raw, int raw_ch, const float* z,
                             const float* rays_d, int ld_d, const float* noise, int white_bkgd, int det,
                             const float* u, float* rgb, float* disp, float* acc, float* depth,
                             float* weights, float* z_out, float* z_samples, float* z_std, void* stream);

/* ------------------------------------------- fused tensor-core MLP (bf16) */

/* Canonical NeRF MLP only: D=8, W=256, skips=[4], multires 10 (63 ch),
 * multires_views 4 (27 ch) or no view directions. */
size_t nerfb_packed_weights_bytes(int use_viewdirs);
/* Packs one network's fp32 parameters (nerf.py:32-55 state_dict order, see
 * NERFB_PARAM_* below) into the bf16 UMMA shared-memory images the fused
 * kernel streams with bulk async copies.  h_params: host array of device
 * pointers, NERFB_NUM_PARAMS entries; entries not used by the architecture
 * may be NULL.  out_ch = rows of output_linear (4 or 5) when !use_viewdirs. */
enum {
  NERFB_P_PTS0_W = 0, NERFB_P_PTS0_B, /* ... pts_linears.i at 2i, 2i+1 for i<8 */
  NERFB_P_VIEWS_W = 16, NERFB_P_VIEWS_B,
  NERFB_P_FEATURE_W, NERFB_P_FEATURE_B,
  NERFB_P_ALPHA_W, NERFB_P_ALPHA_B,
  NERFB_P_RGB_W, NERFB_P_RGB_B,
  NERFB_P_OUTPUT_W, NERFB_P_OUTPUT_B,
  NERFB_NUM_PARAMS
};
int nerfb_pack_weights(const float* const* h_params, int use_viewdirs, int out_ch,
                       void* packed, void* stream);
/* run_network + NeRF.forward (ipynb:790-855, nerf.py:57-111) for all R*S
 * samples: encoding in registers, 8x256 MLP + heads on tcgen05 tensor cores
 * (bf16 operands, fp32 TMEM accumulators, fp32 density head), activations
 * never leave the SM.  Sample positions are o + d z from `rays`/`z`, or the
 * explicit `pts` ([R*S,3]) when not NULL; `viewdirs` as in nerfb_embed_samples.
 * raw: [R*S, 4] = [rgb logits, sigma] with view directions, else [R*S, out_ch]
 * (output_linear, out_ch = 4 or 5). */
int nerfb_mlp_fwd_bf16(int R, int S, const float* rays, int ray_width, const float* z,
                       const float* pts, const float* viewdirs, int use_viewdirs, int out_ch,
                       const void* packed, float* raw, void* stream);
/* Inference: the same network pass WITH the alpha compositing of raw2outputs (ipynb:254-349) folded in -- the
 * per-sample [rgb, sigma] never leave the SM (north_star (1)).  S must be a multiple of 32: each warp-sized segment of a
 * ray is composited where it was evaluated (fp32 prefix product of 1 - alpha + 1e-10 inside the segment; bf16 mode) and leaves a
 * 32-byte record in `rec` [R][S/32][8 floats] = {Q = product over the segment, unused, sums of w', w' r, w' g, w' b,
 * w' z, unused}; the transmittance in front of a segment is the product of the earlier records' Q.  z [R,S] or NULL: coarse
 * depths near*(1-t)+far*t (lindisp: the disparity-linear form) are computed in-kernel from columns 6/7 of `rays`
 * (ipynb:418-425).  wpart [R,S] or NULL: the per-sample partial weights w' (coarse pass: nerfb_resample_records needs
 * them).  A ray is finished by nerfb_composite_finish (final pass) or nerfb_resample_records (coarse pass). */
int nerfb_mlp_fwd_composite_bf16(int R, int S, const float* rays, int ray_width, const float* z, int lindisp,
                                 int use_viewdirs, int out_ch, const void* packed, float* rec, float* wpart, void* stream);
/* rgb [R,3], disp / acc / depth [R] (NULL to skip) of raw2outputs from the nseg = S/32 records of every ray. */
int nerfb_composite_finish(int R, int nseg, const float* rec, int white_bkgd, float* rgb, float* disp, float* acc,
                           float* depth, void* stream);
/* Second launch of the fused coarse pass (ipynb:451-467): finishes the coarse maps rgb0 / disp0 / acc0 / depth0 from
 * the records, rebuilds the coarse weights (kept in shared memory) and runs z_mid / sample_pdf(weights[1:-1]) /
 * sort(cat) / z_std exactly like nerfb_hierarchical_z.  z NULL: coarse depths recomputed from the rays as above.
 * z_out [R,S+N]; z_samples [R,N] or NULL; z_std [R] or NULL. */
int nerfb_resample_records(int R, int S, int N, const float* rays, int ray_width, const float* z, int lindisp,
                           const float* wpart, const float* rec, int white_bkgd, int det, const float* u, float* rgb,
                           float* disp, float* acc, float* depth, float* z_out, float* z_samples, float* z_std,
                           void* stream);
/* ---- training: parameter gradients of the MLP on tensor cores (replaces the autograd graph of
 * NeRF.forward, nerf.py:57-111, under loss.backward(), ipynb:1274).  These three are for use_viewdirs networks;
 * the *_nv_* variants below cover use_viewdirs=False.
 * Three launches per network pass and ray chunk:
 *   1. nerfb_mlp_fwd_save_bf16: the fused forward, which also bulk-stores every layer input as
 *      bf16 operand images into `acts` and one ReLU sign bit per hidden unit into `masks`;
 *   2. nerfb_mlp_bwd_chain_bf16: data gradients d_raw -> d_hv -> d_feature -> dY_7..dY_0 through
 *      the transposed weights (nerfb_pack_weights_bwd), ReLU masks from `masks`, gradients
 *      bulk-stored into `gstore`;
 *   3. nerfb_mlp_wgrad_bf16: dW += dY^T X, db += colsum(dY) for all 12 Linear layers, fp32
 *      accumulation in TMEM, added (red.global.add.f32) into the caller's fp32 gradient tensors.
 * nerfb_train_store_bytes(nrows, 0|1|2): bytes of `acts` / `gstore` / `masks` for nrows samples (whole
 * 512-row super-tiles).  d_raw: [nrows,4] = d loss / d raw.  d_raw_padded: the same buffer but
 * allocated (and zero beyond nrows) up to a multiple of 512 rows.  h_grads: host array of
 * NERFB_NUM_PARAMS device pointers (NERFB_P_* order, state_dict shapes), accumulated into. */
size_t nerfb_train_store_bytes(long long nrows, int which);
int nerfb_mlp_fwd_save_bf16(int R, int S, const float* rays, int ray_width, const float* z,
                            const float* pts, const float* viewdirs, const void* packed, float* raw,
                            void* acts, void* masks, void* stream);
size_t nerfb_packed_weights_bwd_bytes(void);
int nerfb_pack_weights_bwd(const float* const* h_params, void* packed_bwd, void* stream);
int nerfb_mlp_bwd_chain_bf16(long long nrows, const float* d_raw, const void* packed_bwd,
                             const void* masks, void* gstore, void* stream);
int nerfb_mlp_wgrad_bf16(long long nrows, const void* acts, const void* gstore,
                         const float* d_raw_padded, float* const* h_grads, void* stream);
/* nerfb_mlp_bwd_chain_bf16 + nerfb_mlp_wgrad_bf16 in ONE cooperative launch: `chain_sms` SMs (even) run the data-gradient
 * chain, the others the weight-gradient jobs, which read each 128-sample tile's dY panels as soon as the chain has stored
 * them -- from L2, not from HBM -- and overlap their tensor work with the chain's.  flags: [ntiles * 20] ints (ntiles =
 * 4 * ceil(nrows / 512)), zero when allocated; epoch: a value not used with this flags buffer before (1, 2, 3, ...).
 * Needs at least chain_sms / 2 super-tiles of 512 samples. */
int nerfb_mlp_bwd_fused_bf16(long long nrows, const float* d_raw_padded, const void* packed_bwd, const void* masks,
                             const void* acts, void* gstore, float* const* h_grads, int* flags, int epoch, int chain_sms,
                             void* stream);
/* The same three steps for the network WITHOUT view directions (use_viewdirs=False: output_linear 256 -> out_ch on
 * h7, nerf.py:50-55, 108-110): the forward-with-save stops after layer 7 (h7 is the last stored layer input, raw =
 * [nrows, out_ch]); the chain starts from d_h7 = (d_raw W_out) * [h7 > 0] computed on CUDA cores and runs W7^T..W1^T;
 * the weight-gradient launch has 10 jobs, output_linear's gradient being a CUDA-core reduction over the h7 tiles.
 * d_raw is always [nrows,4] (channels 0..3); with out_ch = 5 the gradient of channel 4 is passed separately as
 * d_raw4 [nrows] (d_raw4_padded: zero-padded to a multiple of 512 rows like d_raw_padded), else NULL.
 * `acts` / `masks` / `gstore` have the sizes nerfb_train_store_bytes reports (same tile layout, view-layer panels unused). */
int nerfb_mlp_fwd_save_nv_bf16(int R, int S, const float* rays, int ray_width, const float* z,
                               const float* pts, int out_ch, const void* packed, float* raw,
                               void* acts, void* masks, void* stream);
int nerfb_pack_weights_bwd_nv(const float* const* h_params, int out_ch, void* packed_bwd, void* stream);
int nerfb_mlp_bwd_chain_nv_bf16(long long nrows, const float* d_raw, const float* d_raw4,
                                const void* packed_bwd, const void* masks, void* gstore, void* stream);
int nerfb_mlp_wgrad_nv_bf16(long long nrows, int out_ch, const void* acts, const void* gstore,
                            const float* d_raw_padded, const float* d_raw4_padded,
                            float* const* h_grads, void* stream);
/* Self-test of the tcgen05 building blocks (smem descriptors, 128B swizzle,
 * instruction descriptor, TMEM load layout, 2-CTA pairing):
 * C[M,N] = A[M,K] B[N,K]^T with bf16-rounded inputs and fp32 accumulation,
 * M = 128*cta_group (cta_group 1 or 2), K multiple of 64 (<= 256), N multiple
 * of 32 (<= 256).  Used by tests only. */
int nerfb_umma_selftest(int cta_group, int N, int K, const float* A, const float* B, float* C,
                        void* stream);
/* Same for MN-major operands (the wgrad form): C[128,N] = At^T Bt, At [K,128], Bt [K,N] row-major,
 * K multiple of 16 (<= 128), N multiple of 64 (<= 256), single CTA. */
int nerfb_umma_selftest_mn(int N, int K, const float* At, const float* Bt, float* C, void* stream);
/* bias-by-MMA form of the fused MLP (2-CTA pair): C [256,N] = A' Bb^T + A Bw^T, A [256,64], Bw [N,64] as
 * 128B-swizzled K-major chunks, Bb [N,16] as a no-swizzle K-major tile read by ONE K=16 MMA whose A operand A' is either
 * A[:, 48:64] (A8 NULL) or the [8,16] tile A8 repeated for every 8 rows (stride between 8-row groups = 0: a 256-byte
 * constant operand); N = 128 or 256. */
int nerfb_umma_selftest_bias(int N, const float* A, const float* Bw, const float* Bb, const float* A8, float* C,
                             void* stream);
/* Test hook (host only, no GPU): static schedule of the fused MLP kernel, which = 0 forward with view directions,
 * 1 forward without, 2 data-gradient chain, 3 data-gradient chain without view directions.  h_loads: 2 ints per load {blob offset, bytes per CTA}; h_ops: 5 ints per
 * op {load, tile, A source chunk (4 = encoding buffer), flags (1 first, 2 last, 4 release slot, 8 release enc,
 * 16 N=128), block length on the first op of a block}.  Arrays must hold 44 loads / 84 ops. */
int nerfb_debug_schedule(int which, int* h_nloads, int* h_nops, int* h_nlayers, int* h_loads, int* h_ops,
                         int* h_ring_slots);
/* Probe (tools/probe_bulk.py): cycles per chunk of the 1-D bulk-copy path, per CTA, ring of `depth`. */
int nerfb_bulk_probe(const void* src, int src_bytes, int nchunks, int chunk_bytes, int pieces, int depth,
                     int nblocks, long long* out_cycles, void* stream);
/* Probe (tools/probe_contention.py): bulk-copy stream concurrent with UMMAs (mode 0 none, 1 SS, 2 TS) and
 * sts_warps warps of shared-memory stores; out[4*block + {0,1,2,3}] = cycles/chunk, cycles/MMA, #MMAs, #stores. */
int nerfb_contention_probe(const void* src, int src_bytes, int nchunks, int mode, int sts_warps, int nblocks,
                           long long* out, void* stream);
/* Debug: when dev_buf != NULL, block 0 of the fused MLP kernel records a clock64 timeline
 * (8 regions x 4096 entries: loader, MMA issuer, epilogue, relay/observer of the leader CTA, then of the peer;
 * entry = tag<<44 | cycles).
 * Pass NULL to switch tracing off. */
int nerfb_set_trace(long long* dev_buf);
/* Host pointer to the 8 x u64 watchdog record the tensor-core kernels write
 * before trapping when a barrier wait times out ([0]=0xDEAD when fired). */
const unsigned long long* nerfb_watchdog_record(void);

/* ---------------------------------------------------- keyed random draws */

/* Random draws of render_rays keyed by (seed, GLOBAL ray index) -- SURVEY.md 8(e).
 * Replaces torch.rand / torch.randn of ipynb:428-444 (t_rand), nerf_helpers.py:404 (u) and
 * ipynb:312-325 (density noise), whose values depend on chunking and sharding.
 * Value = f(seed, ray_offset + row, column, stream_id): Philox4x32-10 with
 * counter (ray_lo, ray_hi, column/4, stream_id), key (seed_lo, seed_hi), word column%4;
 * uniform = (word>>8) * 2^-24 in [0,1); normal = Box-Muller on word pairs, times `scale`. */
enum { NERFB_RNG_T_RAND = 0, NERFB_RNG_U = 1, NERFB_RNG_NOISE_COARSE = 2, NERFB_RNG_NOISE_FINE = 3 };
int nerfb_keyed_rng(unsigned long long seed, long long ray_offset, int R, int S, int stream_id,
                    int normal, float scale, float* out /* [R,S] */, void* stream);
/* All draws of one render_rays call in ONE launch; NULL outputs are skipped.
 * t_rand [R,N_samples], u [R,N_importance] uniform; noise_coarse [R,N_samples],
 * noise_fine [R,N_samples+N_importance] = N(0,1) * noise_std. */
int nerfb_render_rays_rng(unsigned long long seed, long long ray_offset, int R, int N_samples,
                          int N_importance, float noise_std, float* t_rand, float* u,
                          float* noise_coarse, float* noise_fine, void* stream);

/* -------------------------------------------------------- image metrics */

/* calculate_metrics / calculate_ssim (nerf_helpers.py:21-111, 157-203) for two [H,W,C] fp32 images in one
 * pass: out3 = { img2mse of the [0,1]-clipped images, mse2psnr of it, mean SSIM }.  SSIM: images clamped to
 * [0,max_val], separable Gaussian window (filter_size odd <= 33, zero padding), variance clamps and constants
 * of the reference.  ssim_map: optional [C,H,W] per-pixel map (the reference's return_map layout) or NULL.
 * workspace: nerfb_image_metrics_workspace_bytes(H,W) bytes, 8-byte aligned.  Deterministic (fixed-order sums). */
size_t nerfb_image_metrics_workspace_bytes(int H, int W);
int nerfb_image_metrics(int H, int W, int C, const float* img1, const float* img2, float max_val,
                        int filter_size, float filter_sigma, float k1, float k2, float* ssim_map,
                        void* workspace, float* out3, void* stream);

/* ------------------------------------------------------------- optimiser */

/* torch.optim.Adam step (ipynb:905, 1275) for up to NERFB_ADAM_MAX tensors in
 * one launch: p -= lr * mhat / (sqrt(vhat) + eps), with g scaled by grad_scale
 * first.  h_* are host arrays of device pointers / sizes. */
#define NERFB_ADAM_MAX 64
int nerfb_adam_step(int n_tensors, float* const* h_p, const float* const* h_g,
                    float* const* h_m, float* const* h_v, const long long* h_numel,
                    float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERF_B200_H_ */
