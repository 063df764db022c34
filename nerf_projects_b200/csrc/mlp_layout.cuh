// Geometry, packed-weight layout and the training tile stores shared by the fused MLP
// kernels (mlp_tc.cu) and the weight-gradient kernel (wgrad_tc.cu).
#pragma once
#include <cstdint>

namespace nerfb {

// ------------------------------------------------------------ geometry
constexpr int kTileRows = 128;                 // rows per CTA per tile (UMMA M / 2)
constexpr int kSuperRows = 512;                // rows per cluster per iteration (2 CTAs x 2 tiles)
constexpr int kChunkBytes = 16384;             // 128 rows x 64 bf16 (one K chunk of A, or half of a B chunk)
#ifndef NERFB_RING_SLOTS
#define NERFB_RING_SLOTS 4
#endif
constexpr int kNS = NERFB_RING_SLOTS;          // weight ring slots of 16 KB per CTA (4 = one 256x256 layer)
constexpr int kThreads = 384;                  // warps: 0 loader, 1 mma, 2 tmem, 3 relay, 4..11 epilogue
constexpr int kEpiThreads = 256;

// packed weight blob: bf16 chunk images, then the fp32 parameter block
constexpr int kChunkFull = 2 * kChunkBytes;    // N=256 chunk: two CTA halves
constexpr int kChunkViews = kChunkBytes;       // N=128 chunk: two halves of 64 rows
constexpr int kNumFullChunks = 34;             // L0(1) L1-4(16) L5(1+4) L6-7(8) FEAT(4)
constexpr int kNumViewChunks = 5;              // gamma(d) part + 4 feature parts
constexpr int kChunkAreaVD = kNumFullChunks * kChunkFull + kNumViewChunks * kChunkViews;
constexpr int kChunkAreaNoVD = 30 * kChunkFull;
// Bias tiles (forward): a layer's bias is added ON THE TENSOR CORES by one extra K = 16 MMA at the start of each tile's
// block: A operand = a constant 256-byte tile ([8 rows x 16] bf16, columns 0 and 1 = 1.0, read by every 8-row group through
// a descriptor whose row-group stride is 0), B operand = the layer's [N x 16] bf16 tile with column 0 = bf16(bias) and
// column 1 = bf16(bias - column 0) (a 16-bit mantissa: the rounding of a bias is systematic over all samples, so a single
// bf16 is not enough -- measured on the calibrated scenes); no-swizzle K-major (tc05::nosw16_offset), two CTA halves of
// 4 KB.  Tiles in blob order: pts_linears 0..7[, feature, views].
constexpr int kBiasTileHalf = 4096;
constexpr int kBiasTileBytes = 2 * kBiasTileHalf;
constexpr int kNumBiasTilesVD = 10, kNumBiasTilesNoVD = 8;
constexpr int kBiasAreaVD = kNumBiasTilesVD * kBiasTileBytes, kBiasAreaNoVD = kNumBiasTilesNoVD * kBiasTileBytes;
constexpr int kParamOffVD = kChunkAreaVD + kBiasAreaVD, kParamOffNoVD = kChunkAreaNoVD + kBiasAreaNoVD;

// fp32 parameter block (floats)
constexpr int kOffBias = 0;                    // 8 x 256 pts biases
constexpr int kOffBiasFeat = 2048;             // 256
constexpr int kOffBiasViews = 2304;            // 128
constexpr int kOffHeadW = 2432;                // 5 x 256: alpha_linear.weight (row 0) or output_linear.weight
constexpr int kOffRgbW = 3712;                 // 3 x 128
constexpr int kOffHeadB = 4096;                // 8
constexpr int kOffRgbB = 4104;                 // 4
constexpr int kParamFloats = 4108;
constexpr int kParamBytes = kParamFloats * 4;  // 16432

// shared memory map (bytes from the 1024-aligned dynamic smem base)
constexpr int kSmAct = 0;                      // 2 tiles x 4 chunks
constexpr int kSmEnc = 2 * 4 * kChunkBytes;    // 131072: gamma(x) / gamma(d) operand chunk
constexpr int kSmRing = kSmEnc + kChunkBytes;  // 147456
constexpr int kSmParams = kSmRing + kNS * kChunkBytes;  // 212992: fp32 parameter block; its first 8 KB (the pts biases,
                                                        // not needed in shared memory any more) = two bias-tile buffers
constexpr int kSmBiasBuf = kSmParams;
constexpr int kSmOnes = kSmParams + 2 * 4096;           // the constant A operand of the bias MMAs (256 B; feature / views bias floats)
constexpr int kSmSig = kSmParams + kParamBytes;         // 229424: [2 tiles][2 halves][128] fp32
constexpr int kSmBars = kSmSig + 2 * 2 * 128 * 4;       // 231472
constexpr int kNumBars = 2 * kNS + 2 + 2 + 1 + 4;       // w_full, w_empty, act_ready[2], mma_done[2], enc_free, stage_full[2], stage_free[2]
constexpr int kSmTmemSlot = kSmBars + kNumBars * 8;
constexpr int kSmOps = kSmTmemSlot + 16;                 // copy of Sched::ops (u32 each)
constexpr int kSmLoads = kSmOps + 84 * 4;                // copy of Sched::loads (2 x u32 each)
constexpr int kSmemBytes = kSmLoads + 44 * 8;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");

// ---------------------------------------------------- training tile stores
// Training keeps, per 128-sample tile, bf16 images of every layer input/gradient in HBM in
// exactly the shared-memory operand layout (64-column panels of 128 rows x 128 B, 128B swizzle):
// the forward writes them with bulk copies from its activation buffers, the backward chain
// does the same for the gradients, and the wgrad kernel reads both back as MN-major operands.
constexpr int kPanel = kChunkBytes;
constexpr int kFXg = 0;                               // gamma(x)            1 panel
constexpr int kFH = kPanel;                           // h_l, l=0..7         4 panels each
constexpr int kFFeat = kFH + 8 * 4 * kPanel;          // feature             4 panels
constexpr int kFDg = kFFeat + 4 * kPanel;             // gamma(d)            1 panel
constexpr int kFHv = kFDg + kPanel;                   // hv (views layer)    2 panels
constexpr int kFwdTileBytes = kFHv + 2 * kPanel;      // 40 panels = 640 KB
constexpr int kGHv = 0;                               // d hv                2 panels
constexpr int kGFeat = 2 * kPanel;                    // d feature           4 panels
constexpr int kGY = 6 * kPanel;                       // dY_l, l=0..7        4 panels each
constexpr int kBwdTileBytes = kGY + 8 * 4 * kPanel;   // 38 panels = 608 KB
// ReLU masks for the data-gradient chain: one SIGN bit per unit (1 = pre-activation negative = inactive), collected by
// the forward-with-save epilogue; per 128-sample tile [9 layers: h0..h7, hv][128 rows][2 column halves][4 words], word b of
// a half covers its columns 32b..32b+31 with column j at bit 31 - j.  36 KB per tile (the bf16 activations are 640 KB).
constexpr int kMaskLayers = 9;
constexpr int kMaskRowBytes = 32;
constexpr int kMaskTileBytes = kMaskLayers * 128 * kMaskRowBytes;
constexpr int kNumBwdChunks = 34;                     // views^T(2) feat^T(4) L7^T..L1^T(28)
constexpr int kChunkAreaBwd = kNumBwdChunks * kChunkFull;
// *_NV: the network without view directions (output_linear head) -- separate instantiations, so the view-dependent
// training kernels compile exactly as before
// MODE_FWD_COMP: the forward with the alpha compositing of raw2outputs folded in (inference): raw is never written
// MODE_FWD_NV / MODE_FWD_COMP_NV: the inference kernels of the network without view directions, as their own
// instantiations: use_viewdirs is a compile-time constant in every kernel, so none carries the other head's epilogue
enum { MODE_FWD = 0, MODE_FWD_SAVE = 1, MODE_BWD = 2, MODE_FWD_SAVE_NV = 3, MODE_BWD_NV = 4, MODE_FWD_COMP = 5,
       MODE_FWD_NV = 6, MODE_FWD_COMP_NV = 7 };
// compositing records: per ray and 32-sample segment 8 floats = {Q = product of 1-alpha+1e-10 over the segment, unused,
// sum w', sum w' r, sum w' g, sum w' b, sum w' z, unused} with w' = alpha * (exclusive product inside the segment)
constexpr int kRecFloats = 8;

}  // namespace nerfb
