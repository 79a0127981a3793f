/*
 * rdeic_b200.h — C ABI of librdeic_b200.so, the sm_100a kernel library underneath the
 * RDEIC relay-diffusion decode path.
 *
 * The reference (ShreyasBhaktharam/RDEIC) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md §8b); every entry point below therefore names the *Python call site* in the
 * reference whose arithmetic it replaces.  The Python drop-in classes in rdeic_b200/ keep the
 * reference signatures and call these symbols through ctypes.
 *
 * Conventions
 *   - every pointer is a CUDA *device* pointer unless the name ends in _host;
 *   - the caller owns all memory, including workspaces (size queries are provided);
 *   - `stream` is a cudaStream_t passed as void*; the library never synchronises the device,
 *     never allocates tensor memory and keeps no mutable global state except the per-thread
 *     error string and cached function attributes, so it is re-entrant across streams and
 *     capturable into CUDA graphs;
 *   - return value: 0 = ok, non-zero = error; the message is in rdeic_last_error().
 *   - activations inside the network kernels are NHWC bf16 ("pixel-major"); the entropy
 *     front end and the sampler updates work on the reference's NCHW fp32 tensors directly.
 */
#ifndef RDEIC_B200_H
#define RDEIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* rdeic_stream_t;

/* ---- library ------------------------------------------------------------------------- */
const char* rdeic_last_error(void);
int rdeic_abi_version(void);

/* ---- entropy-model front end: integer / bit-exact kernels ------------------------------
 * Tensors are NCHW fp32 exactly as the reference passes them. */

/* utils/ckbd.py:35-45  ckbd_anchor / ckbd_nonanchor.  which: 0 = anchor, 1 = non-anchor. */
int rdeic_ckbd_mask(const float* y, float* out, int B, int C, int H, int W, int which,
                    rdeic_stream_t stream);
/* utils/ckbd.py:6-24  ckbd_split (both halves in one pass). */
int rdeic_ckbd_split(const float* y, float* anchor, float* nonanchor, int B, int C, int H,
                     int W, rdeic_stream_t stream);
/* utils/ckbd.py:26-33  ckbd_merge = anchor + nonanchor. */
int rdeic_ckbd_merge(const float* anchor, const float* nonanchor, float* out, int64_t numel,
                     rdeic_stream_t stream);
/* utils/ckbd.py:47-59  ckbd_{anchor,nonanchor}_sequeeze: [B,C,H,W] -> [B,C,H,W/2]. */
int rdeic_ckbd_squeeze(const float* y, float* out, int B, int C, int H, int W, int which,
                       rdeic_stream_t stream);
/* utils/ckbd.py:61-73  ckbd_{anchor,nonanchor}_unsequeeze: [B,C,H,Wh] -> [B,C,H,2*Wh].
 * `out` must be 8-byte aligned; Wh % 4 == 0 (or % 8) with 16- (32-) byte aligned pointers selects the vector paths. */
int rdeic_ckbd_unsqueeze(const float* sq, float* out, int B, int C, int H, int Wh, int which,
                         rdeic_stream_t stream);
/* compressai 1.2.4 EntropyModel.quantize(x, "symbols", means) as called at
 * utils/ckbd.py:82,93: sym = int32(rint(x - means)); means may be NULL. */
int rdeic_quantize_symbols(const float* x, const float* means, int32_t* symbols, int64_t numel,
                           rdeic_stream_t stream);
/* utils/ckbd.py:85,96,104,113: x_hat = float(sym) + means. */
int rdeic_dequantize(const int32_t* symbols, const float* means, float* out, int64_t numel,
                     rdeic_stream_t stream);
/* compressai 1.2.4 GaussianConditional.build_indexes as called at utils/ckbd.py:81,92,102,111:
 * s = max(scale, lower_bound); idx = (levels-1) - #{k < levels-1 : s <= table[k]}.
 * `table` is the fp32 scale table produced by utils/func.py:10-13 (device pointer).
 * `scales` and `indexes` must be 16-byte aligned; 32-byte alignment selects the 256-bit path. */
int rdeic_build_indexes(const float* scales, const float* table, int levels, float lower_bound,
                        int32_t* indexes, int64_t numel, rdeic_stream_t stream);
/* Fused decode-side hand-off (utils/ckbd.py:99-115 minus the rANS call): squeeze scales and
 * means of one checkerboard phase and build the CDF indexes in one pass.
 * scales/means [B,C,H,W] -> means_sq [B,C,H,W/2] fp32, indexes [B,C,H,W/2] int32. */
int rdeic_ckbd_squeeze_indexes(const float* scales, const float* means, const float* table,
                               int levels, float lower_bound, float* means_sq,
                               int32_t* indexes, int B, int C, int H, int W, int which,
                               rdeic_stream_t stream);
/* Fused encode-side (utils/ckbd.py:76-97 minus the rANS call): squeeze y/scales/means,
 * build indexes, quantise symbols, and write y_hat = unsqueeze(sym + means). */
int rdeic_ckbd_encode_phase(const float* y, const float* scales, const float* means,
                            const float* table, int levels, float lower_bound,
                            int32_t* symbols, int32_t* indexes, float* y_hat, int B, int C,
                            int H, int W, int which, rdeic_stream_t stream);
/* Fused decode-side second half (utils/ckbd.py:104-105,113-114): y_hat = unsqueeze(float(sym)
 * + means_sq);  symbols/means_sq [B,C,H,Wh] -> y_hat [B,C,H,2*Wh]. */
int rdeic_ckbd_decode_phase(const int32_t* symbols, const float* means_sq, float* y_hat, int B,
                            int C, int H, int Wh, int which, rdeic_stream_t stream);
/* model/compression_modules.py:309-331 VectorQuantiser.quant: z [B,D,H,W] fp32,
 * codebook [K,D] fp32 -> idx [B,H,W] int64 (first minimum of |z|^2+|e|^2-2 z.e), zq [B,D,H,W]. */
int rdeic_vq_quant(const float* z, const float* codebook, int64_t* indices, float* zq, int B,
                   int D, int HW, int K, rdeic_stream_t stream);
/* model/compression_modules.py:333-338 get_codebook_entry: idx [B,H,W] -> [B,D,H,W]. */
int rdeic_vq_lookup(const int64_t* indices, const float* codebook, float* out, int B, int D,
                    int HW, int K, rdeic_stream_t stream);

/* ---- relay sampler updates (NCHW fp32, elementwise, unfused fp32 rounding order) -------- */

/* ldm/models/diffusion/ddpm.py:357-360 q_sample: out = a*x0 + b*noise. */
int rdeic_q_sample(const float* x0, const float* noise, float* out, int64_t numel, float a,
                   float b, rdeic_stream_t stream);
/* model/spaced_sampler_relay.py:369-384 (p_sample_spaced, cond_fn None):
 * pred = r*x - rm1*eps; mean = c1*pred + c2*x; out = mean + sigma*noise (sigma = 0 at index 0).
 * eps_uncond != NULL applies the guidance combine of :281-283 first:
 * eps = eps_uncond + scale*(eps - eps_uncond). */
int rdeic_relay_update(const float* x, const float* eps, const float* eps_uncond,
                       float guidance_scale, const float* noise, float* out, int64_t numel,
                       float r, float rm1, float c1, float c2, float sigma,
                       rdeic_stream_t stream);
/* model/ddim_sampler_relay.py:203-231 (p_sample_ddim): pred_x0 = (x - s1m*e)/sqrt_at;
 * out = sqrt_aprev*pred_x0 + dir*e + sigma*noise.  pred_x0_out may be NULL. */
int rdeic_ddim_update(const float* x, const float* eps, const float* eps_uncond,
                      float guidance_scale, const float* noise, float* out, float* pred_x0_out,
                      int64_t numel, float sqrt_one_minus_at, float sqrt_at, float sqrt_aprev,
                      float dir_coef, float sigma, rdeic_stream_t stream);

/* ---- layout / glue kernels ------------------------------------------------------------ */

/* NCHW fp32 -> NHWC bf16 into channel window [c_off, c_off+C) of a [B,H,W,ldc] tensor. */
int rdeic_nchw_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, int ldc,
                            int c_off, rdeic_stream_t stream);
/* NHWC (bf16 or fp32) [B,H,W,ldc] channels [0,C) -> NCHW fp32. */
int rdeic_nhwc_to_nchw_f32(const void* src, int src_is_f32, float* dst, int B, int C, int H,
                           int W, int ldc, rdeic_stream_t stream);
/* fp32 -> bf16 copy (weights repack), numel elements. */
int rdeic_f32_to_bf16(const float* src, void* dst, int64_t numel, rdeic_stream_t stream);
/* ldm/modules/diffusionmodules/util.py:161-181 timestep_embedding (+ optional SiLU is not
 * applied here): t [B] int64 -> [B,dim] bf16 (cos | sin). */
int rdeic_timestep_embedding(const int64_t* t, void* out_bf16, int B, int dim, float max_period,
                             rdeic_stream_t stream);
/* y = silu(x) on bf16 or fp32 input, bf16 output. */
int rdeic_silu_bf16(const void* x, int x_is_f32, void* out, int64_t numel,
                    rdeic_stream_t stream);
/* ldm/modules/attention.py:49-56 GEGLU: in [rows, 2F] (value | gate) -> out [rows, F]. */
int rdeic_geglu(const void* in_bf16, void* out_bf16, int64_t rows, int F,
                rdeic_stream_t stream);
/* nearest x2 upsample NHWC bf16 (openaimodel.py:106-113, model.py:63-67). */
int rdeic_upsample2x_nhwc(const void* in, void* out, int B, int H, int W, int C,
                          rdeic_stream_t stream);
/* nn.PixelShuffle(2) of the sub-pixel convs (model/layers/conv.py:7-10) on NHWC bf16:
 * in [B,H,W,4C] with channels ordered (i, j, c) -> out [B,2H,2W,C]. */
int rdeic_pixel_shuffle2_nhwc(const void* in, void* out, int B, int H, int W, int C,
                              rdeic_stream_t stream);
/* im2col for the stride-2 3x3 convs: in NHWC [B,H,W,C] -> out [B*(H/2)*(W/2), 9*Cp] with
 * Cp = C rounded up to 64.  pad_lo = 1: symmetric padding 1 (openaimodel.py:150-152 Downsample,
 * model/layers/res_blk.py:17 conv3x3 stride 2); pad_lo = 0: the VAE encoder's bottom/right-only
 * padding (ldm/modules/diffusionmodules/model.py:82-84 `pad = (0,1,0,1)`). */
int rdeic_im2col_3x3_s2(const void* in, void* out, int B, int H, int W, int C, int pad_lo,
                        rdeic_stream_t stream);
/* row softmax of fp32/bf16 logits with scale, bf16 output: [rows, n]. */
int rdeic_softmax_rows(const void* in, int in_is_f32, void* out_bf16, int64_t rows, int n,
                       float scale, rdeic_stream_t stream);
/* batched transpose bf16 [batch, R, C] -> [batch, C, R]. */
int rdeic_transpose_bf16(const void* in, void* out, int batch, int R, int C,
                         rdeic_stream_t stream);
/* inference.py:85-87: NHWC (fp32) [B,H,W,ldc] rgb in [-1,1] -> uint8 HWC [B,H,W,3]
 * via ((x+1)/2).clamp(0,1)*255 truncated. */
int rdeic_image_to_u8(const float* in, uint8_t* out, int64_t pixels, int ldc,
                      rdeic_stream_t stream);

/* ABI 5.  src fp32 [rows, C] (row stride ld_src) -> two bf16 channel windows of dst (row stride ld_dst):
 * hi = bf16(x) at column off_hi, lo = bf16(x - hi) at off_lo.  A GEMM over [hi | lo] with duplicated weight columns
 * reads x to ~16 mantissa bits: used for the tensors entering the network (x of rdeic.py:176, guide_hint :180,
 * context :172, timestep_embedding util.py:161), whose rounding would otherwise reach the output through every
 * skip connection. */
int rdeic_split_bf16_hilo(const float* src, int64_t rows, int C, int64_t ld_src, void* dst, int64_t ld_dst,
                          int off_hi, int off_lo, rdeic_stream_t stream);

/* ABI 5.  Blend T overlapping decoded uint8 HWC tiles [T, th, tw, 3] (tile t's top-left pixel at origin_yx[2t],
 * origin_yx[2t+1]) into one uint8 image [H, W, 3]: weighted mean with a separable linear ramp over `overlap` pixels
 * at every tile edge.  Tiling large frames is this framework's own behaviour (BASELINE config 4; the reference has
 * only the dead fold/unfold helpers of ldm/models/diffusion/ddpm.py:724); 2 * overlap <= th, tw. */
int rdeic_blend_tiles_u8(const uint8_t* tiles, const int32_t* origin_yx, int T, int th, int tw, int overlap,
                         uint8_t* out, int H, int W, rdeic_stream_t stream);

/* ---- normalisation (HBM-bound) --------------------------------------------------------- */

/* GroupNorm(+SiLU) over NHWC bf16 (or, with in_is_f32, the fp32 master copy of a residual
 * stream), bf16 output, fp32 statistics (util.py:209-226 GroupNorm32,
 * model.py:48-49 Normalize, rdeic.py:473-485).  The input may be the channel concat of two
 * tensors x1 [B,HW,C1] and x2 [B,HW,C2] (openaimodel.py:804 `th.cat([h, hs.pop()])`) without
 * materialising it; x2 may be NULL with C2 = 0.  out [B,HW,C1+C2] bf16.
 * workspace: rdeic_groupnorm_workspace_bytes(B, HW, C) bytes. */
int64_t rdeic_groupnorm_workspace_bytes(int B, int64_t HW, int C);
/* Same, with the statistics taken from the producers' epilogues (rdeic_conv_params.stats_out of the
 * GEMMs that wrote x1 / x2: [B*HW/32][C1] and [B*HW/32][C2] (sum, sumsq) pairs) instead of a pass
 * over the tensors; HW must be a multiple of 32. */
int rdeic_groupnorm_from_stats(const void* x1, int C1, const float* stats1, const void* x2, int C2,
                               const float* stats2, int in_is_f32, const float* gamma,
                               const float* beta, void* out, int B, int64_t HW, int groups,
                               float eps, int silu, void* workspace, rdeic_stream_t stream);
/* VAE decoder tail in one kernel: norm_out -> swish -> conv_out (3x3, pad 1, C = 128 -> n_out <= 4)
 * -> optionally the caller's uint8 post-process (ldm/modules/diffusionmodules/model.py:683-686,
 * inference.py:85-87).  x [B,H,W,C] bf16 NHWC (pre-norm), stats = the slab statistics its producer
 * emitted (rdeic_conv_params.stats_out), w_packed bf16 [9 taps][8 (n_out zero-padded)][C],
 * bias fp32 [n_out].  Outputs (either or both): out_f32 [B,H,W,ldo] fp32 (columns >= n_out zeroed),
 * out_u8 [B,H,W,3] ((x+1)/2 clamp *255 truncated; needs n_out == 3).
 * workspace: rdeic_groupnorm_workspace_bytes(B, H*W, C) bytes. */
int rdeic_gn_silu_conv3x3_tail(const void* x, const float* stats, const float* gamma, const float* beta,
                               const void* w_packed, const float* bias, int n_out, float* out_f32, int ldo,
                               uint8_t* out_u8, int B, int H, int W, int C, int groups, float eps,
                               void* workspace, rdeic_stream_t stream);
/* ABI 5.  1 when rdeic_groupnorm_nhwc runs this problem as ONE kernel (one CTA per sample and group, no workspace
 * traffic, the group's elements stay in registers between the statistics and the apply pass): tensors of at most 4 M
 * elements whose groups have 8 or more (an even number of) channels and at most 12 288 elements.  Callers holding fused
 * statistics skip rdeic_groupnorm_from_stats for such tensors: two launches cost more than the statistics save. */
int rdeic_groupnorm_is_small(int B, int64_t HW, int C1, int C2, int groups);
int rdeic_groupnorm_nhwc(const void* x1, int C1, const void* x2, int C2, int in_is_f32,
                         const float* gamma, const float* beta, void* out, int B, int64_t HW,
                         int groups, float eps, int silu, void* workspace,
                         rdeic_stream_t stream);
/* LayerNorm over the last dim of [rows, C] bf16 or fp32 -> bf16 (attention.py:273-275). */
int rdeic_layernorm(const void* x, int in_is_f32, const float* gamma, const float* beta,
                    void* out, int64_t rows, int C, float eps, rdeic_stream_t stream);

/* ---- tensor-core contractions (tcgen05 / TMEM / TMA) ------------------------------------ */

/* One implicit-GEMM descriptor covers conv3x3 (stride 1, pad 1), conv1x1 and Linear:
 *   out[m, n] = epilogue( sum_{tap, c} A[pixel(m)+tap, c] * Wp[n, tap, c] )
 * A is NHWC bf16 [a_n, a_h, a_w, a_c] (optionally the channel concat of a and a2);
 * a Linear over [M, K] is a_n = a_h = 1, a_w = M, a_c = K, taps = 1.
 * Wp is the packed weight produced by rdeic_pack_conv_weight: bf16 [n_out][taps][cpad] with
 * cpad = 64*ceil(c/64) per source, source a first then a2.
 * epilogue: v = acc + bias[n] + row_bias[sample(m)*row_bias_ld + n];  v = act(v);
 *           v = alpha*v + resid[m, n];  written as bf16 and/or fp32 with row stride ldo.
 * Batched GEMM (VAE mid attention, model.py:181-205): w_batch_stride != 0 selects a
 * different weight matrix per a_n index (elements). */
typedef struct rdeic_conv_params {
    const void* a;        int a_n, a_h, a_w, a_c;
    const void* a2;       int a2_c;
    int taps;             /* 1 (1x1 / linear), 9 (3x3, pad 1) or 25 (5x5, pad 2: model/compression.py:23,
                             compression_modules.py:80-84) */
    const void* w;        /* packed bf16 weights */
    int64_t w_batch_stride;
    int w_k, w_ld;        /* 0,0 = packed layout; else true K extent / row stride (elements) of an
                             activation used as the B operand (Q K^T, P V) */
    int n_out;
    const float* bias;
    const float* row_bias; int row_bias_ld;
    const void* resid;    int resid_is_f32; int ld_resid;
    float alpha;
    int act;              /* 0 none, 1 SiLU, 2 GEGLU: weight rows interleaved in blocks of
                             16 values | 16 gates (rdeic_b200/engine.py), writes n_out/2 columns;
                             3 LeakyReLU(act_param) (model/layers/res_blk.py:18-21,50-53,79),
                             4 exact GELU (compression_modules.py:81-83,96-99) */
    void* out_bf16;
    float* out_f32;
    int ldo;
    int tile_n_hint;      /* 0 = library picks BLOCK_N */
    void* workspace;      /* optional split-K scratch (fp32 partials); NULL disables split-K */
    int64_t workspace_bytes;
    /* ABI 2 */
    float act_param;      /* LeakyReLU negative slope */
    int a_ld, a2_ld;      /* pixel stride (elements) of a / a2 when they are channel slices of a wider
                             NHWC buffer (torch.cat inputs of compression.py:170-190); 0 = a_c / a2_c */
    /* ABI 3 */
    float* stats_out;     /* optional: [ceil(M/32)][n_out] pairs (sum, sum of squares) of the values
                             written, per 32-row slab and output column — the GroupNorm statistics of
                             the consumer (util.py:224, model.py:48) fused into the producer's epilogue.
                             Needs rdeic_conv_stats_supported(a_n, a_h, a_w); disables split-K. */
    /* ABI 5 */
    int up2;              /* 1: nearest x2 upsample + conv3x3 (openaimodel.py:106-113, model.py:63-67) with the
                             upsample folded away: output pixel (2y+py, 2x+px) only sees a 2x2 window of the INPUT,
                             so each parity class (py, px) is a 2x2 conv with pre-summed weights.  taps = 4; w holds
                             the four packed [n_out][4][cpad] matrices (parity 2 py + px) w_batch_stride elements
                             apart; a is the [a_n, a_h, a_w] input, outputs (and resid) are [a_n, 2 a_h, 2 a_w]
                             with row stride ldo; stats_out slabs are ordered [sample][parity][h][w]. */
    int a2_center;        /* 1: the second source a2 (on the OUTPUT pixel grid) contributes the centre tap only, with its
                             own packed weights w2 [n_out][cpad2]; w covers a alone ([n_out][taps][cpad1]).  Fuses
                             `h = h + zero_conv(h_ctr) * scale` (rdeic.py:194,203,207: scale folded into w2 and the
                             bias by the caller) or a 1x1 skip_connection into the conv that produces h. */
    const void* w2;
    int in_stride2;       /* 1: stride-2 conv (openaimodel.py:150-152, model.py:76-89, res_blk.py:15-25): a is the
                             [a_n, 2 a_h, 2 a_w] input, [a_n, a_h, a_w] the output grid; taps 1 or 9 */
    int pad_lo;           /* with in_stride2 and taps = 9: 1 = padding 1 all round, 0 = bottom/right only (model.py:82-84) */
} rdeic_conv_params;

/* 1 if rdeic_conv_gemm should emit stats_out for an [a_n, a_h, a_w] pixel grid, n_out columns and
 * k_blocks = taps * ceil(C/64) reduction blocks: every M tile full and row-contiguous, and the layer
 * is not one the library runs split-K (those keep split-K and a separate statistics pass); else 0. */
int rdeic_conv_stats_supported(int a_n, int a_h, int a_w, int n_out, int k_blocks);

int rdeic_conv_gemm(const rdeic_conv_params* p, rdeic_stream_t stream);
/* openaimodel.py:203 etc.: OIHW fp32 [n_out, c1+c2, kh, kw] -> packed bf16 (see above).
 * dst must hold n_out * taps * (cpad1 + cpad2) bf16. */
int rdeic_pack_conv_weight(const float* w_oihw, void* dst, int n_out, int c1, int c2, int kh,
                           int kw, rdeic_stream_t stream);

/* ---- attention ------------------------------------------------------------------------- */

/* Fused softmax(Q K^T * scale) V (attention.py:171-203).  q [B, Nq, *] with row stride ldq
 * (elements), heads laid out as consecutive d-wide column groups; k, v [B, Nk, *] likewise;
 * batch strides in elements.  d in {16, 64}, or {256, 512} with Nq and Nk multiples of 128: the VAE mid-block
 * attention of ldm/modules/diffusionmodules/model.py:181-205 (one head, d = C = 512) as a flash kernel -- the [B, N, N]
 * score tensor of the reference's bmm / softmax / bmm is never materialised.  out [B, Nq, heads*d] row stride ldo.
 * scale > 0 (the reference passes dim_head ** -0.5, attention.py:163): row maxima are taken on the raw
 * logits and the scale is folded into the multiply-add that feeds the exponential. */
int rdeic_attention(const void* q, const void* k, const void* v, void* out, int B, int heads,
                    int Nq, int Nk, int d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                    int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, float scale,
                    rdeic_stream_t stream);

/* ---- fp32 kernel mode (north_star: per-step rel-L2 <= 1e-5) ----------------------------------
 * The same UNet + control step on plain fp32 NHWC tensors and CUDA-core fp32 FMA with fp64 folding
 * of the k reduction (rdeic_b200/csrc/fp32_mode.cu).  Verification mode, not the throughput mode. */

/* conv2d / F.linear (openaimodel.py:203,229,240,106,150,566,750; attention.py:52,72,162-169,314,328):
 * a [a_n,a_h,a_w,c1] (+ a2 [..,c2], the torch.cat of openaimodel.py:804 / rdeic.py:190), ksize 1|3|5
 * (pad ksize/2; 5x5: model/compression.py:23, compression_modules.py:80-84), stride 1|2, optional nearest
 * x2 upsample of the input first (openaimodel.py:106-113);
 * w fp32 [n_out][ksize*ksize][c1+c2]; out = resid + alpha * act(conv + bias + row_bias[sample]). */
typedef struct rdeic_conv_f32_params {
    const float* a;       int a_n, a_h, a_w, c1;
    const float* a2;      int c2;
    int ksize, stride, up;
    const float* w;
    int n_out;
    const float* bias;
    const float* row_bias; int row_bias_ld;
    const float* resid;   int ld_resid;
    float alpha;
    int act;              /* 0 none, 1 SiLU, 3 LeakyReLU(act_param), 4 exact GELU (numbering of rdeic_conv_params) */
    float* out;           int ldo;
    /* ABI 5: the learned compressor's entropy-parameter nets in fp32 (model/compression.py:215-273:
     * CDF indexes must equal the reference's fp32 nets' or the arithmetic decoder desynchronises) */
    float act_param;      /* LeakyReLU negative slope */
    int a_ld, a2_ld;      /* pixel stride (elements) of a / a2 when they are channel windows of a wider NHWC
                             buffer; 0 = c1 / c2 */
    void* workspace;      /* optional split-K scratch (fp64 partials, summed in a fixed order); NULL disables it */
    int64_t workspace_bytes;
} rdeic_conv_f32_params;
int rdeic_conv_f32(const rdeic_conv_f32_params* p, rdeic_stream_t stream);
/* attention.py:171-203 in fp32: q [B,Nq,*], k/v [B,Nk,*] fp32, heads as consecutive d-wide column groups, d <= 512
 * (the VAE mid attention, model.py:181-205, is one 512-wide head). */
int rdeic_attention_f32(const float* q, const float* k, const float* v, float* out, int B, int heads, int Nq,
                        int Nk, int d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs,
                        int64_t k_bs, int64_t v_bs, int64_t o_bs, float scale, rdeic_stream_t stream);
/* attention.py:54-56 GEGLU, exact erf: in [rows, 2F] (value | gate) fp32 -> out [rows, F] fp32. */
int rdeic_geglu_f32(const float* in, float* out, int64_t rows, int F, rdeic_stream_t stream);
/* util.py:161-181 with an fp32 result. */
int rdeic_timestep_embedding_f32(const int64_t* t, float* out, int B, int dim, float max_period,
                                 rdeic_stream_t stream);
/* GroupNorm(+SiLU) / LayerNorm with fp32 input and fp32 output (same statistics kernels as the bf16 path). */
int rdeic_groupnorm_nhwc_f32(const float* x1, int C1, const float* x2, int C2, const float* gamma,
                             const float* beta, float* out, int B, int64_t HW, int groups, float eps,
                             int silu, void* workspace, rdeic_stream_t stream);
int rdeic_layernorm_f32(const float* x, const float* gamma, const float* beta, float* out, int64_t rows, int C,
                        float eps, rdeic_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RDEIC_B200_H */
