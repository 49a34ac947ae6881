/* libdove_b200 — C ABI of the B200-native DOVE one-step inference hot path.
 *
 * The reference (zhengchen1999/DOVE) has NO plugin / operator / FFI interface: its hot path
 * (/root/reference/inference_script.py:394-503, `process_video`) calls Python methods of a diffusers
 * `CogVideoXPipeline`, which dispatch to PyTorch library kernels.  Each entry point
 * below replaces one group of those library calls; the comment on each names the reference call site and the
 * diffusers module whose arithmetic it implements.  The Python host in `dove_b200/` (mirror of the pipeline
 * surface) binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions (every function):
 *   - plain pointers + sizes; all pointers are DEVICE pointers owned by the caller (no allocation, no
 *     ownership transfer); bf16 = IEEE bfloat16 stored as uint16; `stream` is a cudaStream_t passed as void*;
 *   - launches are asynchronous on `stream`, no internal synchronisation;
 *   - returns 0 on success or a negative DOVE_E_* code; never throws; `dove_last_error()` returns a
 *     thread-local message for the last failure;
 *   - thread-safe for concurrent calls on distinct streams.
 *   - activations inside the VAE are CHANNELS-LAST: [T, H, W, C] bf16 (one clip, batch 1).
 */
#ifndef DOVE_B200_H
#define DOVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOVE_OK 0
#define DOVE_E_BAD_ARG (-1)
#define DOVE_E_CUDA (-2)
#define DOVE_E_NOT_INIT (-3)
#define DOVE_E_UNSUPPORTED (-4)

#define DOVE_ABI_VERSION 1

/* GEMM / conv epilogues.  r = bf16(acc + bias[n]) in every mode (the reference rounds each op to bf16). */
#define DOVE_EPI_BIAS 0        /* out = r                                   (nn.Linear / conv)               */
#define DOVE_EPI_GELU_TANH 1   /* out = bf16(gelu_tanh(r))                  (FeedForward net.0, GELU-tanh)  */
#define DOVE_EPI_GATED_RES 2   /* out = bf16(aux + bf16(gate[seg(m)][n]*r)) (CogVideoXBlock gated residual) */
#define DOVE_EPI_ADD 3         /* out = bf16(r + aux)                       (ResnetBlock3D shortcut add)    */
#define DOVE_EPI_QKV_NORM_ROPE 4 /* internal to dove_gemm_qkv_norm_rope_bf16: per-head q/k LayerNorm + RoPE      */

int dove_abi_version(void);
/* Idempotent: selects the device's attributes (SM count, opt-in shared memory), resolves the driver's
 * cuTensorMapEncodeTiled.  Must be called once per process before any other entry point. */
int dove_init(int device);
const char* dove_last_error(void);
int dove_num_sms(void);
/* Tuning / test switches.  "conv2cta": 1 (default) = big stride-1 3x3(x3) convs on images >= 256 wide run on the
 * specialised kernels (256 output channels: CTA pair, cta_group::2 MMA + in-smem reuse of the W taps; 128 output
 * channels: swapped operands + W-tap reuse out of a 258-voxel halo row), 2 = the same without the halo-row kernel
 * (A/B runs, tests of the generic swapped-operand kernel), 0 = always the generic 1-CTA kernel (tests).
 * "attn_variant": -1 (default) = automatic; 0 = one query tile per CTA, two CTAs per SM (round-1 kernel, best below
 * ~3 000 rows); 1 + e (e = 0..5) = two query tiles per CTA sharing the K/V stages, no row-max pass in the steady state,
 * e/8 of the softmax exponentials evaluated on the FMA pipe; 7 + e (e = 0..4) = the same tiling with the S row held in
 * registers (setmaxnreg) so that the next QK^T is issued while the softmax runs, two MMA issuer warps — the default
 * for long sequences (see attn.cu). */
int dove_set_option(const char* name, int value);

/* ---- DiT -------------------------------------------------------------------------------------------------- */

/* C[M,N] = epilogue(A[M,K] * W[N,K]^T + bias).  tcgen05 + TMA persistent GEMM.
 * Replaces nn.Linear on the DiT path (to_q/k/v, to_out, ff.net.0.proj, ff.net.2, patch_embed.proj/text_proj,
 * proj_out; diffusers attention_processor.py / attention.py; ref call: inference_script.py:483-489) and the
 * VAE's 1x1x1 convs (ResnetBlock3D.conv_shortcut).
 * A: [M, lda] bf16 row-major; W: [N, ldw] bf16 row-major (PyTorch Linear weight layout); K % 64 == 0,
 * N % 16 == 0, lda/ldw/ldc % 8 == 0, 16-byte aligned pointers.  bias: bf16[N] or NULL.
 * DOVE_EPI_GATED_RES: aux [M, ld_aux] = residual stream, gate0 applies to rows < split_row (text tokens),
 * gate1 to rows >= split_row (video tokens); both bf16[N].  DOVE_EPI_ADD: aux [M, ld_aux]. C may alias aux. */
int dove_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M, int N,
                   int K, const void* bias, int epilogue, const void* aux, int64_t ld_aux, const void* gate0,
                   const void* gate1, int split_row, void* stream);

/* y[n] = bf16(sum_k act(x[k]) * W[n,k] + b[n]); act = SiLU if silu_in.  One row (the timestep embedding path:
 * TimestepEmbedding, CogVideoXLayerNormZero.linear, AdaLayerNorm.linear — constant for t = 399, run at load). */
int dove_gemv_bf16(const void* x, const void* W, const void* b, void* y, int N, int K, int silu_in, void* stream);

/* out[r,:] = bf16(bf16(LN(x[r,:]; w, b, eps)) * (1 + scale_s) + shift_s), s = 0 for r < split_row else 1.
 * CogVideoXLayerNormZero / AdaLayerNorm apply (diffusers normalization.py).  scale/shift: bf16[D] each, or NULL
 * for a plain affine LayerNorm (norm_final).  D % 256 == 0, D <= 4096. */
int dove_layernorm_mod_bf16(const void* x, void* out, int rows, int D, const void* ln_w, const void* ln_b,
                            float eps, const void* scale0, const void* shift0, const void* scale1,
                            const void* shift1, int split_row, void* stream);

/* In place on qkv [rows, 3*heads*64]: per-head LayerNorm(64, eps) with affine on q and k, then 3D RoPE
 * (interleaved pairs, fp32) on rows >= text_len using cos/sin [rows - text_len, 64] fp32.
 * CogVideoXAttnProcessor2_0 (norm_q/norm_k + apply_rotary_emb). */
int dove_qk_norm_rope_bf16(void* qkv, int rows, int heads, const void* q_w, const void* q_b, const void* k_w,
                           const void* k_b, float eps, const float* cos, const float* sin, int text_len,
                           void* stream);

/* qkv[M, 3*heads*64] = [to_q | to_k | to_v](A) with the per-head LayerNorm(64, eps) on q and k and the 3-D RoPE
 * (rows >= text_len) applied in the GEMM epilogue: dove_gemm_bf16 + dove_qk_norm_rope_bf16 in ONE kernel (one head =
 * 64 accumulator columns held by one thread).  W: [3*heads*64, ldw] = cat(to_q, to_k, to_v weights), bias likewise;
 * heads*64 % 256 == 0.  cos_t / sin_t: the RoPE tables TRANSPOSED and pair-deduplicated, fp32 [32][M - text_len] with
 * cos_t[i][token] = cos[token][2i] (= cos[token][2i+1]: get_3d_rotary_pos_embed repeats every frequency twice), so a
 * warp's 32 rows read contiguous memory.  CogVideoXAttnProcessor2_0 up to (excluding) F.scaled_dot_product_attention. */
int dove_gemm_qkv_norm_rope_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int M,
                                 int heads, int K, const void* bias, const void* q_w, const void* q_b,
                                 const void* k_w, const void* k_b, float eps, const float* cos_t, const float* sin_t,
                                 int text_len, void* stream);

/* out[rows, heads*64] = softmax(q k^T * scale) v per head, non-causal, no mask.  q,k,v read from the fused
 * qkv [rows, 3*heads*64] buffer.  tcgen05 flash-attention kernel (S, P and O in TMEM, lazy rescaling).
 * F.scaled_dot_product_attention in CogVideoXAttnProcessor2_0. */
int dove_attention_bf16(const void* qkv, void* out, int rows, int heads, float scale, void* stream);

/* tokens[T/2*h/2*w/2, 128] <- latent [F, 16, h, w] (patchify gather, feature order (C,pt,ph,pw));
 * CogVideoXPatchEmbed reshape/permute. */
int dove_patchify_bf16(const void* latent, void* tokens, int F, int C, int h, int w, void* stream);

/* x0[F,16,h,w] = bf16(bf16(a*latent) - bf16(b*unpatchify(tokens))) — un-patchify of proj_out fused with
 * scheduler.get_velocity (ref: inference_script.py:491-493); a, b are the dtype-rounded sqrt(alpha_bar),
 * sqrt(1-alpha_bar).  pred_out (optional, may be NULL) receives the un-patchified model output. */
int dove_unpatchify_velocity_bf16(const void* tokens, const void* latent, void* x0, void* pred_out, int F, int C,
                                  int h, int w, float a, float b, void* stream);

/* out = bf16(bf16(a*noise) - bf16(b*sample)) — CogVideoXDPMScheduler.get_velocity(sample, noise, t) standalone
 * (the call-surface path; the fused path uses dove_unpatchify_velocity_bf16). */
int dove_velocity_bf16(const void* sample, const void* noise, void* out, int64_t n, float a, float b, void* stream);

/* ---- VAE (channels-last) ---------------------------------------------------------------------------------- */

/* Implicit-GEMM convolution on channels-last bf16, tcgen05 + TMA (no im2col buffer).
 *   x: [Tin, Hin, Win, Cin], Cin % 64 == 0.  For causal 3x3x3 convs the caller provides the temporally padded
 *      input (Tin = Tout + 2: two cached / replicated frames first); dove_conv3d_causal_bf16 below avoids that copy.
 *   w: [Cout_pad, kt*kh*kw*Cin] bf16, K index = ((dt*kh + dh)*kw + dw)*Cin + c; Cout_pad % 16 == 0.
 *   y: out_mode 0: [Tout, Ho, Wo, ldy] channels-last (ldy >= cout_valid); out_mode 1: planar, element (n, voxel) at y[n*ldy + voxel]
 *      (ldy = plane stride >= Tout*Ho*Wo: lets a frame batch land inside a longer NCDHW clip).
 *      out_mode 2: planar like 1 with the reference's post-processing fused (ref: inference_script.py:501):
 *      bf16 clamp(bf16(bf16(r*0.5)+0.5), 0, 1).  out_mode 3: the same value quantised as the reference's savers do
 *      (ref: inference_script.py:124, 143, 168: `(video*255).clamp(0,255).to(uint8)` on the fp32 copy): y is uint8,
 *      y[n*ldy + voxel] = trunc(fp32(v) * 255) — the multi-GPU gather and the D2H then move 1 byte per element.
 *   stride: spatial stride (1 or 2); pad: low-side spatial zero pad (1 for "same" 3x3, 0 for the
 *   CogVideoXDownsample3D conv whose (0,1,0,1) pad is right/bottom only — high-side pad comes from TMA OOB fill).
 *   epilogue: DOVE_EPI_BIAS or DOVE_EPI_ADD (aux: [Tout,Ho,Wo,ld_aux]).
 *   gn_partial / gn_done (both may be NULL): fused GroupNorm statistics of the OUTPUT tensor.  When the kernel chosen
 *   for this shape supports it (the swapped-operand and CTA-pair kernels, i.e. all large convs), every CTA writes
 *   per-group (sum, sum of squares) of the bf16-rounded outputs to gn_partial (dove_gn_partial_floats() floats,
 *   zeroed by this call) and *gn_done (HOST int) is set to 1; the next GroupNorm then only needs dove_gn_finalize —
 *   the separate statistics pass over the tensor disappears.  *gn_done = 0: run dove_gn_stats_bf16 as usual.
 * Replaces F.conv3d / F.conv2d in CogVideoXCausalConv3d, CogVideoXDownsample3D, CogVideoXUpsample3D. */
int dove_conv_cl_bf16(const void* x, const void* w, const void* bias, void* y, int Tout, int Hin, int Win, int Cin,
                      int Cout_pad, int cout_valid, int64_t ldy, int kt, int kh, int kw, int stride, int pad,
                      int Ho, int Wo, int epilogue, const void* aux, int64_t ld_aux, int out_mode, float* gn_partial,
                      int* gn_done, void* stream);

/* Causal 3x3x3 convolution (CogVideoXCausalConv3d, stride 1, spatial zero pad 1) on the UN-padded frame batch
 * x [T,H,W,Cin] with ZERO-COPY temporal padding: the two frames preceding x are read from x_prev [2,H,W,Cin] (the
 * conv cache = the last two frames of the previous frame batch's input, typically a view of that buffer) through a
 * second tensor map, or, when x_prev is NULL (first frame batch), input frame 0 is replicated — exactly
 * `fake_context_parallel_forward` + `conv_cache` of diffusers, without materialising the padded tensor.
 * Other arguments as dove_conv_cl_bf16. */
int dove_conv3d_causal_bf16(const void* x, const void* x_prev, const void* w, const void* bias, void* y, int T, int H,
                            int W, int Cin, int Cout_pad, int cout_valid, int64_t ldy, int epilogue, const void* aux,
                            int64_t ld_aux, int out_mode, float* gn_partial, int* gn_done, void* stream);

/* GroupNorm statistics over a channels-last tensor [nvox, C]: mean/rstd per group -> stats[2*groups] fp32.
 * partial: workspace of dove_gn_partial_floats(nvox, groups) floats. */
size_t dove_gn_partial_floats(int64_t nvox, int groups);
int dove_gn_stats_bf16(const void* x, int64_t nvox, int C, int groups, float eps, float* partial, float* stats,
                       void* stream);

/* mean / rstd per group from per-CTA partials written by a conv epilogue (see dove_conv_cl_bf16 gn_partial). */
int dove_gn_finalize(const float* partial, int64_t nvox, int C, int groups, float eps, float* stats, void* stream);

/* y = [silu]( bf16(bf16(GN(x))*cy + cb) ) or [silu](bf16(GN(x))) written to out (may be the frame-offset view of a
 * temporally padded conv input).  Spatial-norm variant (zq_y != NULL): cy/cb = conv_y/conv_b(zq) evaluated at
 * latent resolution [Tz, hz, wz, C] and gathered with the nearest-neighbour index of
 * CogVideoXSpatialNorm3D (first frame mapped separately when T is odd and > 1).
 * GroupNorm+SiLU of CogVideoXResnetBlock3D / Encoder3D.norm_out / Decoder3D.norm_out. */
int dove_gn_apply_bf16(const void* x, void* out, int T, int H, int W, int C, int groups, const float* stats,
                       const void* gamma, const void* beta, int apply_silu, const void* zq_y, const void* zq_b,
                       int Tz, int hz, int wz, void* stream);

/* Temporal average pooling of CogVideoXDownsample3D(compress_time): T odd -> keep frame 0, avg pairs of the
 * rest; T even -> avg pairs.  x [T, HWC] -> y [Tout, HWC]. */
int dove_time_pool_bf16(const void* x, void* y, int T, int64_t frame_elems, void* stream);

/* Nearest-neighbour upsample x2 in (H, W) and optionally T, CogVideoXUpsample3D semantics (T odd > 1: frame 0
 * is not duplicated in time).  x [T,H,W,C] -> y [Tout,2H,2W,C]. */
int dove_upsample_nearest_bf16(const void* x, void* y, int T, int H, int W, int C, int time_x2, void* stream);

/* pixels NCDHW (fp32 or bf16) [3, T, H, W] -> channels-last bf16 [T(+t_off), H, W, Cpad] zero-padded channels. */
int dove_pixels_to_cl_bf16(const void* x, int x_is_fp32, void* y, int T, int H, int W, int Cpad, void* stream);

/* Layout glue for the small latent tensors: NCDHW [C,T,H,W] <-> channels-last [T,H,W,Cpad] (bf16). */
int dove_ncthw_to_cl_bf16(const void* x, void* y, int C, int T, int H, int W, int Cpad, float scale, void* stream);
int dove_cl_to_ncthw_bf16(const void* x, void* y, int C, int T, int H, int W, int ldx, void* stream);

/* z = bf16(bf16(mean + bf16(std * noise)) * scaling), std = exp(0.5*clamp(logvar,-30,20));
 * moments channels-last [nvox, 32] (mean = ch 0..15, logvar = ch 16..31); noise, z: [16, nvox] (NCDHW).
 * DiagonalGaussianDistribution.sample() * scaling_factor (ref: inference_script.py:409). */
int dove_gaussian_sample_bf16(const void* moments, const void* noise, void* z, int64_t nvox, float scaling,
                              void* stream);

/* Host -> device copy of a strided box without a staging pass: `planes` planes of `rows` rows of `row_bytes` bytes, read
 * from (pinned) host memory with row pitch `src_row_pitch` bytes and `src_rows_per_plane` rows between plane starts,
 * written densely to dst.  One cudaMemcpy3DAsync (a DMA descriptor, no CPU touch of the payload).  Replaces the
 * per-unit `video_chunk.to(device)` of a chunk x tile view of the clip (ref: inference_script.py:407, :690-700). */
int dove_h2d_box_async(const void* src, int64_t src_row_pitch, int64_t src_rows_per_plane, void* dst, int64_t row_bytes,
                       int64_t rows, int64_t planes, void* stream);

/* Pre-processing of the low-quality clip on the GPU (ref: inference_script.py:672-679): lr [F,3,h,w] fp32 in 0..255
 * -> out [3, F, scale*h, scale*w] fp32 = bilinear (align_corners=False) upscale, then x/255*2-1. */
int dove_upscale_normalize_f32(const float* lr, float* out, int F, int h, int w, int scale, void* stream);

/* Linear-ramp blend of two overlapping VAE tiles, in place on b (AutoencoderKLCogVideoX.blend_v / blend_h used by
 * tiled_encode / tiled_decode, i.e. `--is_vae_st`):  for p < extent:
 *   b[o,p,q,c] = bf16(bf16(a[o, a_len-extent+p, q, c]*(1-p/extent)) + bf16(b[o,p,q,c]*(p/extent)))
 * with element strides (so, sp, sq) of the outer / blended / other axes and a contiguous inner axis. */
int dove_blend_bf16(const void* a, void* b, int outer, int extent, int other, int inner, int64_t a_so, int64_t a_sp,
                    int64_t a_sq, int a_len, int64_t b_so, int64_t b_sp, int64_t b_sq, void* stream);

/* y = clamp(x*0.5+0.5, 0, 1) elementwise bf16 (ref: inference_script.py:501). */
int dove_post_scale_bf16(const void* x, void* y, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DOVE_B200_H */
