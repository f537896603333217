/* mvd_b200.h — C ABI of libmvd_b200.so: the sm_100a kernels behind MVD's multi-view denoising hot path.
 *
 * The reference (pananananas/MVD) has no FFI: every op below replaces a stock PyTorch/diffusers call made
 * from the reference's Python hot path. Each entry point cites the reference call site it stands in for
 * (paths relative to the reference repo) and, where the arithmetic lives in the un-vendored
 * diffusers==0.32.2, the diffusers module it restates.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in _host. Memory is owned by the caller.
 *  - Activations are bf16, channels-last: [N, H, W, C] == [N, H*W, C] row-major ("NLC"). Weights are bf16
 *    [out_features, in_features] row-major (PyTorch Linear layout); 3x3 conv weights are [Cout, 3, 3, Cin].
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous and capture-safe (no sync,
 *    no allocation, no host read).
 *  - Return value: MVD_OK (0) or a negative MVD_ERR_*; mvd_last_error() gives the text (thread-local).
 *    Unsupported shapes are hard errors: there is no CPU fallback and no alternate backend.
 */
#ifndef MVD_B200_H_
#define MVD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVD_ABI_VERSION 1
#define MVD_OK 0
#define MVD_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define MVD_ERR_CUDA (-2)    /* CUDA runtime / driver error      */

const char* mvd_last_error(void);
int mvd_abi_version(void);
/* Number of CUDA kernels this library has launched (or recorded into a stream capture) in this process. */
int64_t mvd_kernel_launch_count(void);
/* Process-wide switch for programmatic dependent launch of every kernel of this library (the next kernel's prologue —
 * barrier init, TMEM allocation, descriptor prefetch — overlaps the previous kernel's tail; data dependences are kept by
 * griddepcontrol.wait). Pays for launch-bound steps (1-2 samples per GPU), costs ~3 % for throughput-bound ones; the
 * environment variable MVD_PDL=0/1 overrides it. No reference analogue (scheduling detail). */
int mvd_set_launch_overlap(int on);

/* ---------------------------------------------------------------------------------------------------------
 * Tensor-core contractions (tcgen05 + TMEM + TMA), csrc/gemm.cu
 * ------------------------------------------------------------------------------------------------------- */

/* out[M, N] = [a | a2][M, k1+k2] @ w[N, k1+k2]^T (+ bias[N]) (+ row_group_bias[row / rows_per_group, N])
 *             (+ residual[M, N]);   geglu != 0: w rows are interleaved [a-block | g-block] per tile_n columns
 *             and out[M, N/2] = a * gelu_erf(g).
 * Replaces nn.Linear at src/models/attention.py:125,129,131,157 (to_q_ref/to_k_ref/to_v_ref/to_out_ref) and
 * diffusers Attention.to_q/to_k/to_v/to_out[0], Transformer2DModel.proj_in/proj_out, FeedForward (GEGLU,
 * Linear) called from src/models/mvd_unet.py:318. k1, k2 multiples of 64; N multiple of 32; ld* in elements. */
int mvd_linear_bf16(const void* a, int64_t lda, int k1, const void* a2, int64_t lda2, int k2, const void* w,
                    int64_t ldw, const void* bias, const float* row_group_bias, int row_group_bias_ld,
                    int rows_per_group, const void* residual, int64_t ldr, void* out, int64_t ldo, int M, int N,
                    int geglu, int tile_n, void* stream);

/* 3x3 convolution, padding 1, stride 1 or 2, NHWC bf16, implicit GEMM.
 * out[n, y, x, :] = sum_taps [x | x2](n, s*y+ky-1, s*x+kx-1, :) @ w[:, ky, kx, :]^T + bias + img_bias[n, :] (fp32, row stride
 *                   img_bias_ld floats, a multiple of 4; 0 = ONE row shared by every image, e.g. the time-embedding
 *                   projection of a scalar timestep)
 *                   (+ residual[n, y, x, :]).
 * Replaces diffusers ResnetBlock2D.conv1/conv2, Downsample2D.conv (stride 2), Upsample2D.conv
 * (invoked from src/models/mvd_unet.py:318 and src/models/image_encoder.py:105).
 * x2 (optional) is a second channel block concatenated after x (skip connections of the up path).
 * cin1, cin2 multiples of 64; c_out multiple of 32; h_out/w_out are OUTPUT sizes. */
int mvd_conv3x3_bf16(const void* x, int cin1, const void* x2, int cin2, const void* w, const void* bias,
                     const float* img_bias, int img_bias_ld, const void* residual, void* out, int n_img, int h_out,
                     int w_out, int c_out, int stride, int tile_n, void* stream);

/* Optional fused neighbours of a GEMM / conv launch (mvd_linear_ex_bf16, mvd_conv3x3_ex_bf16); every pointer may be NULL.
 *  - LayerNorm of the A rows folded into the GEMM (diffusers BasicTransformerBlock.norm1/2/3 in front of attn1 / attn2 /
 *    ff): the caller passes weights already scaled by gamma per input channel, `ln_colsum[n]` =
 *    sum_k of the scaled bf16 weight row n (fp32), folds W.beta into the bias, and `ln_stats` = [M][ln_parts][2] fp32
 *    partial (sum x, sum x^2) of each A row as written by the producing launch's `stats_out`. The epilogue computes
 *    rstd * (acc - mean * ln_colsum[n]) — identical to LayerNorm followed by the GEMM, one kernel and one pass less.
 *  - stats_out: [M][stats_parts][2] fp32 partial row statistics of the bf16 OUTPUT, one pair per column tile
 *    (stats_parts must equal the launch's column-tile count: mvd_gemm_plan's ceil(N / bn)); plain linears only.
 *  - FiLM on the output (src/models/camera_encoder.py:221-234 applied by the forward hooks of src/models/mvd_unet.py:
 *    354-385 to a block's output): out = v * film_scale[g][n] + film_shift[g][n], g = image (conv) or row group
 *    (linear, rows_per_group), fp32 rows of stride film_ld.
 *  - workspace: with it, the part of a launch that does not fill the machine — every tile of a launch with fewer tiles
 *    than SMs (the 16x16 / 8x8 UNet levels, a view-sharded rank), or the last partial wave of a big one — shares its
 *    k-blocks evenly over all SMs (stream-K); the fp32 partial tiles are reduced in k order by the last-arriving CTA
 *    (deterministic), and for launches under two waves the tile width is re-chosen for the shared schedule. Launches
 *    with ln_stats / stats_out or GEGLU ignore it. 64 MiB cover every case of the UNet; too small a buffer simply
 *    means whole tiles. mvd_gemm_plan reports the plan WITHOUT a workspace (the one the stats_out contract refers to). */
typedef struct mvd_gemm_extras {
  const float* ln_stats;
  const float* ln_colsum;
  int ln_parts;
  float ln_eps;
  float* stats_out;
  int stats_parts;
  const float* film_scale;
  const float* film_shift;
  int film_ld;
  void* workspace;          /* optional scratch (16-byte aligned, zero-filled once, private to launches that cannot run */
  int64_t workspace_bytes;  /* concurrently): stream-K partial tiles + tickets                                       */
} mvd_gemm_extras;

int mvd_linear_ex_bf16(const void* a, int64_t lda, int k1, const void* a2, int64_t lda2, int k2, const void* w,
                       int64_t ldw, const void* bias, const float* row_group_bias, int row_group_bias_ld,
                       int rows_per_group, const void* residual, int64_t ldr, void* out, int64_t ldo, int M, int N,
                       int geglu, int tile_n, const mvd_gemm_extras* extras, void* stream);
int mvd_conv3x3_ex_bf16(const void* x, int cin1, const void* x2, int cin2, const void* w, const void* bias,
                        const float* img_bias, int img_bias_ld, const void* residual, void* out, int n_img, int h_out,
                        int w_out, int c_out, int stride, int tile_n, const mvd_gemm_extras* extras, void* stream);

/* The scheduling decision mvd_linear_bf16 (n_img = h_out = 1, w_out = M, ntaps = 1) / mvd_conv3x3_bf16 (ntaps = 9)
 * would take for a problem, without launching anything: tile width, whether the weight tile stays resident in shared
 * memory (weight-stationary tiles of the K <= 320 linears) and the grid size. Any output pointer may be NULL. For
 * tests and tuning; no reference analogue. */
int mvd_gemm_plan(int n_img, int h_out, int w_out, int c_in, int c_out, int ntaps, int stride, int geglu, int tile_n,
                  int* bn, int* weight_stationary, int* grid);

/* The stream-K part of that decision for a plain launch that is handed `workspace_bytes` of scratch: tile width, tile
 * and k-block counts, and the schedule — tiles [0, sk_first) are whole-tile work items, the k-blocks of tiles
 * [sk_first, tiles) are shared by CTAs [0, sk_ctas) (CTA i owns units [floor(i T / G), floor((i + 1) T / G)) of the
 * T = (tiles - sk_first) * k_blocks units), with sk_slots partial-tile slots per shared tile. sk_ctas = 0: no stream-K.
 * For tests of the schedule (coverage, slot bound) without a GPU; no reference analogue. */
int mvd_gemm_plan_streamk(int n_img, int h_out, int w_out, int c_in, int c_out, int ntaps, int stride, int tile_n,
                          int64_t workspace_bytes, int* bn, int* tiles, int* k_blocks, int* sk_first, int* sk_ctas,
                          int* sk_slots);

/* ---------------------------------------------------------------------------------------------------------
 * Fused flash-attention forward, head_dim 64 (tcgen05 + TMEM + TMA), csrc/attn.cu
 * ------------------------------------------------------------------------------------------------------- */

/* out[b, i, h*64:(h+1)*64] = softmax_j(scale * q[b,i,h] . k[b,j,h]) @ v[b,j,h]   (no mask, no dropout)
 * Replaces F.scaled_dot_product_attention at src/models/attention.py:148-150 (reference-image / cross-view
 * branch, S_kv arbitrary) and the SDPA inside diffusers AttnProcessor2_0 = `original_processor`
 * (src/models/attention.py:62-70). q/k/v/out are [batch, S, ld*] bf16 with head h at columns h*64..h*64+63
 * (i.e. the un-transposed output of the projection GEMMs); ld* and *_batch_stride in elements. */
int mvd_attention_bf16(const void* q, int64_t ldq, int64_t q_batch_stride, const void* k, int64_t ldk,
                       int64_t k_batch_stride, const void* v, int64_t ldv, int64_t v_batch_stride, void* out,
                       int64_t ldo, int64_t o_batch_stride, int batch, int heads, int s_q, int s_kv, float scale,
                       void* stream);

/* Same operator with a caller-owned scratch buffer of mvd_attention_workspace_bytes() bytes (16-byte aligned,
 * zero-filled once, private to launches that cannot run concurrently, e.g. one per stream). With it, the units of a
 * launch's last partial wave are split along S_kv over the idle SMs and merged by whichever CTA finishes a unit last
 * (deterministic part order); without it (NULL) every (256-row tile pair, head, batch) unit runs on one CTA.
 * co_units: number of such units of OTHER attention launches the caller has in flight on another stream (0 if none):
 * the last wave is then shared with them and only this launch's share of it is split.
 * k_batch_stride = v_batch_stride = 0 shares one K/V sequence between all batch entries: the cross-view reference
 * mode of configs[3], where every view attends over the concatenated tokens of all views
 * (src/models/attention.py:190-197 accepts a 3-D reference; :126-132 projects it once). */
int64_t mvd_attention_workspace_bytes(void);
int mvd_attention_bf16_ws(const void* q, int64_t ldq, int64_t q_batch_stride, const void* k, int64_t ldk,
                          int64_t k_batch_stride, const void* v, int64_t ldv, int64_t v_batch_stride, void* out,
                          int64_t ldo, int64_t o_batch_stride, int batch, int heads, int s_q, int s_kv, float scale,
                          void* workspace, int64_t workspace_bytes, int co_units, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Bandwidth-bound normalisation kernels, csrc/norm.cu
 * ------------------------------------------------------------------------------------------------------- */

/* GroupNorm(groups, eps)(+SiLU) over NHWC bf16 [n_img, hw, c1(+c2)]; x2 (optional) is channel-concatenated after
 * x1 (fuses torch.cat of the up-path skip connection). out is dense [n_img, hw, c1+c2].
 * Replaces diffusers ResnetBlock2D.norm1/norm2 + nonlinearity, Transformer2DModel.norm (eps 1e-6, no SiLU) and
 * UNet2DConditionModel.conv_norm_out + conv_act. workspace: 4-byte-element scratch of
 * mvd_groupnorm_workspace_floats() elements that MUST be zero-filled once before its first use (it starts with
 * per-image ticket counters which the kernel re-arms itself; results are deterministic). */
int64_t mvd_groupnorm_workspace_floats(int n_img, int hw, int groups);
int mvd_groupnorm_bf16(const void* x1, int c1, const void* x2, int c2, const void* gamma, const void* beta, void* out,
                       int n_img, int hw, int groups, float eps, int silu, float* workspace, int64_t workspace_floats,
                       void* stream);

/* LayerNorm over the last dim of bf16 [M, C] (BasicTransformerBlock.norm1/2/3) / of small fp32 rows with
 * optional SiLU (CameraEncoder MLPs, src/models/camera_encoder.py:31-76,81-85). gamma/beta bf16. */
int mvd_layernorm_bf16(const void* x, int64_t ldx, const void* gamma, const void* beta, void* out, int64_t ldo, int M,
                       int C, float eps, void* stream);
int mvd_layernorm_f32(const float* x, const void* gamma, const void* beta, float* out, int M, int C, float eps,
                      int silu, void* stream);

/* Reference-feature normalisation, src/models/attention.py:95-103:
 *   out = (x - mean) / clamp(std_unbiased, 1e-6) * 0.5, statistics over dims (0,1) of the RAW reference tensor.
 * x/out: bf16 [batch, seq, channels] (channels-last). per_pixel=1: the reference tensor was 4-D [B,C,H,W] ->
 * statistics per pixel over (batch, channel); per_pixel=0: it was 3-D [B,S,C] -> per channel over (batch, seq). */
int64_t mvd_refnorm_workspace_floats(int channels);
int mvd_refnorm_bf16(const void* x, void* out, int batch, int seq, int channels, int per_pixel, float* workspace,
                     int64_t workspace_floats, void* stream);
/* per_pixel = 0 form for a reference that is `replication` IDENTICAL copies of x along the batch (cross-view mode: every
 * sample gets the tokens of all views, src/models/attention.py:190-197 leaves a 3-D reference as is): the statistics
 * of the replicated tensor (mean unchanged, unbiased std with rep*rows - 1 in the denominator) without materialising
 * it; out holds ONE normalised copy, shared by all samples through the attention op's zero batch stride. */
int mvd_refnorm_replicated_bf16(const void* x, void* out, int batch, int seq, int channels, int per_pixel,
                                int replication, float* workspace, int64_t workspace_floats, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Elementwise / embedding / layout kernels, csrc/elementwise.cu
 * ------------------------------------------------------------------------------------------------------- */

/* Camera FiLM, src/models/camera_encoder.py:221-234 (hooked at src/models/mvd_unet.py:354-385):
 *   out = x * (2*sigmoid(mod[:, :C])*strength) + mod[:, C:]*strength; mod fp32 [n_cam, 2C]; image n uses
 *   camera n % n_cam. x/out NHWC bf16 [n_img, hw, C] (in place allowed). */
int mvd_film_bf16(const void* x, void* out, const float* mod, int n_img, int n_cam, int hw, int channels,
                  float strength, void* stream);

/* out[M,N] = act_out(act_in(x[M,K]) @ w[N,K]^T + bias), M <= 16, fp32 activations, bf16 weights.
 * TimestepEmbedding, ResnetBlock2D.time_emb_proj (all blocks in one call), CameraEncoder MLPs and modulators
 * (src/models/camera_encoder.py:31-76,81-88,185-194,221). */
int mvd_small_linear_f32(const float* x, int64_t ldx, const void* w, const void* bias, float* out, int64_t ldo, int M,
                         int N, int K, int silu_in, int silu_out, void* stream);

/* diffusers Timesteps(dim, flip_sin_to_cos=True, freq_shift=0): out[b] = [cos(t f) | sin(t f)], fp32. */
int mvd_timestep_embedding_f32(const float* timesteps, int n_timesteps, float* out, int batch, int dim, void* stream);

/* CameraEncoder.compute_relative_transform + the sinusoidal part of positional_encoding
 * (src/models/camera_encoder.py:107-120,137-151). cams fp32 [V,3,4]; r_flat [V,9]; t_enc [V, 6*pos_enc_dim];
 * t_rel [V,3] (optional, may be NULL). */
int mvd_camera_front_f32(const float* source_cam, const float* target_cam, float* r_flat, float* t_enc, float* t_rel,
                         int n_views, int pos_enc_dim, float max_freq, void* stream);

/* UNet conv_in (Conv2d(4,Cout,3,p=1)) on fp32 NCHW latents -> NHWC bf16, with the input-latent FiLM of
 * src/models/mvd_unet.py:256-258 (mod fp32 [n_cam, 8] or NULL) and the CFG duplication of
 * src/models/pipeline.py:141 (image n reads latent n % n_latents) folded in. w: [Cout,3,3,4]. */
int mvd_conv_in_f32_bf16(const float* latents, int n_latents, const float* mod, int n_cam, float strength,
                         const void* w, const void* bias, void* out, int n_img, int h, int wdt, int c_out,
                         void* stream);

/* UNet conv_out (Conv2d(Cin,4,3,p=1)) on NHWC bf16 -> fp32 NCHW [n_img,4,h,w]. w: [4,3,3,Cin]. */
int mvd_conv_out_bf16_f32(const void* x, const void* w, const void* bias, float* out, int n_img, int h, int wdt,
                          int c_in, void* stream);

/* The first 4 channels of an NHWC bf16 tensor (pixel stride ld elements) as fp32 NCHW [n_img,4,hw]: the tail of the
 * UNet's conv_out (diffusers unet_2d_condition.py conv_out; reference call site src/models/mvd_unet.py:318) when that
 * conv runs through mvd_conv3x3_bf16 with its 4 weight rows zero-padded to 32. */
int mvd_head4_to_nchw_f32(const void* x, int64_t ld, float* out, int n_img, int64_t hw, void* stream);

/* F.interpolate(scale_factor=2, mode="nearest") of diffusers Upsample2D, NHWC bf16. */
int mvd_upsample_nearest2x_bf16(const void* x, void* out, int n_img, int h, int wdt, int channels, void* stream);

int mvd_add_bf16(const void* a, const void* b, void* out, int64_t n, void* stream);
int mvd_cast_f32_bf16(const float* x, void* out, int64_t n, void* stream);
/* x: [batch, rows, cols] -> out: [batch, cols, rows]; dtype codes 0 = fp32, 1 = bf16 (NCHW <-> NHWC). */
int mvd_transpose_batched(const void* x, void* out, int batch, int rows, int cols, int src_dtype, int dst_dtype,
                          void* stream);

/* CFG combine + DDPM v-prediction step, src/models/pipeline.py:156-158,161 (diffusers DDPMScheduler.step):
 *   v = v_u + g (v_c - v_u); x0 = sqrt_abar*x - sqrt_1m_abar*v; x <- coef_x0*x0 + coef_xt*x + sigma*noise.
 * model_out fp32 [cfg*n]; latents fp32 [n] updated in place; noise fp32 [n] or NULL. */
int mvd_cfg_ddpm_step_f32(const float* model_out, float* latents, const float* noise, int64_t n, int cfg,
                          float guidance, float sqrt_alpha_bar, float sqrt_one_minus_alpha_bar, float coef_x0,
                          float coef_xt, float sigma, void* stream);

/* Device-table variants for replaying the whole sampling loop as one CUDA graph: the per-step
 * scalars live in coef_table[steps][8] = {t, sqrt_abar, sqrt_1m_abar, c_x0, c_xt, sigma, 0, 0} and are selected by
 * the device-side counter *step_idx; noise_table is fp32 [steps][n] or NULL. mvd_advance_step increments the
 * counter (wrapping at n_steps) and writes the next timestep for mvd_timestep_embedding_f32. */
int mvd_cfg_ddpm_step_table_f32(const float* model_out, float* latents, const float* noise_table, int64_t n, int cfg,
                                float guidance, const float* coef_table, const int* step_idx, void* stream);
int mvd_advance_step(int* step_idx, const float* coef_table, float* timestep_out, int n_steps, void* stream);
/* mvd_advance_step that additionally copies row *step_idx (after the increment) of row_table [n_steps][row_len] fp32
 * into row_out: the per-schedule table of everything that depends on the timestep only — diffusers Timesteps +
 * TimestepEmbedding + the 22 ResnetBlock2D.time_emb_proj(SiLU(temb)) outputs, computed once
 * per session instead of once per step. row_len multiple of 4, rows 16-byte aligned. */
int mvd_advance_step_rows(int* step_idx, const float* coef_table, float* timestep_out, int n_steps,
                          const float* row_table, float* row_out, int row_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVD_B200_H_ */
