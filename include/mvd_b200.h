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
 * out[n, y, x, :] = sum_taps [x | x2](n, s*y+ky-1, s*x+kx-1, :) @ w[:, ky, kx, :]^T + bias + img_bias[n, :]
 *                   (+ residual[n, y, x, :]).
 * Replaces diffusers ResnetBlock2D.conv1/conv2, Downsample2D.conv (stride 2), Upsample2D.conv
 * (invoked from src/models/mvd_unet.py:318 and src/models/image_encoder.py:105).
 * x2 (optional) is a second channel block concatenated after x (skip connections of the up path).
 * cin1, cin2 multiples of 64; c_out multiple of 32; h_out/w_out are OUTPUT sizes. */
int mvd_conv3x3_bf16(const void* x, int cin1, const void* x2, int cin2, const void* w, const void* bias,
                     const float* img_bias, const void* residual, void* out, int n_img, int h_out, int w_out,
                     int c_out, int stride, int tile_n, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* MVD_B200_H_ */
