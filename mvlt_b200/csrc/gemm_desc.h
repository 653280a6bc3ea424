// Plain-C description of one (strided-batched) GEMM launch. Shared by include/mvlt_b200.h (C-ABI),
// the tcgen05 kernel and the SIMT cross-check kernel.
//
//   D[b1,b2][m,n] = epilogue( alpha * sum_k A[b1,b2][m,k] * B[b1,b2][n,k] )
//
// A is M x K, B is N x K (i.e. D = A * B^T in BLAS terms). Each operand is bf16 and either
//   K-major  (x_mn = 0): element (r,k) at  base + r*ld + k      (the reference's nn.Linear weight layout)
//   MN-major (x_mn = 1): element (r,k) at  base + k*ld + r      (a transposed view, no copy)
// so the three GEMMs of a Linear layer (y = x W^T, dx = dy W, dW = dy^T x) and the four batched
// products of attention all map onto this one descriptor without materialising a transpose.
//
// Implicit 3x3 convolution (stride 1, zero padding 1) over an NHWC bf16 tensor X[b, y, x, c] (conv_* fields): the
// im2col matrix col[(b,y,x), tap*C + c] = X[b, y + tap/3 - 1, x + tap%3 - 1, c] is never materialised; TMA fetches
// shifted boxes of X (out-of-bounds = the zero padding) straight into the swizzled operand tiles.
//   conv_mode = MVLT_CONV_A : A := col    (M = B*H*W, K = 9*C, a_mn = 0)   forward / input-gradient convolutions
//   conv_mode = MVLT_CONV_BT: B := col^T  (N = 9*C, K = B*H*W, b_mn = 1)   weight-gradient GEMM dW = dY^T col
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MVLT_ACT_NONE = 0,
  MVLT_ACT_GELU = 1,   // D = gelu_erf(v); optional D2 = v (pre-activation, bf16) for the backward pass
  MVLT_ACT_DGELU = 2,  // D = v * gelu_erf'(aux[m,n])   (aux = saved pre-activation, bf16)
  MVLT_ACT_GELU_SAVE_GRAD = 3,  // D = gelu_erf(v); D2 = gelu_erf'(v) (bf16): the backward pass is then a plain multiply
  MVLT_ACT_MUL_AUX = 4,         // D = v * aux[m,n]        (aux = saved gelu'(pre-activation), bf16)
  // Row-wise epilogues; need the whole row in one tile (N <= 256, N % 32 == 0) and a bf16 output:
  MVLT_ACT_SOFTMAX = 5,         // D = softmax_n(alpha * acc)                        (attention probabilities)
  MVLT_ACT_SOFTMAX_BWD = 6,     // D = alpha * aux * (acc - sum_n aux * acc)         (aux = P; acc = dP -> dS)
};

// MVLT_CONV_PATCH_A: A := the patch matrix of a convolution with kernel = stride = conv_R (PVLT's spatial-reduction and
// patch-embedding convolutions, libs/pvlt.py:103-104,168) over an NHWC bf16 tensor X[b, y, x, c] with pixel stride conv_C:
// patches[(b, oy, ox), (ky*R + kx)*C + c] = X[b, oy*R + ky, ox*R + kx, c] is never materialised -- a 5-D tensor map
// (kx*C + c | ox | ky | oy | b) describes it in place and one box per k-block lands as the usual K-major operand tile.
// Needs (H/R)*(W/R) == 64 output pixels per image and (R*C) % 64 == 0; M = B*64, K = R*R*C, a_mn = 0.
enum { MVLT_CONV_NONE = 0, MVLT_CONV_A = 1, MVLT_CONV_BT = 2, MVLT_CONV_PATCH_A = 3 };

typedef struct mvlt_gemm_desc {
  const void* A;  // bf16
  const void* B;  // bf16
  void* D;        // bf16 or fp32 (out_f32)
  void* D2;       // optional second bf16 output (pre-activation) or NULL
  const float* bias;      // [N] fp32 or NULL; added after alpha scaling
  const void* aux;        // bf16 [.., M, N] with D's strides, for MVLT_ACT_DGELU
  const float* residual;  // fp32 with D's strides or NULL:  D = residual + rowscale * v
  const float* rowscale;  // fp32 [ceil(M / rows_per_scale)] or NULL (drop-path keep/scale per sample)
  float* rowsum;          // fp32 [M] or NULL: rowsum[m] += alpha * sum_k A[m,k] (atomic), computed on the tensor core by one
                          // extra N=16 MMA against a tile of ones: the bias gradient db = dY^T 1 rides on the dW = dY^T X GEMM
  int32_t M, N, K;
  int32_t a_mn, b_mn;
  int64_t lda, ldb, ldd;  // leading dimensions in ELEMENTS
  int32_t batch1, batch2;  // >= 1 each; total batch = batch1 * batch2
  int64_t sA1, sA2, sB1, sB2, sD1, sD2;  // batch strides in elements (0 allowed = broadcast)
  float alpha;
  int32_t act;
  int32_t out_f32;         // 1: D is fp32, 0: bf16
  int32_t atomic_add;      // 1: D (fp32) += v via red.global.add (split-K / grad accumulation)
  int32_t rows_per_scale;  // rows of D per rowscale entry
  int32_t split_k;         // 0/1 = no split; >1 requires atomic_add
  int32_t block_n;         // 0 = auto
  // implicit 3x3 convolution operand (see above); the replaced operand pointer (A or B) is the NHWC base
  int32_t conv_mode;
  int32_t conv_B, conv_H, conv_W, conv_C;           // C % 64 == 0; W <= 64 and 64 % W == 0; (H*W) % 64 == 0
  int64_t conv_pix_stride, conv_batch_stride;       // element strides of X (x -> x+1, b -> b+1); y stride = W * pix_stride
  // LayerNorm of the OUTPUT rows fused into the residual epilogue (active when ln_gamma != NULL). Needs: residual, fp32 D,
  // N == 64 or 128 == the tile width (the whole row in one tile), no batch, no split-K, 16-byte aligned D / D2, and
  // D2 = bf16 [M, N] with D's leading dimension:  D2[m, :] = (D[m, :] - mean_m) * rstd_m * ln_gamma + ln_beta
  // (statistics over the N columns of the final fp32 row, biased variance, rstd = rsqrt(var + ln_eps));
  // ln_mean / ln_rstd: optional fp32 [M] outputs for the LayerNorm backward. Replaces a separate LayerNorm pass over D
  // (reference: x = x + drop_path(attn(norm1(x))); norm2(x), libs/pvlt.py:140-143).
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  int32_t conv_R;          // MVLT_CONV_PATCH_A / patch_store: kernel = stride of the convolution (conv_H / conv_W: INPUT map size)
  // patch_store != 0: the OUTPUT is the same patch view (the input gradient of that convolution, i.e. "unpatchify" fused into the
  // store): D is the fp32 NHWC tensor dX[b, y, x, c] (pixel stride conv_C, image stride conv_batch_stride) and element
  // (m = (b, oy, ox), n = (ky*R + kx)*C + c) of the product is written to dX[b, oy*R + ky, ox*R + kx, c] -- every pixel exactly
  // once. Needs out_f32, no epilogue operands, M = conv_B*64, N = R*R*C, the geometry rules of MVLT_CONV_PATCH_A; ldd is ignored.
  int32_t patch_store;
} mvlt_gemm_desc;

#ifdef __cplusplus
}
#endif
