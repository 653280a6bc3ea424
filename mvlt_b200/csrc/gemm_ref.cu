// One-thread-per-output SIMT GEMM with exactly the semantics of mvlt_gemm (gemm_desc.h).
// It exists ONLY as an on-device cross-check for the tcgen05 kernel in tests/ (and to bisect a bad layout
// on the GPU box); no model path calls it.
#include "common.cuh"
#include "gemm_desc.h"

namespace {

__global__ void gemm_ref_kernel(const mvlt_gemm_desc g) {
  pdl_prologue();
  const long long total = (long long)g.batch1 * g.batch2 * g.M * g.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % g.N);
    long long t = i / g.N;
    const int m = (int)(t % g.M);
    t /= g.M;
    const int b2 = (int)(t % g.batch2);
    const int b1 = (int)(t / g.batch2);
    const __nv_bfloat16* A = reinterpret_cast<const __nv_bfloat16*>(g.A) + b1 * g.sA1 + b2 * g.sA2;
    const __nv_bfloat16* B = reinterpret_cast<const __nv_bfloat16*>(g.B) + b1 * g.sB1 + b2 * g.sB2;
    float acc = 0.f;
    for (int k = 0; k < g.K; ++k) {
      const float a = __bfloat162float(g.a_mn ? A[(long long)k * g.lda + m] : A[(long long)m * g.lda + k]);
      const float b = __bfloat162float(g.b_mn ? B[(long long)k * g.ldb + n] : B[(long long)n * g.ldb + k]);
      acc = fmaf(a, b, acc);
    }
    const long long off = b1 * g.sD1 + b2 * g.sD2 + (long long)m * g.ldd + n;
    float v = acc * g.alpha;
    if (g.bias) v += g.bias[n];
    if (g.act == MVLT_ACT_GELU) {
      if (g.D2) reinterpret_cast<__nv_bfloat16*>(g.D2)[off] = __float2bfloat16(v);
      v = gelu_erf(v);
    } else if (g.act == MVLT_ACT_GELU_SAVE_GRAD) {
      reinterpret_cast<__nv_bfloat16*>(g.D2)[off] = __float2bfloat16(dgelu_erf(v));
      v = gelu_erf(v);
    } else if (g.act == MVLT_ACT_MUL_AUX) {
      v *= __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g.aux)[off]);
    } else if (g.act == MVLT_ACT_DGELU) {
      v *= dgelu_erf(__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g.aux)[off]));
    }
    const float rs = g.rowscale ? g.rowscale[m / g.rows_per_scale] : 1.f;
    if (g.residual) v = g.residual[off] + rs * v;
    else v *= rs;
    if (g.atomic_add) atomicAdd(reinterpret_cast<float*>(g.D) + off, v);
    else if (g.out_f32) reinterpret_cast<float*>(g.D)[off] = v;
    else reinterpret_cast<__nv_bfloat16*>(g.D)[off] = __float2bfloat16(v);
  }
}

}  // namespace

extern "C" int mvlt_gemm_ref(const mvlt_gemm_desc* g, void* stream) {
  MVLT_CHECK_ARG(g && g->A && g->B && g->D, "mvlt_gemm_ref: null argument");
  MVLT_CHECK_ARG(g->rowsum == nullptr, "mvlt_gemm_ref: rowsum is only implemented by the tcgen05 kernel");
  MVLT_CHECK_ARG(g->ln_gamma == nullptr, "mvlt_gemm_ref: the fused LayerNorm epilogue is only implemented by the tcgen05 kernel");
  MVLT_CHECK_ARG(g->conv_mode == MVLT_CONV_NONE, "mvlt_gemm_ref: implicit convolution operands are only implemented by the tcgen05 kernel");
  MVLT_CHECK_ARG(g->act != MVLT_ACT_SOFTMAX && g->act != MVLT_ACT_SOFTMAX_BWD,
                 "mvlt_gemm_ref: the row-wise softmax epilogues are only implemented by the tcgen05 kernel");
  const long long total = (long long)g->batch1 * g->batch2 * g->M * g->N;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  mvlt_launch(gemm_ref_kernel, (int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream), *g);
  MVLT_CHECK_LAUNCH();
  return 0;
}
