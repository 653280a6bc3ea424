// Multi-tensor AdamW: one launch updates every parameter of a group (decoupled weight decay, bias correction),
// reading p, g, m, v once and writing p, m, v once: 28 B per parameter, HBM-bound.
//
// Replaces the per-step optimizer work the reference delegates to timm.create_optimizer(opt='adamw') + torch.optim
// (/root/reference/main_vl.py:308, stepped at engine_grid_masking.py:122-127 through timm's NativeScaler): same
// update rule as torch.optim.AdamW(amsgrad=False, maximize=False):
//   p <- p * (1 - lr * wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;
//   p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "common.cuh"

struct AdamTensor {      // one parameter tensor (88 bytes; mirrored by mvlt_b200/optim.py)
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
  float wd_mult;         // multiplies the group's weight decay (0 for bias / 1-D parameters)
  int shadow;            // bf16 compute copies the engine's GEMMs read, refreshed by this kernel (0 = none):
                         //   1: w16[i] = bf16(p[i])                                  (Linear / embedding weights)
                         //   2: p = [Co, Ci, KK]: w16[co * ld + kk * Ci + ci]        (convolutions as GEMMs, K = (tap, ci))
                         //      and, when w16t != nullptr, w16t[ci, (KK-1-kk) * Co + co]  (flipped + transposed: input gradient)
  __nv_bfloat16* w16;
  __nv_bfloat16* w16t;
  int Co, Ci, KK, ld;
};
struct AdamChunk {       // one contiguous run of <= chunk_elems elements of tensor `t`
  long long off;
  int t;
  int pad;
};

namespace {

__device__ __forceinline__ void shadow_conv(const AdamTensor& t, long long e, float val) {
  const int kk = (int)(e % t.KK);
  const int ci = (int)((e / t.KK) % t.Ci);
  const int co = (int)(e / ((long long)t.Ci * t.KK));
  const __nv_bfloat16 b = __float2bfloat16(val);
  t.w16[(long long)co * t.ld + (long long)kk * t.Ci + ci] = b;
  if (t.w16t) t.w16t[((long long)ci * t.KK + (t.KK - 1 - kk)) * t.Co + co] = b;
}

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr_wd, float b1, float b2, float step, float rbc2,
                                      float eps) {
  p -= p * lr_wd;
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p -= step * m / (sqrtf(v) * rbc2 + eps);
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamTensor* __restrict__ tensors, const AdamChunk* __restrict__ chunks,
                                                          int chunk_elems, float lr, float b1, float b2, float eps, float wd,
                                                          float bc1, float bc2, const float* __restrict__ grad_scale, int zero_grads,
                                                          const float* __restrict__ hyper) {
  pdl_prologue();
  if (hyper != nullptr) {   // {lr, bias_correction1, bias_correction2} read from device memory (CUDA-graph replays, schedulers)
    lr = hyper[0];
    bc1 = hyper[1];
    bc2 = hyper[2];
  }
  const AdamChunk ch = chunks[blockIdx.x];
  const AdamTensor t = tensors[ch.t];
  const long long n = min((long long)chunk_elems, t.n - ch.off);
  float* p = t.p + ch.off;
  float* g = t.g + ch.off;
  float* m = t.m + ch.off;
  float* v = t.v + ch.off;
  const float gs = grad_scale ? *grad_scale : 1.f;
  const float lr_wd = lr * wd * t.wd_mult, step = lr / bc1, rbc2 = rsqrtf(bc2);
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0 &&
                   (t.shadow != 1 || (((uintptr_t)t.w16) & 7) == 0);
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    adam1(p4.x, g4.x * gs, m4.x, v4.x, lr_wd, b1, b2, step, rbc2, eps);
    adam1(p4.y, g4.y * gs, m4.y, v4.y, lr_wd, b1, b2, step, rbc2, eps);
    adam1(p4.z, g4.z * gs, m4.z, v4.z, lr_wd, b1, b2, step, rbc2, eps);
    adam1(p4.w, g4.w * gs, m4.w, v4.w, lr_wd, b1, b2, step, rbc2, eps);
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    if (zero_grads) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t.shadow == 1) {          // ch.off is a multiple of 4 elements and w16 is 8-byte aligned (checked by the host)
      uint2 o;
      o.x = pack_bf16x2(p4.x, p4.y);
      o.y = pack_bf16x2(p4.z, p4.w);
      reinterpret_cast<uint2*>(t.w16 + ch.off)[i] = o;
    } else if (t.shadow == 2) {
      shadow_conv(t, ch.off + 4 * i, p4.x);
      shadow_conv(t, ch.off + 4 * i + 1, p4.y);
      shadow_conv(t, ch.off + 4 * i + 2, p4.z);
      shadow_conv(t, ch.off + 4 * i + 3, p4.w);
    }
  }
  for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam1(pp, g[i] * gs, mm, vv, lr_wd, b1, b2, step, rbc2, eps);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grads) g[i] = 0.f;
    if (t.shadow == 1) t.w16[ch.off + i] = __float2bfloat16(pp);
    else if (t.shadow == 2) shadow_conv(t, ch.off + i, pp);
  }
}

// ---- gradient-norm clipping folded into the optimizer step (timm NativeScaler / torch.nn.utils.clip_grad_norm_ semantics,
// engine_grid_masking.py:126-127 with --clip-grad): the gradients are READ once more (4 B per parameter) for their squared sum,
// never rewritten: the clip coefficient is a device scalar that adamw_multi multiplies into every gradient as it loads it.
// Deterministic (no atomics): one partial per chunk, summed in a fixed order, so data-parallel ranks holding identical
// gradients compute bit-identical coefficients and their parameters stay bit-identical.
__global__ void __launch_bounds__(256) grad_sumsq_multi_kernel(const AdamTensor* __restrict__ tensors, const AdamChunk* __restrict__ chunks,
                                                               int chunk_elems, float* __restrict__ partials) {
  pdl_prologue();
  __shared__ float sh[32];
  const AdamChunk ch = chunks[blockIdx.x];
  const AdamTensor t = tensors[ch.t];
  const long long n = min((long long)chunk_elems, t.n - ch.off);
  const float* g = t.g + ch.off;
  float s = 0.f;
  const long long n4 = ((((uintptr_t)g) & 15) == 0) ? n / 4 : 0;
  for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    s = fmaf(g4.x, g4.x, s);
    s = fmaf(g4.y, g4.y, s);
    s = fmaf(g4.z, g4.z, s);
    s = fmaf(g4.w, g4.w, s);
  }
  for (long long i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s = fmaf(g[i], g[i], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// out[0] = in_scale * min(1, max_norm / (norm + 1e-6)), out[1] = norm = |in_scale| * sqrt(sum of the partials)
__global__ void __launch_bounds__(256) grad_clip_scale_kernel(const float* __restrict__ partials, int n, float max_norm,
                                                              const float* __restrict__ in_scale, float* __restrict__ out) {
  pdl_prologue();
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float is = in_scale ? *in_scale : 1.f;
    const float norm = (float)sqrt(sh[0]) * fabsf(is);
    const float coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
    out[0] = coef * is;
    out[1] = norm;
  }
}

}  // namespace

// tensors: device array of AdamTensor; chunks: device array of n_chunks AdamChunk (one CTA each).
// hyper_dev (optional): device float[3] = {lr, bias_correction1, bias_correction2} overriding the by-value arguments at
// execution time (the values a captured CUDA graph must pick up on every replay)
extern "C" int mvlt_adamw_multi(const void* tensors, const void* chunks, int n_chunks, int chunk_elems, float lr, float beta1,
                                float beta2, float eps, float weight_decay, float bias_correction1, float bias_correction2,
                                const float* grad_scale_dev, int zero_grads, const float* hyper_dev, void* stream_) {
  MVLT_CHECK_ARG(tensors && chunks && n_chunks > 0 && chunk_elems > 0, "adamw_multi: empty tables");
  MVLT_CHECK_ARG(bias_correction1 > 0.f && bias_correction2 > 0.f, "adamw_multi: bias corrections must be positive");
  mvlt_launch(adamw_multi_kernel, n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const AdamTensor*>(tensors), reinterpret_cast<const AdamChunk*>(chunks), chunk_elems, lr, beta1, beta2, eps,
      weight_decay, bias_correction1, bias_correction2, grad_scale_dev, zero_grads, hyper_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// partials_out[c] = sum of squared gradients of chunk c (the tables of mvlt_adamw_multi; n_chunks floats are written).
extern "C" int mvlt_grad_sumsq_multi(const void* tensors, const void* chunks, int n_chunks, int chunk_elems, float* partials_out,
                                     void* stream_) {
  MVLT_CHECK_ARG(tensors && chunks && partials_out && n_chunks > 0 && chunk_elems > 0, "grad_sumsq_multi: empty tables");
  mvlt_launch(grad_sumsq_multi_kernel, n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const AdamTensor*>(tensors), reinterpret_cast<const AdamChunk*>(chunks), chunk_elems, partials_out);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// Gradient-clipping coefficient from the chunk partials of every group (summed in a fixed order):
//   out2[1] = total_norm = |s| * sqrt(sum partials),  out2[0] = s * min(1, max_norm / (total_norm + 1e-6))     (s = *in_scale_dev or 1)
// out2[0] is the grad_scale_dev operand of mvlt_adamw_multi (torch.nn.utils.clip_grad_norm_ semantics; max_norm <= 0: no clipping).
extern "C" int mvlt_grad_clip_scale(const float* partials, int n_partials, float max_norm, const float* in_scale_dev, float* out2,
                                    void* stream_) {
  MVLT_CHECK_ARG(partials && out2 && n_partials > 0, "grad_clip_scale: empty partials");
  mvlt_launch(grad_clip_scale_kernel, 1, 256, 0, reinterpret_cast<cudaStream_t>(stream_), partials, n_partials, max_norm, in_scale_dev, out2);
  MVLT_CHECK_LAUNCH();
  return 0;
}
