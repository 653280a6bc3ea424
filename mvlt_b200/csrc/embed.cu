// BERT text embeddings: LN_{1e-12}(word[ids] + pos[t] + type[0]) (+ dropout), one warp per token, fused gather +
// LayerNorm; and its backward (LN backward + scatter-add into the three embedding tables).
// Reference: transformers BertEmbeddings.forward, called at /root/reference/libs/pvlt.py:326 (ctor :232-233).
#include "common.cuh"

namespace {

constexpr int HID = 768;
constexpr int NV = HID / 128;  // float4 per lane

__device__ __forceinline__ void gather_row(const float* __restrict__ word, const float* __restrict__ pos,
                                           const float* __restrict__ type, long long id, int t, int lane, float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(word + id * HID + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(pos + (long long)t * HID + c));
    const float4 d = __ldg(reinterpret_cast<const float4*>(type + c));
    v[i] = make_float4(a.x + b.x + d.x, a.y + b.y + d.y, a.z + b.z + d.z, a.w + b.w + d.w);
  }
}

__global__ void __launch_bounds__(256) bert_embed_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ word,
                                                             const float* __restrict__ pos, const float* __restrict__ type,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, int rows, int T, float eps,
                                                             float p_drop, unsigned long long seed,
                                                             const unsigned long long* __restrict__ seed_dev) {
  pdl_prologue();
  if (seed_dev != nullptr) seed = *seed_dev;   // seed read from device memory: a captured CUDA graph draws new masks per replay
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    float4 v[NV];
    gather_row(word, pos, type, ids[r], r % T, lane, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
    const float mean = warp_sum(s) * (1.f / HID);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / HID) + eps);
    if (lane == 0 && mean_out) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float o[4] = {(v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                    (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w};
      if (p_drop > 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = hash_uniform(seed, (unsigned long long)r * HID + c + j) >= p_drop ? o[j] * keep_scale : 0.f;
      }
      uint2 u;
      u.x = pack_bf16x2(o[0], o[1]);
      u.y = pack_bf16x2(o[2], o[3]);
      *reinterpret_cast<uint2*>(out + (long long)r * HID + c) = u;
    }
  }
}

__global__ void __launch_bounds__(256) bert_embed_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const long long* __restrict__ ids,
                                                             const float* __restrict__ word, const float* __restrict__ pos,
                                                             const float* __restrict__ type, const float* __restrict__ gamma,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ dword, float* __restrict__ dpos,
                                                             float* __restrict__ dtype_, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, int rows, int T, float p_drop,
                                                             unsigned long long seed, int pad_id,
                                                             const unsigned long long* __restrict__ seed_dev) {
  pdl_prologue();
  if (seed_dev != nullptr) seed = *seed_dev;
  __shared__ float shg[8][HID];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  float4 ag[NV], ab[NV], at[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = at[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = blockIdx.x * wpb + warp; r < rows; r += gridDim.x * wpb) {
    const long long id = ids[r];
    const int t = r % T;
    float4 v[NV];
    gather_row(word, pos, type, id, t, lane, v);
    const float mu = mean[r], rs = rstd[r];
    float4 gd[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      const uint2 u = *reinterpret_cast<const uint2*>(dy + (long long)r * HID + c);
      const float2 d0 = unpack_bf16x2(u.x), d1 = unpack_bf16x2(u.y);
      float d[4] = {d0.x, d0.y, d1.x, d1.y};
      if (p_drop > 0.f) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          d[j] = hash_uniform(seed, (unsigned long long)r * HID + c + j) >= p_drop ? d[j] * keep_scale : 0.f;
      }
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      xh[i] = make_float4((v[i].x - mu) * rs, (v[i].y - mu) * rs, (v[i].z - mu) * rs, (v[i].w - mu) * rs);
      ag[i].x += d[0] * xh[i].x; ag[i].y += d[1] * xh[i].y; ag[i].z += d[2] * xh[i].z; ag[i].w += d[3] * xh[i].w;
      ab[i].x += d[0]; ab[i].y += d[1]; ab[i].z += d[2]; ab[i].w += d[3];
      gd[i] = make_float4(d[0] * g.x, d[1] * g.y, d[2] * g.z, d[3] * g.w);
      s1 += gd[i].x + gd[i].y + gd[i].z + gd[i].w;
      s2 += gd[i].x * xh[i].x + gd[i].y * xh[i].y + gd[i].z * xh[i].z + gd[i].w * xh[i].w;
    }
    s1 = warp_sum(s1) * (1.f / HID);
    s2 = warp_sum(s2) * (1.f / HID);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = i * 128 + lane * 4;
      const float e0 = rs * (gd[i].x - s1 - xh[i].x * s2), e1 = rs * (gd[i].y - s1 - xh[i].y * s2);
      const float e2 = rs * (gd[i].z - s1 - xh[i].z * s2), e3 = rs * (gd[i].w - s1 - xh[i].w * s2);
      if (id != pad_id) {  // nn.Embedding(padding_idx=0): the pad row receives no gradient from the gather
        float* dw = dword + id * HID + c;
        atomicAdd(dw, e0); atomicAdd(dw + 1, e1); atomicAdd(dw + 2, e2); atomicAdd(dw + 3, e3);
      }
      float* dp = dpos + (long long)t * HID + c;
      atomicAdd(dp, e0); atomicAdd(dp + 1, e1); atomicAdd(dp + 2, e2); atomicAdd(dp + 3, e3);
      at[i].x += e0; at[i].y += e1; at[i].z += e2; at[i].w += e3;
    }
  }
  // three block reductions (dgamma, dbeta, dtype row 0) through one smem buffer
  float4* parts[3] = {ag, ab, at};
  float* outs[3] = {dgamma, dbeta, dtype_};
  for (int k = 0; k < 3; ++k) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(&shg[warp][i * 128 + lane * 4]) = parts[k][i];
    __syncthreads();
    for (int c = threadIdx.x; c < HID; c += blockDim.x) {
      float s = 0.f;
      for (int w = 0; w < wpb; ++w) s += shg[w][c];
      atomicAdd(outs[k] + c, s);
    }
  }
}

// ---- stochastic-depth / dropout keep factors from the counter-based hash (common.cuh) ---------------------------------
// out[r, b] = hash_uniform(seed, r * cols + b) >= rate[r] ? 1 / (1 - rate[r]) : 0   (rate_per_row == nullptr: `rate` for all)
__global__ void keep_scale_kernel(float* __restrict__ out, int rows, int cols, const float* __restrict__ rate_per_row, float rate,
                                  unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
  pdl_prologue();
  if (seed_dev != nullptr) seed = *seed_dev;
  const long long n = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = rate_per_row ? rate_per_row[i / cols] : rate;
    out[i] = (r <= 0.f || hash_uniform(seed, (unsigned long long)i) >= r) ? 1.f / (1.f - r) : 0.f;
  }
}

}  // namespace

// Keep factors of timm's DropPath (per sample, libs/pvlt.py:135,141-142 with the rates of :197) for all residual branches of
// one step in ONE launch: out fp32 [rows, cols] (rows = 2 * #blocks, cols = batch), rate_per_row fp32 [rows] on the device.
// With rate_per_row == nullptr it writes the element-wise dropout factors bert_embed_fwd applies for (seed, rate) to a
// [rows, cols = 768] activation (transformers BertEmbeddings.dropout) -- lets tests rebuild the mask a step drew.
// seed_dev (optional, all three entry points of this file): when non-null the seed is READ FROM DEVICE MEMORY at execution time
// and `seed` is ignored, so that a captured CUDA graph draws fresh masks on every replay.
extern "C" int mvlt_keep_scale(float* out, int rows, int cols, const float* rate_per_row, float rate, unsigned long long seed,
                               const unsigned long long* seed_dev, void* stream_) {
  MVLT_CHECK_ARG(out && rows > 0 && cols > 0 && rate >= 0.f && rate < 1.f, "keep_scale: bad arguments");
  const long long n = (long long)rows * cols;
  int grid = (int)((n + 255) / 256);
  const int cap = mvlt_num_sms() * 8;
  if (grid > cap) grid = cap;
  mvlt_launch(keep_scale_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_), out, rows, cols, rate_per_row, rate, seed, seed_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_bert_embed_fwd(const long long* ids, const float* word, const float* pos, const float* type,
                                   const float* gamma, const float* beta, void* out_bf16, float* mean, float* rstd,
                                   int rows, int T, float eps, float p_drop, unsigned long long seed,
                                   const unsigned long long* seed_dev, void* stream_) {
  MVLT_CHECK_ARG(rows > 0 && T > 0 && T <= 512, "bert_embed_fwd: bad rows/T");
  int grid = (rows + 7) / 8;
  const int cap = mvlt_num_sms() * 8;
  if (grid > cap) grid = cap;
  mvlt_launch(bert_embed_fwd_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_), ids, word, pos, type, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out_bf16), mean, rstd, rows, T, eps, p_drop, seed, seed_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_bert_embed_bwd(const void* dy_bf16, const long long* ids, const float* word, const float* pos,
                                   const float* type, const float* gamma, const float* mean, const float* rstd,
                                   float* dword, float* dpos, float* dtype_, float* dgamma, float* dbeta, int rows, int T,
                                   float p_drop, unsigned long long seed, int pad_id, const unsigned long long* seed_dev, void* stream_) {
  MVLT_CHECK_ARG(rows > 0 && T > 0, "bert_embed_bwd: bad rows/T");
  int grid = (rows + 31) / 32;
  const int cap = mvlt_num_sms() * 2;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  mvlt_launch(bert_embed_bwd_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(dy_bf16), ids, word, pos, type, gamma, mean, rstd, dword, dpos, dtype_,
      dgamma, dbeta, rows, T, p_drop, seed, pad_id, seed_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}
