// Losses, label compaction (masked-token gather/scatter), the small ITM/CLS linear heads and retrieval ranking.
//
// Reference ops replaced:
//   CrossEntropyLoss(ignore_index=-1) / CrossEntropyLoss()   /root/reference/engine_grid_masking.py:84,90,94-95
//   preds[target != index] boolean selects                    libs/vl_scores.py:16-18
//   ITMHead / CLSHead final Linear + two biases                libs/vl_heads.py:84-87,101-104
//   softmax -> sort -> rank of candidate 0                     engine_grid_masking.py:360-384
#include "common.cuh"

namespace {

// ---- deterministic compaction of labelled positions: idx[k] = k-th i with labels[i] != ignore -----------------
__global__ void __launch_bounds__(1024) compact_labels_kernel(const long long* __restrict__ labels, int n, long long ignore,
                                                              int* __restrict__ idx, long long* __restrict__ lab_out,
                                                              int* __restrict__ count, int cap, float* __restrict__ count_f32,
                                                              float* __restrict__ inv_count, float* __restrict__ overflow) {
  pdl_prologue();
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += blockDim.x) {
    const int i = start + threadIdx.x;
    const long long lab = i < n ? labels[i] : ignore;
    const int flag = (lab != ignore) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (flag && (cap <= 0 || off + within < cap)) {
      idx[off + within] = i;
      if (lab_out) lab_out[off + within] = lab;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
      base += t;
    }
    __syncthreads();
  }
  // fixed-capacity mode (cap > 0): the consumers run on exactly `cap` rows whatever the count is (static shapes: CUDA graphs,
  // no device->host read of the count): the tail is padded with row 0 / the ignore label, which contributes nothing
  const int cnt = base;
  for (int k = cnt + (int)threadIdx.x; k < cap; k += blockDim.x) {
    idx[k] = 0;
    if (lab_out) lab_out[k] = ignore;
  }
  if (threadIdx.x == 0) {
    *count = cnt;
    if (count_f32) *count_f32 = (float)cnt;
    if (inv_count) *inv_count = 1.0f / (float)(cnt > 0 ? cnt : 1);
    if (overflow && cap > 0 && cnt > cap) *overflow = (float)cnt;   // more labelled rows than the capacity: the caller must check
  }
}

struct RowMap3 { int group, stride, offset; };
__device__ __forceinline__ long long map_row3(const RowMap3& m, long long r) {
  return (r / m.group) * m.stride + m.offset + (r % m.group);
}

// dst[i, :] = src[map(idx[i]), :]
template <typename TO>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, RowMap3 sm, long long lds,
                                                          const int* __restrict__ idx, int n_idx, TO* __restrict__ dst, int C) {
  pdl_prologue();
  const int c4n = C / 4;
  const long long total = (long long)n_idx * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c4n) * 4;
    const int r = (int)((unsigned int)i / (unsigned int)c4n);
    const float4 v = *reinterpret_cast<const float4*>(src + map_row3(sm, idx[r]) * lds + c);
    if constexpr (sizeof(TO) == 2) {
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(dst + (long long)r * C + c) = u;
    } else {
      *reinterpret_cast<float4*>(dst + (long long)r * C + c) = v;
    }
  }
}
// dst[map(idx[i]), :] (+)= src[i, :]   (idx entries are unique)
template <typename TI>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const TI* __restrict__ src, const int* __restrict__ idx, int n_idx,
                                                           float* __restrict__ dst, RowMap3 dm, long long ldd, int C,
                                                           int accumulate) {
  pdl_prologue();
  const int c4n = C / 4;
  const long long total = (long long)n_idx * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c4n) * 4;
    const int r = (int)((unsigned int)i / (unsigned int)c4n);
    float4 v;
    if constexpr (sizeof(TI) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(src + (long long)r * C + c);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      v = make_float4(a.x, a.y, b.x, b.y);
    } else {
      v = *reinterpret_cast<const float4*>(src + (long long)r * C + c);
    }
    float4* d = reinterpret_cast<float4*>(dst + map_row3(dm, idx[r]) * ldd + c);
    if (accumulate) {
      const float4 o = *d;
      v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *d = v;
  }
}

// ---- cross entropy: one block per row ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float ldf(const T* p) {
  if constexpr (sizeof(T) == 2) return __bfloat162float(*p);
  else return *p;
}

template <typename T>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                                                     int n_cls, long long ignore, float* __restrict__ lse_out,
                                                     float* __restrict__ loss_sum, float* __restrict__ total_sum, float scale,
                                                     int* __restrict__ argmax_out, float* __restrict__ correct,
                                                     const float* __restrict__ scale_dev) {
  pdl_prologue();
  if (scale_dev != nullptr) scale *= *scale_dev;   // e.g. 1 / (#labelled rows) computed on the device by compact_labels
  __shared__ float sh[32];
  __shared__ int shi[32];
  const int r = blockIdx.x;
  const T* row = logits + (long long)r * ld;
  float m = -INFINITY;
  int am = 0;
  // bf16 rows with a 16-byte aligned pitch (the 30522-wide vocabulary logits): 8 logits per 16-byte load
  const bool vec8 = (sizeof(T) == 2) && ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int n8 = vec8 ? (n_cls >> 3) : 0;
  for (int i = threadIdx.x; i < n8; i += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      if (f.x > m) { m = f.x; am = 8 * i + 2 * j; }
      if (f.y > m) { m = f.y; am = 8 * i + 2 * j + 1; }
    }
  }
  for (int c = 8 * n8 + threadIdx.x; c < n_cls; c += blockDim.x) {
    const float v = ldf(row + c);
    if (v > m) { m = v; am = c; }
  }
  // block argmax (first index wins ties, like torch.argmax on CPU for our test sizes)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  if (lane == 0) { sh[warp] = m; shi[warp] = am; }
  __syncthreads();
  if (warp == 0) {
    float mm = lane < (int)(blockDim.x >> 5) ? sh[lane] : -INFINITY;
    int aa = lane < (int)(blockDim.x >> 5) ? shi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mm, o);
      const int oa = __shfl_xor_sync(0xffffffffu, aa, o);
      if (om > mm || (om == mm && oa < aa)) { mm = om; aa = oa; }
    }
    if (lane == 0) { sh[0] = mm; shi[0] = aa; }
  }
  __syncthreads();
  m = sh[0];
  am = shi[0];
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n8; i += blockDim.x) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + 8 * i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      s += __expf(f.x - m) + __expf(f.y - m);
    }
  }
  for (int c = 8 * n8 + threadIdx.x; c < n_cls; c += blockDim.x) s += __expf(ldf(row + c) - m);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) {
    const float lse = m + logf(s);
    lse_out[r] = lse;
    const long long lab = labels[r];
    if (argmax_out) argmax_out[r] = am;
    if (lab != ignore) {
      const float l = (lse - ldf(row + lab)) * scale;
      atomicAdd(loss_sum, l);
      if (total_sum) atomicAdd(total_sum, l);
      if (correct && am == (int)lab) atomicAdd(correct, 1.0f);
    }
  }
}

// dlogits = (softmax - onehot) * scale * g   (0 for ignored rows); may alias logits
template <typename T>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                                                     int n_cls, long long ignore, const float* __restrict__ lse,
                                                     T* __restrict__ dlogits, long long ldd, float scale,
                                                     const float* __restrict__ gscale, const float* __restrict__ scale_dev) {
  pdl_prologue();
  if (scale_dev != nullptr) scale *= *scale_dev;
  const int r = blockIdx.x;
  const long long lab = labels[r];
  const float g = scale * (gscale ? *gscale : 1.f);
  const float l = lse[r];
  const T* row = logits + (long long)r * ld;
  T* drow = dlogits + (long long)r * ldd;
  const bool vec8 = (sizeof(T) == 2) && ((ld & 7) == 0) && ((ldd & 7) == 0) &&
                    (((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(dlogits)) & 15) == 0);
  const int n8 = vec8 ? (n_cls >> 3) : 0;
  for (int i = threadIdx.x; i < n8; i += blockDim.x) {
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (lab != ignore) {
      u = *reinterpret_cast<const uint4*>(row + 8 * i);
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        const int c = 8 * i + 2 * j;
        w[j] = pack_bf16x2((__expf(f.x - l) - (c == lab ? 1.f : 0.f)) * g, (__expf(f.y - l) - (c + 1 == lab ? 1.f : 0.f)) * g);
      }
      u = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4*>(drow + 8 * i) = u;
  }
  for (int c = 8 * n8 + threadIdx.x; c < n_cls; c += blockDim.x) {
    float v = 0.f;
    if (lab != ignore) v = (__expf(ldf(row + c) - l) - (c == lab ? 1.f : 0.f)) * g;
    if constexpr (sizeof(T) == 2) drow[c] = __float2bfloat16(v);
    else drow[c] = v;
  }
}

// ---- small linear head: logits[m, n] = h[m,:] . W[n,:] + b1[n] + b2[n]  (n = 2 / 48 / 122) ---------------------
__global__ void __launch_bounds__(256) small_linear_fwd_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ W,
                                                               const float* __restrict__ b1, const float* __restrict__ b2,
                                                               float* __restrict__ out, int M, int n, int K) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= M * n) return;
  const int m = wid / n, j = wid % n;
  float acc = 0.f;
  for (int k = lane * 2; k < K; k += 64) {
    const float2 hv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(h + (long long)m * K + k));
    const float2 wv = *reinterpret_cast<const float2*>(W + (long long)j * K + k);
    acc += hv.x * wv.x + hv.y * wv.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(long long)m * n + j] = acc + (b1 ? b1[j] : 0.f) + (b2 ? b2[j] : 0.f);
}
// dh[m,k] = sum_n dl[m,n] W[n,k]  (bf16 out);  dW[n,k] += sum_m dl[m,n] h[m,k];  db1[n], db2[n] += sum_m dl[m,n]
__global__ void __launch_bounds__(256) small_linear_bwd_kernel(const float* __restrict__ dl, const __nv_bfloat16* __restrict__ h,
                                                               const float* __restrict__ W, __nv_bfloat16* __restrict__ dh,
                                                               float* __restrict__ dW, float* __restrict__ db1,
                                                               float* __restrict__ db2, int M, int n, int K) {
  pdl_prologue();
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long n_dh = (long long)M * K, n_dw = (long long)n * K;
  if (tid < n_dh) {
    const int m = (int)(tid / K), k = (int)(tid % K);
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += dl[(long long)m * n + j] * W[(long long)j * K + k];
    dh[tid] = __float2bfloat16(acc);
  } else if (tid < n_dh + n_dw) {
    const long long t = tid - n_dh;
    const int j = (int)(t / K), k = (int)(t % K);
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc += dl[(long long)m * n + j] * __bfloat162float(h[(long long)m * K + k]);
    dW[t] += acc;
  } else if (tid < n_dh + n_dw + n) {
    const int j = (int)(tid - n_dh - n_dw);
    float acc = 0.f;
    for (int m = 0; m < M; ++m) acc += dl[(long long)m * n + j];
    if (db1) db1[j] += acc;
    if (db2) db2[j] += acc;
  }
}

// ---- retrieval: rank of candidate 0 under descending p(match) -------------------------------------------------
__global__ void itm_rank_kernel(const float* __restrict__ logits, int n_query, int n_cand, int* __restrict__ rank_out,
                                float* __restrict__ prob_out) {
  pdl_prologue();
  __shared__ float sh[32];
  const int q = blockIdx.x;
  const float* lq = logits + (long long)q * n_cand * 2;
  auto pmatch = [&](int j) {
    const float a = lq[2 * j], b = lq[2 * j + 1];
    const float m = fmaxf(a, b);
    const float ea = expf(a - m), eb = expf(b - m);
    return eb / (ea + eb);
  };
  const float p0 = pmatch(0);
  float cnt = 0.f;
  for (int j = threadIdx.x; j < n_cand; j += blockDim.x) {
    const float p = pmatch(j);
    if (prob_out) prob_out[(long long)q * n_cand + j] = p;
    if (p > p0) cnt += 1.f;
  }
  cnt = block_sum(cnt, sh);
  if (threadIdx.x == 0) rank_out[q] = (int)(cnt + 0.5f);
}

// sum of squared differences (compute_psnr, vl_scores.py:54-63)
__global__ void __launch_bounds__(256) sq_diff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                          float* __restrict__ out) {
  pdl_prologue();
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s += d * d;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

inline int cap_grid(long long work_items, int threads, int per_sm = 8) {
  // the kernels decompose their linear work index with 32-bit arithmetic: refuse (grid 0 -> launch error) beyond that
  if (work_items >= (1ll << 32)) {
    mvlt_set_error("work size %lld exceeds the 32-bit index range of the elementwise kernels", work_items);
    return 0;
  }
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)mvlt_num_sms() * per_sm;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

}  // namespace

// cap > 0: fixed-capacity mode: idx_out / labels_out hold `cap` entries, the tail beyond the count is padded (row 0, ignore label);
// count_f32_out / inv_count_out / overflow_out (optional device floats) receive the count, 1 / max(count, 1) and -- only when the
// count exceeds the capacity -- the count (left untouched otherwise).
extern "C" int mvlt_compact_labels(const long long* labels, int n, long long ignore, int* idx_out, long long* labels_out,
                                   int* count_out, int cap, float* count_f32_out, float* inv_count_out, float* overflow_out,
                                   void* stream_) {
  MVLT_CHECK_ARG(n > 0 && cap >= 0, "compact_labels: n must be positive");
  mvlt_launch(compact_labels_kernel, 1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_), labels, n, ignore, idx_out, labels_out,
                                                                                 count_out, cap, count_f32_out, inv_count_out, overflow_out);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_gather_rows(const float* src, const int* smap, long long lds, const int* idx, int n_idx, void* dst,
                                int dst_f32, int C, void* stream_) {
  MVLT_CHECK_ARG(C % 4 == 0 && n_idx > 0, "gather_rows: bad C / n_idx");
  const int big = 1 << 30;
  RowMap3 sm{smap && smap[0] > 0 ? smap[0] : big, smap && smap[0] > 0 ? smap[1] : big, smap && smap[0] > 0 ? smap[2] : 0};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = cap_grid((long long)n_idx * (C / 4), 256);
  if (dst_f32) mvlt_launch(gather_rows_kernel<float>, grid, 256, 0, st, src, sm, lds, idx, n_idx, reinterpret_cast<float*>(dst), C);
  else mvlt_launch(gather_rows_kernel<__nv_bfloat16>, grid, 256, 0, st, src, sm, lds, idx, n_idx, reinterpret_cast<__nv_bfloat16*>(dst), C);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_scatter_rows(const void* src, int src_f32, const int* idx, int n_idx, float* dst, const int* dmap,
                                 long long ldd, int C, int accumulate, void* stream_) {
  MVLT_CHECK_ARG(C % 4 == 0 && n_idx > 0, "scatter_rows: bad C / n_idx");
  const int big = 1 << 30;
  RowMap3 dm{dmap && dmap[0] > 0 ? dmap[0] : big, dmap && dmap[0] > 0 ? dmap[1] : big, dmap && dmap[0] > 0 ? dmap[2] : 0};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = cap_grid((long long)n_idx * (C / 4), 256);
  if (src_f32) mvlt_launch(scatter_rows_kernel<float>, grid, 256, 0, st, reinterpret_cast<const float*>(src), idx, n_idx, dst, dm, ldd, C, accumulate);
  else mvlt_launch(scatter_rows_kernel<__nv_bfloat16>, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(src), idx, n_idx, dst, dm, ldd, C, accumulate);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_ce_fwd(const void* logits, int logits_f32, long long ld, const long long* labels, int rows, int n_cls,
                           long long ignore, float* lse, float* loss_sum, float* total_sum, float scale, int* argmax_out,
                           float* correct, const float* scale_dev, void* stream_) {
  MVLT_CHECK_ARG(rows > 0 && n_cls > 0, "ce_fwd: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (logits_f32)
    mvlt_launch(ce_fwd_kernel<float>, rows, 256, 0, st, reinterpret_cast<const float*>(logits), ld, labels, n_cls, ignore, lse,
                                               loss_sum, total_sum, scale, argmax_out, correct, scale_dev);
  else
    mvlt_launch(ce_fwd_kernel<__nv_bfloat16>, rows, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(logits), ld, labels, n_cls,
                                                       ignore, lse, loss_sum, total_sum, scale, argmax_out, correct, scale_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_ce_bwd(const void* logits, int logits_f32, long long ld, const long long* labels, int rows, int n_cls,
                           long long ignore, const float* lse, void* dlogits, long long ldd, float scale,
                           const float* gscale_dev, const float* scale_dev, void* stream_) {
  MVLT_CHECK_ARG(rows > 0 && n_cls > 0, "ce_bwd: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (logits_f32)
    mvlt_launch(ce_bwd_kernel<float>, rows, 256, 0, st, reinterpret_cast<const float*>(logits), ld, labels, n_cls, ignore, lse,
                                               reinterpret_cast<float*>(dlogits), ldd, scale, gscale_dev, scale_dev);
  else
    mvlt_launch(ce_bwd_kernel<__nv_bfloat16>, rows, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(logits), ld, labels, n_cls,
                                                       ignore, lse, reinterpret_cast<__nv_bfloat16*>(dlogits), ldd, scale,
                                                       gscale_dev, scale_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_small_linear_fwd(const void* h_bf16, const float* W, const float* b1, const float* b2, float* out,
                                     int M, int n, int K, void* stream_) {
  MVLT_CHECK_ARG(K % 2 == 0 && M > 0 && n > 0, "small_linear_fwd: bad shape");
  const long long warps = (long long)M * n;
  mvlt_launch(small_linear_fwd_kernel, (int)((warps + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(h_bf16), W, b1, b2, out, M, n, K);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_small_linear_bwd(const float* dlogits, const void* h_bf16, const float* W, void* dh_bf16, float* dW,
                                     float* db1, float* db2, int M, int n, int K, void* stream_) {
  const long long total = (long long)M * K + (long long)n * K + n;
  mvlt_launch(small_linear_bwd_kernel, (int)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), dlogits, reinterpret_cast<const __nv_bfloat16*>(h_bf16), W, reinterpret_cast<__nv_bfloat16*>(dh_bf16), dW, db1, db2,
      M, n, K);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_itm_rank(const float* logits, int n_query, int n_cand, int* rank_out, float* prob_out, void* stream_) {
  MVLT_CHECK_ARG(n_query > 0 && n_cand > 0, "itm_rank: bad shape");
  mvlt_launch(itm_rank_kernel, n_query, 128, 0, reinterpret_cast<cudaStream_t>(stream_), logits, n_query, n_cand, rank_out, prob_out);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_sq_diff_sum(const float* a, const float* b, long long n, float* out, void* stream_) {
  mvlt_launch(sq_diff_sum_kernel, cap_grid(n, 256, 4), 256, 0, reinterpret_cast<cudaStream_t>(stream_), a, b, n, out);
  MVLT_CHECK_LAUNCH();
  return 0;
}
