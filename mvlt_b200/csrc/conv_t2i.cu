// Kernels of the MVM / text-to-image reconstruction head (ITGHead): 3x3 convolutions as im2col + tcgen05 GEMM,
// train-mode BatchNorm (batch statistics, running-stat update), fused BN-apply / elementwise products written
// straight into channel slices of the concat buffers, bilinear x2 / x8 upsampling (align_corners=True) and the
// SmoothL1 reconstruction loss fused with the final x8 upsample (the [B,3,H,W] prediction is never stored on the
// loss path). Activations are NHWC bf16.
//
// Reference: /root/reference/libs/vl_heads.py:107-165 (ITGHead), loss at engine_grid_masking.py:101.
#include "common.cuh"

namespace {

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

template <typename T>
__device__ __forceinline__ uint4 load8_as_bf16(const T* p) {
  if constexpr (sizeof(T) == 2) {
    return *reinterpret_cast<const uint4*>(p);
  } else {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    uint4 u;
    u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
    u.z = pack_bf16x2(b.x, b.y); u.w = pack_bf16x2(b.z, b.w);
    return u;
  }
}

// col[(b,y,x), (ky*3+kx)*C + c] = src[b, y+ky-1, x+kx-1, c] (zero padded)
template <typename T>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const T* __restrict__ src, long long batch_stride, int pix_stride,
                                                        __nv_bfloat16* __restrict__ col, int B, int H, int W, int C) {
  pdl_prologue();
  const int c8n = C / 8;
  const long long total = (long long)B * H * W * 9 * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)((unsigned int)i % (unsigned int)c8n);
    unsigned int t = (unsigned int)i / (unsigned int)c8n;
    const int tap = (int)(t % 9); t /= 9;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      u = load8_as_bf16<T>(src + (long long)b * batch_stride + ((long long)yy * W + xx) * pix_stride + c8 * 8);
    *reinterpret_cast<uint4*>(col + (((long long)b * H + y) * W + x) * (9LL * C) + (long long)tap * C + c8 * 8) = u;
  }
}

// dsrc[b,y,x,c] (+)= sum_taps dcol[(b, y-ky+1, x-kx+1), tap*C + c]
template <typename T>
__global__ void __launch_bounds__(256) col2im3x3_kernel(const __nv_bfloat16* __restrict__ dcol, T* __restrict__ dst,
                                                        long long batch_stride, int pix_stride, int B, int H, int W, int C,
                                                        int accumulate) {
  pdl_prologue();
  const int c8n = C / 8;
  const long long total = (long long)B * H * W * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)((unsigned int)i % (unsigned int)c8n);
    unsigned int t = (unsigned int)i / (unsigned int)c8n;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int oy = y - (tap / 3 - 1), ox = x - (tap % 3 - 1);
      if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
      const uint4 u = *reinterpret_cast<const uint4*>(dcol + (((long long)b * H + oy) * W + ox) * (9LL * C) + (long long)tap * C + c8 * 8);
      float2 f;
      f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
      f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
      f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
      f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
    }
    T* d = dst + (long long)b * batch_stride + ((long long)y * W + x) * pix_stride + c8 * 8;
    if constexpr (sizeof(T) == 4) {
      if (accumulate) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += d[j];
      }
      *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
      if (accumulate) {
        const uint4 o = *reinterpret_cast<const uint4*>(d);
        float2 f;
        f = unpack_bf16x2(o.x); acc[0] += f.x; acc[1] += f.y;
        f = unpack_bf16x2(o.y); acc[2] += f.x; acc[3] += f.y;
        f = unpack_bf16x2(o.z); acc[4] += f.x; acc[5] += f.y;
        f = unpack_bf16x2(o.w); acc[6] += f.x; acc[7] += f.y;
      }
      uint4 u;
      u.x = pack_bf16x2(acc[0], acc[1]); u.y = pack_bf16x2(acc[2], acc[3]);
      u.z = pack_bf16x2(acc[4], acc[5]); u.w = pack_bf16x2(acc[6], acc[7]);
      *reinterpret_cast<uint4*>(d) = u;
    }
  }
}

// ---- BatchNorm -------------------------------------------------------------------------------------------------
// column sums of x and x^2 (forward statistics), or of dy and dy*xhat (backward), over a bf16 [rows, C] matrix
__global__ void __launch_bounds__(256) bn_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                        const float* __restrict__ mean, const float* __restrict__ invstd,
                                                        long long rows, int C, float* __restrict__ out0,
                                                        float* __restrict__ out1, long long rows_per_block) {
  pdl_prologue();
  __shared__ float sh0[8][64], sh1[8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + tx * 2;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  if (c < C) {
    float m0 = 0.f, m1 = 0.f, i0 = 1.f, i1 = 1.f;
    if (dy) { m0 = mean[c]; m1 = mean[c + 1]; i0 = invstd[c]; i1 = invstd[c + 1]; }
    for (long long r = r0 + ty; r < r1; r += 8) {
      const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * C + c));
      if (dy) {
        const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + r * C + c));
        a0 += d.x; a1 += d.y;
        b0 += d.x * (xv.x - m0) * i0; b1 += d.y * (xv.y - m1) * i1;
      } else {
        a0 += xv.x; a1 += xv.y;
        b0 += xv.x * xv.x; b1 += xv.y * xv.y;
      }
    }
  }
  sh0[ty][tx * 2] = a0; sh0[ty][tx * 2 + 1] = a1;
  sh1[ty][tx * 2] = b0; sh1[ty][tx * 2 + 1] = b1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0 += sh0[j][threadIdx.x]; s1 += sh1[j][threadIdx.x]; }
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < C) { atomicAdd(out0 + cc, s0); atomicAdd(out1 + cc, s1); }
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, long long rows,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, int training, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out, int C) {
  pdl_prologue();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    mean = sum[c] / (float)rows;
    var = fmaxf(sumsq[c] / (float)rows - mean * mean, 0.f);
    if (running_mean) {
      const float unbiased = rows > 1 ? var * (float)rows / (float)(rows - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
}

// out[r, coff + c] = (x[r,c]*scale[c] + shift[c]) * m1[r,c] * m2[r,c]      (m1/m2 optional, bf16)
__global__ void __launch_bounds__(256) bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const __nv_bfloat16* __restrict__ m1,
                                                       int m1_ld, const __nv_bfloat16* __restrict__ m2, int m2_ld,
                                                       __nv_bfloat16* __restrict__ out, int out_ld, int out_coff,
                                                       long long rows, int C) {
  pdl_prologue();
  const int c2n = C / 2;
  const long long total = rows * c2n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c2n) * 2;
    const long long r = (long long)((unsigned int)i / (unsigned int)c2n);
    float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * C + c));
    v.x = v.x * scale[c] + shift[c];
    v.y = v.y * scale[c + 1] + shift[c + 1];
    if (m1) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(m1 + r * m1_ld + c));
      v.x *= a.x; v.y *= a.y;
    }
    if (m2) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(m2 + r * m2_ld + c));
      v.x *= a.x; v.y *= a.y;
    }
    *reinterpret_cast<uint32_t*>(out + r * out_ld + out_coff + c) = pack_bf16x2(v.x, v.y);
  }
}

// training BN backward: dx = scale * (dy - sum_dy/rows - xhat * sum_dy_xhat/rows)    (eval: dx = scale * dy)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                           const float* __restrict__ scale, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const float* __restrict__ sum_dy,
                                                           const float* __restrict__ sum_dy_xhat, __nv_bfloat16* __restrict__ dx,
                                                           long long rows, int C, int training) {
  pdl_prologue();
  const int c2n = C / 2;
  const long long total = rows * c2n;
  const float inv_rows = 1.f / (float)rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c2n) * 2;
    const long long r = (long long)((unsigned int)i / (unsigned int)c2n);
    const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + r * C + c));
    float o0, o1;
    if (training) {
      const float2 xv = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * C + c));
      const float xh0 = (xv.x - mean[c]) * invstd[c], xh1 = (xv.y - mean[c + 1]) * invstd[c + 1];
      o0 = scale[c] * (d.x - sum_dy[c] * inv_rows - xh0 * sum_dy_xhat[c] * inv_rows);
      o1 = scale[c + 1] * (d.y - sum_dy[c + 1] * inv_rows - xh1 * sum_dy_xhat[c + 1] * inv_rows);
    } else {
      o0 = scale[c] * d.x;
      o1 = scale[c + 1] * d.y;
    }
    *reinterpret_cast<uint32_t*>(dx + r * C + c) = pack_bf16x2(o0, o1);
  }
}

// dst[r, dcoff + c] (+)= a[r, acoff + c] * b[r,c] * c2[r,c]    (a: bf16 or fp32; b, c2 optional bf16; dst bf16/fp32)
template <typename TA, typename TD>
__global__ void __launch_bounds__(256) ew_mul_kernel(const TA* __restrict__ a, int a_ld, int a_coff, const __nv_bfloat16* __restrict__ b,
                                                     int b_ld, const __nv_bfloat16* __restrict__ c2, int c2_ld, TD* __restrict__ dst,
                                                     int d_ld, int d_coff, long long rows, int C, int accumulate) {
  pdl_prologue();
  const int c2n = C / 2;
  const long long total = rows * c2n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c2n) * 2;
    const long long r = (long long)((unsigned int)i / (unsigned int)c2n);
    float2 v;
    if constexpr (sizeof(TA) == 2) v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(a + r * a_ld + a_coff + c));
    else v = *reinterpret_cast<const float2*>(a + r * a_ld + a_coff + c);
    if (b) {
      const float2 t = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(b + r * b_ld + c));
      v.x *= t.x; v.y *= t.y;
    }
    if (c2) {
      const float2 t = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(c2 + r * c2_ld + c));
      v.x *= t.x; v.y *= t.y;
    }
    TD* d = dst + r * d_ld + d_coff + c;
    if constexpr (sizeof(TD) == 2) {
      if (accumulate) {
        const float2 o = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(d));
        v.x += o.x; v.y += o.y;
      }
      *reinterpret_cast<uint32_t*>(d) = pack_bf16x2(v.x, v.y);
    } else {
      if (accumulate) { v.x += d[0]; v.y += d[1]; }
      *reinterpret_cast<float2*>(d) = v;
    }
  }
}

// ---- bilinear x2 upsample, align_corners=True, NHWC ------------------------------------------------------------
__device__ __forceinline__ void ac_index(int dst, int in_size, int out_size, int& i0, int& i1, float& l1) {
  // torch area_pixel_compute_scale(align_corners=True): scale = (in-1)/(out-1) in fp32, src = scale * dst
  const float s = (out_size > 1) ? ((float)(in_size - 1) / (float)(out_size - 1)) * (float)dst : 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

__device__ __forceinline__ void load8f(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8f(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8f(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// 8 channels (one 16-byte vector) per thread
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const T* __restrict__ src, long long batch_stride, int pix_stride,
                                                             __nv_bfloat16* __restrict__ dst, int B, int h, int w, int C) {
  pdl_prologue();
  const int H = 2 * h, W = 2 * w, c8n = C / 8;
  const long long total = (long long)B * H * W * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c8n) * 8;
    unsigned int t = (unsigned int)i / (unsigned int)c8n;
    const int X = (int)(t % W); t /= W;
    const int Y = (int)(t % H);
    const int b = (int)(t / H);
    int y0, y1, x0, x1;
    float ly, lx;
    ac_index(Y, h, H, y0, y1, ly);
    ac_index(X, w, W, x0, x1, lx);
    const T* sb = src + (long long)b * batch_stride + c;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    load8f(sb + ((long long)y0 * w + x0) * pix_stride, v00);
    load8f(sb + ((long long)y0 * w + x1) * pix_stride, v01);
    load8f(sb + ((long long)y1 * w + x0) * pix_stride, v10);
    load8f(sb + ((long long)y1 * w + x1) * pix_stride, v11);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = w00 * v00[j] + w01 * v01[j] + w10 * v10[j] + w11 * v11[j];
    store8f(dst + (((long long)b * H + Y) * W + X) * C + c, o);
  }
}

// gather form of the transposed operator: dx[b,i,j,c] (+)= sum over the <= 5x5 output pixels that read (i,j);
// 8 channels per thread, the (tiny) weight search is amortised over them
template <typename TD>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, TD* __restrict__ dx,
                                                             long long batch_stride, int pix_stride, int B, int h, int w, int C,
                                                             int accumulate) {
  pdl_prologue();
  const int H = 2 * h, W = 2 * w, c8n = C / 8;
  const long long total = (long long)B * h * w * c8n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)idx % (unsigned int)c8n) * 8;
    unsigned int t = (unsigned int)idx / (unsigned int)c8n;
    const int j = (int)(t % w); t /= w;
    const int i = (int)(t % h);
    const int b = (int)(t / h);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // per-axis weights of the <= 7 candidate output coordinates
    float wys[7], wxs[7];
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const int Y = 2 * i - 3 + u, X = 2 * j - 3 + u;
      int a0, a1;
      float l;
      wys[u] = 0.f;
      wxs[u] = 0.f;
      if (Y >= 0 && Y < H) {
        ac_index(Y, h, H, a0, a1, l);
        if (a0 == i) wys[u] += 1.f - l;
        if (a1 == i) wys[u] += l;
      }
      if (X >= 0 && X < W) {
        ac_index(X, w, W, a0, a1, l);
        if (a0 == j) wxs[u] += 1.f - l;
        if (a1 == j) wxs[u] += l;
      }
    }
#pragma unroll
    for (int uy = 0; uy < 7; ++uy) {
      if (wys[uy] == 0.f) continue;
      const int Y = 2 * i - 3 + uy;
#pragma unroll
      for (int ux = 0; ux < 7; ++ux) {
        if (wxs[ux] == 0.f) continue;
        const int X = 2 * j - 3 + ux;
        float g[8];
        load8f(dy + (((long long)b * H + Y) * W + X) * C + c, g);
        const float wgt = wys[uy] * wxs[ux];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += wgt * g[q];
      }
    }
    TD* d = dx + (long long)b * batch_stride + ((long long)i * w + j) * pix_stride + c;
    if (accumulate) {
      float o[8];
      load8f(d, o);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += o[q];
    }
    store8f(d, acc);
  }
}

// ---- 1x1 score conv (Cin -> 3) -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) score_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ W,
                                                        const float* __restrict__ bias, float* __restrict__ out, long long rows,
                                                        int Cin) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long long wid = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= rows) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int c = lane * 2; c < Cin; c += 64) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + wid * Cin + c));
    a0 += v.x * W[c] + v.y * W[c + 1];
    a1 += v.x * W[Cin + c] + v.y * W[Cin + c + 1];
    a2 += v.x * W[2 * Cin + c] + v.y * W[2 * Cin + c + 1];
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
  if (lane == 0) {
    out[wid * 3 + 0] = a0 + bias[0];
    out[wid * 3 + 1] = a1 + bias[1];
    out[wid * 3 + 2] = a2 + bias[2];
  }
}
// dx[r,c] = sum_j ds[r,j] W[j,c] (bf16);  dW[j,c] += sum_r ds[r,j] x[r,c];  db[j] += sum_r ds[r,j]
__global__ void __launch_bounds__(256) score_bwd_kernel(const float* __restrict__ ds, const __nv_bfloat16* __restrict__ x,
                                                        const float* __restrict__ W, __nv_bfloat16* __restrict__ dx,
                                                        float* __restrict__ dW, float* __restrict__ db, long long rows, int Cin,
                                                        long long rows_per_block, const float* __restrict__ gscale) {
  pdl_prologue();
  const float gsc = gscale ? *gscale : 1.f;   // upstream gradient of the total loss (device scalar), folded in here
  extern __shared__ float shw[];  // [3*Cin] partial dW + [3] db
  for (int i = threadIdx.x; i < 3 * Cin + 3; i += blockDim.x) shw[i] = 0.f;
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  // each thread owns one channel (Cin <= blockDim.x) and walks the block's rows
  const int c = threadIdx.x;
  if (c < Cin) {
    float w0 = W[c], w1 = W[Cin + c], w2 = W[2 * Cin + c];
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
      const float d0 = ds[r * 3] * gsc, d1 = ds[r * 3 + 1] * gsc, d2 = ds[r * 3 + 2] * gsc;
      const float xv = __bfloat162float(x[r * Cin + c]);
      dx[r * Cin + c] = __float2bfloat16(d0 * w0 + d1 * w1 + d2 * w2);
      g0 += d0 * xv; g1 += d1 * xv; g2 += d2 * xv;
    }
    atomicAdd(dW + c, g0);
    atomicAdd(dW + Cin + c, g1);
    atomicAdd(dW + 2 * Cin + c, g2);
  } else if (c < Cin + 3) {
    const int j = c - Cin;
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += ds[r * 3 + j];
    atomicAdd(db + j, s * gsc);
  }
}

// ---- x8 bilinear upsample (align_corners=True) of the NHWC score map to NCHW, and SmoothL1 fused with it ----------
__global__ void __launch_bounds__(256) upsample8_fwd_kernel(const float* __restrict__ score, float* __restrict__ out, int B, int h,
                                                            int w, int S) {
  pdl_prologue();
  const int H = h * S, W = w * S;
  const long long total = (long long)B * 3 * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)((unsigned int)i % (unsigned int)W);
    unsigned int t = (unsigned int)i / (unsigned int)W;
    const int Y = (int)(t % H); t /= H;
    const int c = (int)(t % 3);
    const int b = (int)(t / 3);
    int y0, y1, x0, x1;
    float ly, lx;
    ac_index(Y, h, H, y0, y1, ly);
    ac_index(X, w, W, x0, x1, lx);
    const float* sb = score + (long long)b * h * w * 3 + c;
    const float v00 = sb[(y0 * w + x0) * 3], v01 = sb[(y0 * w + x1) * 3], v10 = sb[(y1 * w + x0) * 3], v11 = sb[(y1 * w + x1) * 3];
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

// One block per (b, c, output-row band of S rows). mode 0: g = dpred (given); mode 1: g = d SmoothL1(pred - target) *
// scale * gscale and the loss itself is accumulated. dscore (fp32 NHWC [B,h,w,3]) receives the transposed interpolation
// of g. Everything is separable and lanes always walk the output columns X (coalesced global reads, conflict-free
// shared memory):
//   setup   : per-column source index / weight tables, the first output column of every source column (xs), the weights
//             of the band's S rows onto the <= 3 score rows it touches (wy), and -- mode 1 -- those 3 score rows
//             interpolated along X once per block (ri), so that a prediction is 2 shared loads + 2 FMAs;
//   phase 1 : g[S][W] -> shared (+ loss);
//   phase 2 : along Y first: gy[r][X] = sum_Y wy[r][Y] g[Y][X] for the 3 score rows, stored pre-multiplied by the two
//             X weights (A = (1 - lx) gy, Bv = lx gy) at padded positions X + X/8 (the X phase reads with stride ~8.2);
//   phase 3 : along X: dscore[r][x] += sum_{X in bin(x)} A + sum_{X in bin(x-1)} Bv, one global atomic per touched cell.
__global__ void __launch_bounds__(256) t2i_up8_loss_kernel(const float* __restrict__ score, const float* __restrict__ target,
                                                           const float* __restrict__ dpred, float* __restrict__ dscore,
                                                           float* __restrict__ loss_sum, float* __restrict__ total_sum,
                                                           float loss_scale, float grad_scale, const float* __restrict__ gscale,
                                                           int B, int h, int w, int S, int mode, int want_grad) {
  pdl_prologue();
  extern __shared__ float sh[];
  const int H = h * S, W = w * S;
  const int WP = W + (W >> 3) + 1;   // padded row length of A / Bv
  const int bands = H / S;  // one band = S output rows => touches at most score rows ybase .. ybase+2
  const int band = blockIdx.x % bands;
  const int c = (blockIdx.x / bands) % 3;
  const int b = blockIdx.x / (bands * 3);
  float* gs = sh;               // [S][W]
  float* red = gs + S * W;      // [32]
  float* lxs = red + 32;        // [W]  interpolation weight of output column X
  float* lys = lxs + W;         // [S]
  float* wys = lys + S;         // [3][S] weight of band row yy onto score row ybase + r
  float* ri = wys + 3 * S;      // [3][W] score rows ybase .. ybase+2 interpolated along X (mode 1)
  float* As = ri + 3 * W;       // [3][WP]
  float* Bs = As + 3 * WP;      // [3][WP]
  int* x0s = reinterpret_cast<int*>(Bs + 3 * WP);   // [W]  left source column of output column X
  int* y0s = x0s + W;                               // [S]
  int* xs = y0s + S;                                // [w + 1] first output column whose left source column is x
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    int x0, x1;
    float lx;
    ac_index(X, w, W, x0, x1, lx);
    x0s[X] = x0;
    lxs[X] = lx;
  }
  for (int yy = threadIdx.x; yy < S; yy += blockDim.x) {
    int y0, y1;
    float ly;
    ac_index(band * S + yy, h, H, y0, y1, ly);
    y0s[yy] = y0;
    lys[yy] = ly;
  }
  for (int x = threadIdx.x; x <= w; x += blockDim.x) xs[x] = W;
  __syncthreads();
  const int ybase = y0s[0];
  for (int X = threadIdx.x; X < W; X += blockDim.x)
    if (X == 0 || x0s[X] != x0s[X - 1]) xs[x0s[X]] = X;     // x0 is non-decreasing in X and skips no source column
  for (int e = threadIdx.x; e < 3 * S; e += blockDim.x) {
    const int r = e / S, yy = e - r * S;
    const int y0 = y0s[yy];
    const int y1 = y0 + ((y0 < h - 1) ? 1 : 0);
    const float ly = lys[yy];
    float wy = 0.f;
    if (y0 == ybase + r) wy += 1.f - ly;
    if (y1 == ybase + r) wy += ly;
    wys[e] = wy;
  }
  const float* sb = score + (long long)b * h * w * 3 + c;
  if (mode == 1) {
    for (int r = 0; r < 3; ++r) {
      const int y = min(ybase + r, h - 1);
      for (int X = threadIdx.x; X < W; X += blockDim.x) {
        const int x0 = x0s[X];
        const int x1 = x0 + ((x0 < w - 1) ? 1 : 0);
        const float lx = lxs[X];
        ri[r * W + X] = (1.f - lx) * sb[(y * w + x0) * 3] + lx * sb[(y * w + x1) * 3];
      }
    }
  }
  __syncthreads();
  // phase 1
  const float g_mul = grad_scale * (gscale ? *gscale : 1.f);
  float lsum = 0.f;
  for (int yy = 0; yy < S; ++yy) {
    const int y0 = y0s[yy];
    const int r0 = y0 - ybase;
    const int r1 = min(r0 + ((y0 < h - 1) ? 1 : 0), 2);
    const float ly = lys[yy];
    const long long rowoff = (((long long)b * 3 + c) * H + (band * S + yy)) * W;
    for (int X = threadIdx.x; X < W; X += blockDim.x) {
      float g;
      if (mode == 1) {
        const float pred = (1.f - ly) * ri[r0 * W + X] + ly * ri[r1 * W + X];
        const float d = pred - target[rowoff + X];
        const float ad = fabsf(d);
        lsum += (ad < 1.f) ? 0.5f * d * d : ad - 0.5f;
        g = fminf(fmaxf(d, -1.f), 1.f) * g_mul;
      } else {
        g = dpred[rowoff + X];
      }
      gs[yy * W + X] = g;
    }
  }
  if (mode == 1 && loss_sum != nullptr) {
    lsum = block_sum(lsum, red);
    if (threadIdx.x == 0) {
      atomicAdd(loss_sum, lsum * loss_scale);
      if (total_sum) atomicAdd(total_sum, lsum * loss_scale);
    }
  }
  if (!want_grad) return;
  __syncthreads();
  // phase 2: along Y (each thread re-reads the column it wrote), pre-multiplied by the X weights
  for (int X = threadIdx.x; X < W; X += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int yy = 0; yy < S; ++yy) {
      const float g = gs[yy * W + X];
      a0 = fmaf(wys[yy], g, a0);
      a1 = fmaf(wys[S + yy], g, a1);
      a2 = fmaf(wys[2 * S + yy], g, a2);
    }
    const float lx = lxs[X];
    const float wl = (x0s[X] == w - 1) ? 1.f : 1.f - lx;   // the right neighbour of the last source column is itself
    const float wr = (x0s[X] == w - 1) ? 0.f : lx;
    const int Xp = X + (X >> 3);
    As[Xp] = wl * a0; As[WP + Xp] = wl * a1; As[2 * WP + Xp] = wl * a2;
    Bs[Xp] = wr * a0; Bs[WP + Xp] = wr * a1; Bs[2 * WP + Xp] = wr * a2;
  }
  __syncthreads();
  // phase 3: along X, then the global accumulation (score rows ybase .. ybase+2)
  for (int e = threadIdx.x; e < 3 * w; e += blockDim.x) {
    const int r = e / w, x = e - r * w;
    const int y = ybase + r;
    if (y >= h) continue;
    const int Xa = xs[x], Xb = xs[x + 1];
    float a = 0.f;
    for (int X = Xa; X < Xb; ++X) a += As[r * WP + X + (X >> 3)];
    if (x > 0)
      for (int X = xs[x - 1]; X < Xa; ++X) a += Bs[r * WP + X + (X >> 3)];
    if (a != 0.f) atomicAdd(dscore + ((long long)b * h * w + (long long)y * w + x) * 3 + c, a);
  }
}

}  // namespace

extern "C" int mvlt_im2col3x3(const void* src, int src_f32, long long batch_stride, int pix_stride, void* col, int B, int H,
                              int W, int C, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0 && pix_stride % 8 == 0, "im2col3x3: C must be a multiple of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)B * H * W * 9 * (C / 8);
  if (src_f32)
    mvlt_launch(im2col3x3_kernel<float>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const float*>(src), batch_stride, pix_stride,
                                                                  reinterpret_cast<__nv_bfloat16*>(col), B, H, W, C);
  else
    mvlt_launch(im2col3x3_kernel<__nv_bfloat16>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(src), batch_stride,
                                                                          pix_stride, reinterpret_cast<__nv_bfloat16*>(col), B, H, W, C);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_col2im3x3(const void* dcol, void* dst, int dst_f32, long long batch_stride, int pix_stride, int B, int H,
                              int W, int C, int accumulate, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0 && pix_stride % 8 == 0, "col2im3x3: C must be a multiple of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)B * H * W * (C / 8);
  if (dst_f32)
    mvlt_launch(col2im3x3_kernel<float>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dcol),
                                                                  reinterpret_cast<float*>(dst), batch_stride, pix_stride, B, H, W, C, accumulate);
  else
    mvlt_launch(col2im3x3_kernel<__nv_bfloat16>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dcol),
                                                                          reinterpret_cast<__nv_bfloat16*>(dst), batch_stride, pix_stride, B, H, W, C, accumulate);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// 8 channels (one 16-byte vector) per thread; requires C, leading dimensions and column offsets to be multiples of 8
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void load8p(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__global__ void __launch_bounds__(256) bn_apply8_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const __nv_bfloat16* __restrict__ m1,
                                                        int m1_ld, const __nv_bfloat16* __restrict__ m2, int m2_ld,
                                                        __nv_bfloat16* __restrict__ out, int out_ld, int out_coff,
                                                        long long rows, int C) {
  pdl_prologue();
  const unsigned int c8n = (unsigned int)(C / 8);
  const long long total = rows * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % c8n) * 8;
    const long long r = (long long)((unsigned int)i / c8n);
    float v[8], sc[8], sh[8];
    unpack8(*reinterpret_cast<const uint4*>(x + r * C + c), v);
    load8p(scale + c, sc);
    load8p(shift + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (m1) {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(m1 + r * m1_ld + c), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= a[j];
    }
    if (m2) {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(m2 + r * m2_ld + c), a);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= a[j];
    }
    *reinterpret_cast<uint4*>(out + r * out_ld + out_coff + c) = pack8(v);
  }
}
__global__ void __launch_bounds__(256) bn_bwd_apply8_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ scale, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ sum_dy,
                                                            const float* __restrict__ sum_dy_xhat, __nv_bfloat16* __restrict__ dx,
                                                            long long rows, int C, int training) {
  pdl_prologue();
  const unsigned int c8n = (unsigned int)(C / 8);
  const long long total = rows * c8n;
  const float inv_rows = 1.f / (float)rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % c8n) * 8;
    const long long r = (long long)((unsigned int)i / c8n);
    float d[8], sc[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(dy + r * C + c), d);
    load8p(scale + c, sc);
    if (training) {
      float xv[8], mu[8], is[8], s1[8], s2[8];
      unpack8(*reinterpret_cast<const uint4*>(x + r * C + c), xv);
      load8p(mean + c, mu);
      load8p(invstd + c, is);
      load8p(sum_dy + c, s1);
      load8p(sum_dy_xhat + c, s2);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mu[j]) * is[j];
        o[j] = sc[j] * (d[j] - s1[j] * inv_rows - xh * s2[j] * inv_rows);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = sc[j] * d[j];
    }
    *reinterpret_cast<uint4*>(dx + r * C + c) = pack8(o);
  }
}

static void bn_reduce_launch(const void* x, const void* dy, const float* mean, const float* invstd, long long rows, int C,
                             float* o0, float* o1, cudaStream_t st) {
  const int gx = (C + 63) / 64;
  long long gy = (long long)mvlt_num_sms() * 4 / gx;
  if (gy < 1) gy = 1;
  long long rpb = (rows + gy - 1) / gy;
  if (rpb < 64) rpb = 64;
  gy = (rows + rpb - 1) / rpb;
  mvlt_launch(bn_reduce_kernel, dim3(gx, (unsigned)gy), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(x),
                                                           reinterpret_cast<const __nv_bfloat16*>(dy), mean, invstd, rows, C, o0, o1, rpb);
}

// forward statistics: sum / sumsq must be zeroed by the caller
extern "C" int mvlt_bn_stats(const void* x_bf16, long long rows, int C, float* sum, float* sumsq, void* stream_) {
  MVLT_CHECK_ARG(C % 2 == 0, "bn_stats: C must be even");
  bn_reduce_launch(x_bf16, nullptr, nullptr, nullptr, rows, C, sum, sumsq, reinterpret_cast<cudaStream_t>(stream_));
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_bn_finalize(const float* sum, const float* sumsq, long long rows, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float momentum, float eps, int training, float* scale,
                                float* shift, float* mean_out, float* invstd_out, int C, void* stream_) {
  mvlt_launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream_), sum, sumsq, rows, gamma, beta, running_mean, running_var, momentum, eps, training, scale, shift, mean_out, invstd_out, C);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_bn_apply(const void* x_bf16, const float* scale, const float* shift, const void* m1, int m1_ld, const void* m2,
                             int m2_ld, void* out_bf16, int out_ld, int out_coff, long long rows, int C, void* stream_) {
  MVLT_CHECK_ARG(C % 2 == 0 && out_ld % 2 == 0 && out_coff % 2 == 0, "bn_apply: even sizes required");
  const bool vec8 = (C % 8 == 0) && (out_ld % 8 == 0) && (out_coff % 8 == 0) && (m1 == nullptr || m1_ld % 8 == 0) &&
                    (m2 == nullptr || m2_ld % 8 == 0) &&
                    ((((uintptr_t)x_bf16 | (uintptr_t)out_bf16 | (uintptr_t)m1 | (uintptr_t)m2 | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0);
  if (vec8)
    mvlt_launch(bn_apply8_kernel, cap_grid(rows * (C / 8), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(x_bf16), scale, shift, reinterpret_cast<const __nv_bfloat16*>(m1), m1_ld,
        reinterpret_cast<const __nv_bfloat16*>(m2), m2_ld, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_ld, out_coff, rows, C);
  else
  mvlt_launch(bn_apply_kernel, cap_grid(rows * (C / 2), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(x_bf16), scale, shift, reinterpret_cast<const __nv_bfloat16*>(m1), m1_ld,
      reinterpret_cast<const __nv_bfloat16*>(m2), m2_ld, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_ld, out_coff, rows, C);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// backward: sum_dy / sum_dy_xhat (zeroed by the caller) are ALSO the dbeta / dgamma of the affine parameters
extern "C" int mvlt_bn_bwd(const void* dy_bf16, const void* x_bf16, const float* scale, const float* mean, const float* invstd,
                           float* sum_dy, float* sum_dy_xhat, void* dx_bf16, long long rows, int C, int training, void* stream_) {
  MVLT_CHECK_ARG(C % 2 == 0, "bn_bwd: C must be even");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  bn_reduce_launch(x_bf16, dy_bf16, mean, invstd, rows, C, sum_dy, sum_dy_xhat, st);
  const bool vec8 = (C % 8 == 0) && ((((uintptr_t)dy_bf16 | (uintptr_t)x_bf16 | (uintptr_t)dx_bf16 | (uintptr_t)scale | (uintptr_t)mean |
                                        (uintptr_t)invstd | (uintptr_t)sum_dy | (uintptr_t)sum_dy_xhat) & 15) == 0);
  if (vec8)
    mvlt_launch(bn_bwd_apply8_kernel, cap_grid(rows * (C / 8), 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), reinterpret_cast<const __nv_bfloat16*>(x_bf16), scale, mean, invstd,
        sum_dy, sum_dy_xhat, reinterpret_cast<__nv_bfloat16*>(dx_bf16), rows, C, training);
  else
  mvlt_launch(bn_bwd_apply_kernel, cap_grid(rows * (C / 2), 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dy_bf16), reinterpret_cast<const __nv_bfloat16*>(x_bf16), scale, mean, invstd,
      sum_dy, sum_dy_xhat, reinterpret_cast<__nv_bfloat16*>(dx_bf16), rows, C, training);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_ew_mul(const void* a, int a_f32, int a_ld, int a_coff, const void* b, int b_ld, const void* c2, int c2_ld,
                           void* dst, int dst_f32, int d_ld, int d_coff, long long rows, int C, int accumulate, void* stream_) {
  MVLT_CHECK_ARG(C % 2 == 0, "ew_mul: C must be even");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = cap_grid(rows * (C / 2), 256);
  const __nv_bfloat16* bb = reinterpret_cast<const __nv_bfloat16*>(b);
  const __nv_bfloat16* cc = reinterpret_cast<const __nv_bfloat16*>(c2);
#define LAUNCH(TA, TD)                                                                                                   \
  mvlt_launch(ew_mul_kernel<TA, TD>, grid, 256, 0, st, reinterpret_cast<const TA*>(a), a_ld, a_coff, bb, b_ld, cc, c2_ld,          \
                                              reinterpret_cast<TD*>(dst), d_ld, d_coff, rows, C, accumulate)
  if (a_f32 && dst_f32) LAUNCH(float, float);
  else if (a_f32) LAUNCH(float, __nv_bfloat16);
  else if (dst_f32) LAUNCH(__nv_bfloat16, float);
  else LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_upsample2x_fwd(const void* src, int src_f32, long long batch_stride, int pix_stride, void* dst_bf16, int B,
                                   int h, int w, int C, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0 && pix_stride % 8 == 0 && batch_stride % 8 == 0, "upsample2x_fwd: C / strides must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)B * 4 * h * w * (C / 8);
  if (src_f32)
    mvlt_launch(upsample2x_fwd_kernel<float>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const float*>(src), batch_stride, pix_stride,
                                                                       reinterpret_cast<__nv_bfloat16*>(dst_bf16), B, h, w, C);
  else
    mvlt_launch(upsample2x_fwd_kernel<__nv_bfloat16>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(src), batch_stride,
                                                                               pix_stride, reinterpret_cast<__nv_bfloat16*>(dst_bf16), B, h, w, C);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_upsample2x_bwd(const void* dy_bf16, void* dx, int dx_f32, long long batch_stride, int pix_stride, int B, int h,
                                   int w, int C, int accumulate, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0 && pix_stride % 8 == 0 && batch_stride % 8 == 0, "upsample2x_bwd: C / strides must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)B * h * w * (C / 8);
  if (dx_f32)
    mvlt_launch(upsample2x_bwd_kernel<float>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dy_bf16),
                                                                       reinterpret_cast<float*>(dx), batch_stride, pix_stride, B, h, w, C, accumulate);
  else
    mvlt_launch(upsample2x_bwd_kernel<__nv_bfloat16>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(dy_bf16),
                                                                               reinterpret_cast<__nv_bfloat16*>(dx), batch_stride, pix_stride, B, h, w, C, accumulate);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_score_fwd(const void* x_bf16, const float* W, const float* bias, float* out, long long rows, int Cin, void* stream_) {
  MVLT_CHECK_ARG(Cin % 2 == 0, "score_fwd: Cin must be even");
  mvlt_launch(score_fwd_kernel, (int)((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(x_bf16), W, bias, out, rows, Cin);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// gscale_dev: optional device scalar multiplied into dscore (the upstream gradient of the total loss)
extern "C" int mvlt_score_bwd(const float* dscore, const void* x_bf16, const float* W, void* dx_bf16, float* dW, float* db,
                              long long rows, int Cin, const float* gscale_dev, void* stream_) {
  MVLT_CHECK_ARG(Cin + 3 <= 256, "score_bwd: Cin must be <= 253");
  long long blocks = (long long)mvlt_num_sms() * 8;   // one full wave of 256-thread blocks
  long long rpb = (rows + blocks - 1) / blocks;
  if (rpb < 16) rpb = 16;
  blocks = (rows + rpb - 1) / rpb;
  mvlt_launch(score_bwd_kernel, (int)blocks, 256, (3 * Cin + 3) * sizeof(float), reinterpret_cast<cudaStream_t>(stream_), dscore, reinterpret_cast<const __nv_bfloat16*>(x_bf16), W, reinterpret_cast<__nv_bfloat16*>(dx_bf16), dW, db, rows, Cin, rpb, gscale_dev);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_upsample8_fwd(const float* score, float* out, int B, int h, int w, int S, void* stream_) {
  const long long total = (long long)B * 3 * h * S * w * S;
  mvlt_launch(upsample8_fwd_kernel, cap_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), score, out, B, h, w, S);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// mode 1: SmoothL1(pred, target) loss (+ optional gradient wrt the score map); mode 0: backward of the plain upsample
extern "C" int mvlt_t2i_up_loss(const float* score, const float* target, const float* dpred, float* dscore, float* loss_sum,
                                float* total_sum, float loss_scale, float grad_scale, const float* gscale_dev, int B, int h, int w,
                                int S, int mode, int want_grad, void* stream_) {
  const int W_ = w * S, WP_ = W_ + (W_ >> 3) + 1;
  const size_t smem = (size_t)(S * W_ + 32 + W_ + S + 3 * S + 3 * W_ + 6 * WP_ + W_ + S + (w + 1)) * sizeof(float);
  MVLT_CHECK_ARG(S >= 2 && smem <= 48 * 1024, "t2i_up_loss: bad geometry");
  const int blocks = B * 3 * h;
  mvlt_launch(t2i_up8_loss_kernel, blocks, 256, smem, reinterpret_cast<cudaStream_t>(stream_), score, target, dpred, dscore, loss_sum, total_sum, loss_scale, grad_scale, gscale_dev, B, h, w, S, mode, want_grad);
  MVLT_CHECK_LAUNCH();
  return 0;
}
