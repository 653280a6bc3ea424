// Memory-bound glue kernels: casts, column sums (bias grads), patchify/unpatchify for the k=s convolutions,
// strided row copies, batch reductions and the bilinear position-embedding resize (fwd + bwd).
// All are coalesced and 128-bit vectorised where the layout allows; grids are capped at a multiple of the SM count.
//
// Reference ops replaced: the permute/reshape/contiguous copies around /root/reference/libs/pvlt.py:102-104,168,
// 350-352, F.interpolate at :295-297, torch.cat/split at :107,:346 and autograd's bias/broadcast reductions.
#include <string.h>
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

// ---- dst_bf16[i] = bf16(src_f32[i] * scale[row / rows_per_scale]) -------------------------------------------
__global__ void __launch_bounds__(256) cast_scale_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                         long long n8, int C, const float* __restrict__ rowscale,
                                                         int rows_per_scale, float alpha) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 8;
    float s = alpha;
    if (rowscale) s *= rowscale[(e / C) / rows_per_scale];
    const float4 a = *reinterpret_cast<const float4*>(src + e);
    const float4 b = *reinterpret_cast<const float4*>(src + e + 4);
    uint4 u;
    u.x = pack_bf16x2(a.x * s, a.y * s); u.y = pack_bf16x2(a.z * s, a.w * s);
    u.z = pack_bf16x2(b.x * s, b.y * s); u.w = pack_bf16x2(b.z * s, b.w * s);
    *reinterpret_cast<uint4*>(dst + e) = u;
  }
}

// ---- out[c] += sum_r x[r*ld + c]: each thread owns 8 consecutive columns (one 16 B load for bf16) ---------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long rows, int C, long long ld,
                                                     float* __restrict__ out, long long rows_per_block, int tx_n) {
  pdl_prologue();
  __shared__ float sh[256 * 8];
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const int ry = blockDim.x / tx_n;
  const int c = (blockIdx.x * tx_n + tx) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
#pragma unroll 4
    for (long long r = r0 + ty; r < r1; r += ry) {
      if constexpr (sizeof(T) == 2) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + r * ld + c);
        float2 f;
        f = unpack_bf16x2(u.x); a[0] += f.x; a[1] += f.y;
        f = unpack_bf16x2(u.y); a[2] += f.x; a[3] += f.y;
        f = unpack_bf16x2(u.z); a[4] += f.x; a[5] += f.y;
        f = unpack_bf16x2(u.w); a[6] += f.x; a[7] += f.y;
      } else {
        const float4 f0 = *reinterpret_cast<const float4*>(x + r * ld + c);
        const float4 f1 = *reinterpret_cast<const float4*>(x + r * ld + c + 4);
        a[0] += f0.x; a[1] += f0.y; a[2] += f0.z; a[3] += f0.w;
        a[4] += f1.x; a[5] += f1.y; a[6] += f1.z; a[7] += f1.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[(ty * tx_n + tx) * 8 + j] = a[j];
  __syncthreads();
  for (int i = threadIdx.x; i < tx_n * 8; i += blockDim.x) {
    float s = 0.f;
    for (int y = 0; y < ry; ++y) s += sh[(y * tx_n) * 8 + i];
    const int cc = blockIdx.x * tx_n * 8 + i;
    if (cc < C) atomicAdd(out + cc, s);
  }
}

// ---- patchify: NHWC tokens -> [B*oh*ow, R*R*C] bf16 (k index = (ky*R+kx)*C + c) ------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* __restrict__ src, long long src_batch_stride,
                                                       __nv_bfloat16* __restrict__ dst, int B, int H, int W, int C,
                                                       int R) {
  pdl_prologue();
  const int c8n = C / 8;
  const long long total = (long long)B * H * W * c8n;
  const int ow = W / R, oh = H / R;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)((unsigned int)i % (unsigned int)c8n);
    unsigned int t = (unsigned int)i / (unsigned int)c8n;
    const int xw = (int)(t % W); t /= W;
    const int yh = (int)(t % H);
    const int b = (int)(t / H);
    const T* s = src + (long long)b * src_batch_stride + ((long long)yh * W + xw) * C + c8 * 8;
    uint4 u;
    if constexpr (sizeof(T) == 2) {
      u = *reinterpret_cast<const uint4*>(s);
    } else {
      const float4 a = *reinterpret_cast<const float4*>(s);
      const float4 bb = *reinterpret_cast<const float4*>(s + 4);
      u.x = pack_bf16x2(a.x, a.y); u.y = pack_bf16x2(a.z, a.w);
      u.z = pack_bf16x2(bb.x, bb.y); u.w = pack_bf16x2(bb.z, bb.w);
    }
    const int oy = yh / R, ky = yh % R, ox = xw / R, kx = xw % R;
    const long long drow = ((long long)b * oh + oy) * ow + ox;
    *reinterpret_cast<uint4*>(dst + drow * ((long long)R * R * C) + (long long)(ky * R + kx) * C + c8 * 8) = u;
  }
}

// inverse: dpatch bf16 [B*oh*ow, R*R*C] -> dst fp32 NHWC rows (writes; accumulate != 0: adds to what is there)
__global__ void __launch_bounds__(256) unpatchify_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                                                         long long dst_batch_stride, int B, int H, int W, int C, int R,
                                                         int accumulate) {
  pdl_prologue();
  const int c8n = C / 8;
  const long long total = (long long)B * H * W * c8n;
  const int ow = W / R, oh = H / R;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)((unsigned int)i % (unsigned int)c8n);
    unsigned int t = (unsigned int)i / (unsigned int)c8n;
    const int xw = (int)(t % W); t /= W;
    const int yh = (int)(t % H);
    const int b = (int)(t / H);
    const int oy = yh / R, ky = yh % R, ox = xw / R, kx = xw % R;
    const long long srow = ((long long)b * oh + oy) * ow + ox;
    const uint4 u = *reinterpret_cast<const uint4*>(src + srow * ((long long)R * R * C) + (long long)(ky * R + kx) * C + c8 * 8);
    float* d = dst + (long long)b * dst_batch_stride + ((long long)yh * W + xw) * C + c8 * 8;
    float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
    float4 o0 = make_float4(f0.x, f0.y, f1.x, f1.y), o1 = make_float4(f2.x, f2.y, f3.x, f3.y);
    if (accumulate) {
      const float4 a0 = *reinterpret_cast<const float4*>(d), a1 = *reinterpret_cast<const float4*>(d + 4);
      o0.x += a0.x; o0.y += a0.y; o0.z += a0.z; o0.w += a0.w;
      o1.x += a1.x; o1.y += a1.y; o1.z += a1.z; o1.w += a1.w;
    }
    *reinterpret_cast<float4*>(d) = o0;
    *reinterpret_cast<float4*>(d + 4) = o1;
  }
}

// stage-1 patch embed: NCHW fp32 image -> bf16 [B*(H/P)*(W/P), Kpad], k = (ci*P + ky)*P + kx (the conv weight's own
// flattening), zero padded to Kpad
__global__ void __launch_bounds__(256) patchify_nchw_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ dst,
                                                            int B, int Cin, int H, int W, int P, int Kpad) {
  pdl_prologue();
  // One block per (b, oy) row of patches: the Cin*P image rows it needs are read as whole rows (coalesced), transposed
  // through shared memory into [ow][Kpad] bf16 (zero padded) and written out as one contiguous ow*Kpad*2-byte run.
  extern __shared__ __align__(16) unsigned char patch_sh[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(patch_sh);
  const int ow = W / P, oh = H / P;
  const int K = Cin * P * P;
  for (int blk = blockIdx.x; blk < B * oh; blk += gridDim.x) {
    const int oy = blk % oh, b = blk / oh;
    __syncthreads();
    if (Kpad > K)
      for (int i = threadIdx.x; i < ow * (Kpad - K); i += blockDim.x) tile[(i / (Kpad - K)) * Kpad + K + i % (Kpad - K)] = __float2bfloat16(0.f);
    for (int i = threadIdx.x; i < Cin * P * W; i += blockDim.x) {
      const int x = i % W, r = i / W;            // r = ci*P + ky
      const int ci = r / P, ky = r % P;
      const float v = img[(((long long)b * Cin + ci) * H + (oy * P + ky)) * W + x];
      tile[(x / P) * Kpad + r * P + x % P] = __float2bfloat16(v);
    }
    __syncthreads();
    uint4* out = reinterpret_cast<uint4*>(dst + (long long)blk * ow * Kpad);
    const uint4* tin = reinterpret_cast<const uint4*>(tile);
    for (int i = threadIdx.x; i < ow * Kpad / 8; i += blockDim.x) out[i] = tin[i];
  }
}

// Fast path of the above for the geometry PVLT uses (P = 4, Kpad = Cin * 16, W % 4 == 0, oh % 2 == 0): one float4 load IS the
// four kx of one (patch, ci, ky), i.e. 8 contiguous output bytes. A block iteration covers TWO rows of patches of one image:
// all its loads (6 x 16 B per thread) are issued before the first use, the [2][ow][Kpad] tile is transposed through shared
// memory (rows padded by 8 B: conflict-free 8-byte stores) and leaves as fully coalesced 8-byte stores.
template <int CIN>
__global__ void __launch_bounds__(256) patchify_nchw4_kernel(const float* __restrict__ img, uint2* __restrict__ dst, int B, int H, int W) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char patch_sh[];
  uint2* tile = reinterpret_cast<uint2*>(patch_sh);
  constexpr int U = CIN * 4;                 // 8-byte units per patch row (Kpad * 2 / 8)
  constexpr int UP = U + 1;                  // padded pitch
  const int ow = W >> 2, oh = H >> 2;
  const int per_ch = 8 * ow;                 // float4 loads per channel per iteration (8 image rows)
  const int n_ld = CIN * per_ch;
  const int n_st = 2 * ow * U;
  for (int blk = blockIdx.x; blk < B * (oh >> 1); blk += gridDim.x) {
    const int oy2 = blk % (oh >> 1), b = blk / (oh >> 1);
    const float4* base = reinterpret_cast<const float4*>(img + ((long long)b * CIN * H + (long long)oy2 * 8) * W);
    __syncthreads();                         // the previous iteration's tile has been read
    for (int i0 = threadIdx.x; i0 < n_ld; i0 += 6 * blockDim.x) {
      float4 v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int i = i0 + j * blockDim.x;
        if (i < n_ld) {
          const int ci = i / per_ch, f = i - ci * per_ch;
          v[j] = __ldg(base + (long long)ci * (H * (W >> 2)) + f);
        }
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int i = i0 + j * blockDim.x;
        if (i < n_ld) {
          const int ci = i / per_ch, f = i - ci * per_ch;
          const int yl = f / ow, ox = f - yl * ow;
          tile[((yl >> 2) * ow + ox) * UP + ci * 4 + (yl & 3)] = make_uint2(pack_bf16x2(v[j].x, v[j].y), pack_bf16x2(v[j].z, v[j].w));
        }
      }
    }
    __syncthreads();
    uint2* out = dst + (long long)blk * n_st;
    for (int e = threadIdx.x; e < n_st; e += blockDim.x) {
      const int pch = e / U, j = e - pch * U;
      out[e] = tile[pch * UP + j];
    }
  }
}

// ---- generic strided row copy with dtype conversion ----------------------------------------------------------
struct RowMap2 { int group, stride, offset; };
__device__ __forceinline__ long long map_row2(const RowMap2& m, long long r) {
  if (m.stride == m.group && m.offset == 0) return r;   // identity map: no 64-bit division per element
  const long long q = r / m.group;
  return q * m.stride + m.offset + (r - q * m.group);
}
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) copy_rows_kernel(const TI* __restrict__ src, RowMap2 sm, long long lds,
                                                        TO* __restrict__ dst, RowMap2 dm, long long ldd, long long rows,
                                                        int C, int accumulate) {
  pdl_prologue();
  const int c4n = C / 4;
  const long long total = rows * c4n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((unsigned int)i % (unsigned int)c4n) * 4;
    const long long r = (long long)((unsigned int)i / (unsigned int)c4n);
    const TI* s = src + map_row2(sm, r) * lds + c;
    TO* d = dst + map_row2(dm, r) * ldd + c;
    float4 v;
    if constexpr (sizeof(TI) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(s);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      v = make_float4(a.x, a.y, b.x, b.y);
    } else {
      v = *reinterpret_cast<const float4*>(s);
    }
    if constexpr (sizeof(TO) == 2) {
      if (accumulate) {
        const uint2 u0 = *reinterpret_cast<const uint2*>(d);
        const float2 a = unpack_bf16x2(u0.x), b = unpack_bf16x2(u0.y);
        v.x += a.x; v.y += a.y; v.z += b.x; v.w += b.y;
      }
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(d) = u;
    } else {
      if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(d);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *reinterpret_cast<float4*>(d) = v;
    }
  }
}

// ---- out[r, c] (+)= sum_b x[b*batch_stride + r*C + c]  (position-embedding gradients) -------------------------
__global__ void __launch_bounds__(256) batch_reduce_kernel(const float* __restrict__ x, long long batch_stride, int B,
                                                           long long n4, float* __restrict__ out, int accumulate, int bchunk) {
  pdl_prologue();
  // blockIdx.y owns samples [y*bchunk, (y+1)*bchunk); with more than one chunk the partial sums meet in `out` through
  // vector atomics (the caller zeroes `out` first when it is not accumulating).
  const int b0 = blockIdx.y * bchunk, b1 = min(B, b0 + bchunk);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xi = x + i * 4;
    int b = b0;
    for (; b + 8 <= b1; b += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(xi + (long long)(b + u) * batch_stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; b < b1; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(xi + (long long)b * batch_stride);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* o = out + i * 4;
    if (gridDim.y > 1) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
    } else {
      if (accumulate) {
        const float4 p = *reinterpret_cast<const float4*>(o);
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
      }
      *reinterpret_cast<float4*>(o) = acc;
    }
  }
}

// ---- bilinear resize of a [h*w, C] table to [H*W, C], align_corners=False (torch upsample_bilinear2d) ---------
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = ((float)dst + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}
__global__ void pos_resize_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst, int h, int w, int H, int W,
                                      int C) {
  pdl_prologue();
  const long long total = (long long)H * W * C;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int X = (int)((i / C) % W), Y = (int)(i / ((long long)C * W));
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(Y, sy, h, y0, y1, ly);
    src_index(X, sx, w, x0, x1, lx);
    const float v00 = src[((long long)y0 * w + x0) * C + c], v01 = src[((long long)y0 * w + x1) * C + c];
    const float v10 = src[((long long)y1 * w + x0) * C + c], v11 = src[((long long)y1 * w + x1) * C + c];
    dst[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}
__global__ void pos_resize_bwd_kernel(const float* __restrict__ dsrc_resized, float* __restrict__ dtable, int h, int w,
                                      int H, int W, int C) {
  pdl_prologue();
  const long long total = (long long)H * W * C;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int X = (int)((i / C) % W), Y = (int)(i / ((long long)C * W));
    int y0, y1, x0, x1;
    float ly, lx;
    src_index(Y, sy, h, y0, y1, ly);
    src_index(X, sx, w, x0, x1, lx);
    const float g = dsrc_resized[i];
    atomicAdd(dtable + ((long long)y0 * w + x0) * C + c, g * (1.f - ly) * (1.f - lx));
    atomicAdd(dtable + ((long long)y0 * w + x1) * C + c, g * (1.f - ly) * lx);
    atomicAdd(dtable + ((long long)y1 * w + x0) * C + c, g * ly * (1.f - lx));
    atomicAdd(dtable + ((long long)y1 * w + x1) * C + c, g * ly * lx);
  }
}

// ---- weight preparation: fp32 master -> bf16 compute copy, optionally permuting conv [Co,Ci,kh,kw] -> [Co,kh,kw,Ci]
__global__ void cast_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}
__global__ void cast_conv_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Co, int Ci,
                                        int KK, int dst_ld) {
  pdl_prologue();
  const long long n = (long long)Co * Ci * KK;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int kk = (int)((i / Ci) % KK);
    const int co = (int)(i / ((long long)Ci * KK));
    dst[(long long)co * dst_ld + (long long)kk * Ci + ci] = __float2bfloat16(src[((long long)co * Ci + ci) * KK + kk]);
  }
}
// transposed + spatially flipped copy for the input-gradient convolution: dst[ci, t*Co + co] = src[co, ci, KK-1-t]
__global__ void cast_conv_weight_t_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Co, int Ci,
                                          int KK) {
  pdl_prologue();
  const long long n = (long long)Co * Ci * KK;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Co);
    const int t = (int)((i / Co) % KK);
    const int ci = (int)(i / ((long long)Co * KK));
    dst[i] = __float2bfloat16(src[((long long)co * Ci + ci) * KK + (KK - 1 - t)]);
  }
}
// gradient of the permuted copy back to the master layout: dW[co,ci,kk] += dWp[co,kk,ci]
__global__ void uncast_conv_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int KK,
                                         int src_ld) {
  pdl_prologue();
  const long long n = (long long)Co * Ci * KK;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % KK);
    const int ci = (int)((i / KK) % Ci);
    const int co = (int)(i / ((long long)Ci * KK));
    dw[i] += dwp[(long long)co * src_ld + (long long)kk * Ci + ci];
  }
}

// ---- out = dy * gelu_erf'(pre), all bf16 -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ pre,
                                                       __nv_bfloat16* __restrict__ out, long long n2) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 d = unpack_bf16x2(reinterpret_cast<const uint32_t*>(dy)[i]);
    const float2 x = unpack_bf16x2(reinterpret_cast<const uint32_t*>(pre)[i]);
    reinterpret_cast<uint32_t*>(out)[i] = pack_bf16x2(d.x * dgelu_fast(x.x), d.y * dgelu_fast(x.y));
  }
}

// ---- stand-alone exact-erf GELU (libs/vl_heads.py:7-14 GELU.forward) and its backward on fp32 / bf16 arrays ---------
template <typename T>
__global__ void __launch_bounds__(256) gelu_ew_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ out, long long n) {
  pdl_prologue();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = (float)x[i];
    out[i] = (T)(dy ? (float)dy[i] * dgelu_fast(v) : gelu_erf(v));
  }
}

// ---- generic 2-D cast with arbitrary (unaligned) widths: dst[r*ldd + c] = (TO) src[r*lds + c] * alpha ----------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast2d_kernel(const TI* __restrict__ src, long long lds, TO* __restrict__ dst,
                                                     long long ldd, long long rows, int C, float alpha) {
  pdl_prologue();
  const long long total = rows * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i % C);
    float v;
    if constexpr (sizeof(TI) == 2) v = __bfloat162float(src[r * lds + c]);
    else v = src[r * lds + c];
    v *= alpha;
    if constexpr (sizeof(TO) == 2) dst[r * ldd + c] = __float2bfloat16(v);
    else dst[r * ldd + c] = v;
  }
}

}  // namespace

// out = gelu(x) (dy == nullptr) or out = dy * gelu'(x); f32 selects fp32 (else bf16) for all three arrays
extern "C" int mvlt_gelu_ew(const void* x, const void* dy, void* out, long long n, int f32, void* stream_) {
  MVLT_CHECK_ARG(x && out && n > 0, "gelu_ew: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (f32)
    mvlt_launch(gelu_ew_kernel<float>, cap_grid(n, 256), 256, 0, st, reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(dy), reinterpret_cast<float*>(out), n);
  else
    mvlt_launch(gelu_ew_kernel<__nv_bfloat16>, cap_grid(n, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<__nv_bfloat16*>(out), n);
  MVLT_CHECK_LAUNCH();
  return 0;
}
extern "C" int mvlt_gelu_bwd(const void* dy_bf16, const void* pre_bf16, void* out_bf16, long long n, void* stream_) {
  MVLT_CHECK_ARG(n % 2 == 0, "gelu_bwd: n must be even");
  mvlt_launch(gelu_bwd_kernel, cap_grid(n / 2, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(dy_bf16), reinterpret_cast<const __nv_bfloat16*>(pre_bf16),
      reinterpret_cast<__nv_bfloat16*>(out_bf16), n / 2);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_cast2d(const void* src, int src_f32, long long lds, void* dst, int dst_f32, long long ldd,
                           long long rows, int C, float alpha, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = cap_grid(rows * C, 256);
#define LAUNCH(TI, TO) mvlt_launch(cast2d_kernel<TI, TO>, grid, 256, 0, st, reinterpret_cast<const TI*>(src), lds, reinterpret_cast<TO*>(dst), ldd, rows, C, alpha)
  if (src_f32 && dst_f32) LAUNCH(float, float);
  else if (src_f32) LAUNCH(float, __nv_bfloat16);
  else if (dst_f32) LAUNCH(__nv_bfloat16, float);
  else LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_memset_zero(void* p, long long bytes, void* stream_) {
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, reinterpret_cast<cudaStream_t>(stream_));
  if (e != cudaSuccess) {
    mvlt_set_error("memset failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

// ---- small host -> device parameter block, passed BY VALUE in the kernel arguments ------------------------------------
// Per-step scalars that kernels of a captured CUDA graph read from device memory (dropout / drop-path seeds, AdamW learning
// rate and bias corrections) are refreshed by this one-thread launch ahead of the replay: the 64 bytes travel inside the
// launch itself, so there is no pinned staging buffer whose lifetime the caller would have to manage.
struct SetValuesBlock { unsigned int w[16]; };
__global__ void set_values_kernel(unsigned int* __restrict__ dst, const SetValuesBlock v, int nwords) {
  pdl_prologue();
  if (threadIdx.x < nwords) dst[threadIdx.x] = v.w[threadIdx.x];
}
// dst: device pointer (4-byte aligned); src_host: nbytes (a multiple of 4, <= 64) of host memory, read before the call returns
extern "C" int mvlt_set_values(void* dst, const void* src_host, int nbytes, void* stream_) {
  MVLT_CHECK_ARG(dst && src_host && nbytes > 0 && nbytes <= 64 && nbytes % 4 == 0, "set_values: 4..64 bytes, a multiple of 4");
  SetValuesBlock v;
  memset(&v, 0, sizeof(v));
  memcpy(v.w, src_host, (size_t)nbytes);
  mvlt_launch(set_values_kernel, 1, 32, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<unsigned int*>(dst), v, nbytes / 4);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// ---- measurement aid: keep the stream busy for `ns` nanoseconds -------------------------------------------------------------
// bench.py's instrumented pass puts a CUDA-event pair around every launch; launched one by one from Python the GPU would
// idle between kernels and every event interval would include the host's launch latency. A spin ahead of the step lets the
// host enqueue the whole step first, so that the intervals measure kernels, not the launch path.
__global__ void spin_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(2000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}
extern "C" int mvlt_spin(long long ns, void* stream_) {
  MVLT_CHECK_ARG(ns >= 0 && ns <= 2000000000LL, "spin: 0 .. 2 s");
  spin_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream_)>>>((unsigned long long)ns);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_cast_scale_bf16(const float* src, void* dst, long long rows, int C, const float* rowscale,
                                    int rows_per_scale, float alpha, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0, "cast_scale: C=%d must be a multiple of 8", C);
  const long long n8 = rows * C / 8;
  mvlt_launch(cast_scale_kernel, cap_grid(n8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), src, reinterpret_cast<__nv_bfloat16*>(dst), n8, C, rowscale, rows_per_scale > 0 ? rows_per_scale : 1, alpha);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_colsum(const void* x, int x_f32, long long rows, int C, long long ld, float* out, void* stream_) {
  MVLT_CHECK_ARG(ld % 8 == 0 && (((uintptr_t)x) & 15) == 0, "colsum: ld must be a multiple of 8 and x 16-byte aligned");
  const int groups = (C + 7) / 8;           // 8-column groups (reads may touch the ld padding, results are masked)
  int tx_n = groups < 32 ? groups : 32;
  while (256 % tx_n != 0) --tx_n;
  const int gx = (groups + tx_n - 1) / tx_n;
  long long gy = (long long)mvlt_num_sms() * 8 / gx;
  if (gy < 1) gy = 1;
  const int ry = 256 / tx_n;
  long long rpb = (rows + gy - 1) / gy;
  if (rpb < 4LL * ry) rpb = 4LL * ry;
  gy = (rows + rpb - 1) / rpb;
  dim3 grid(gx, (unsigned)gy);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (x_f32) mvlt_launch(colsum_kernel<float>, grid, 256, 0, st, reinterpret_cast<const float*>(x), rows, C, ld, out, rpb, tx_n);
  else mvlt_launch(colsum_kernel<__nv_bfloat16>, grid, 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(x), rows, C, ld, out, rpb, tx_n);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_patchify(const void* src, int src_f32, long long src_batch_stride, void* dst, int B, int H, int W,
                             int C, int R, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0 && H % R == 0 && W % R == 0, "patchify: bad shape H=%d W=%d C=%d R=%d", H, W, C, R);
  const long long total = (long long)B * H * W * (C / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (src_f32)
    mvlt_launch(patchify_kernel<float>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const float*>(src), src_batch_stride,
                                                                 reinterpret_cast<__nv_bfloat16*>(dst), B, H, W, C, R);
  else
    mvlt_launch(patchify_kernel<__nv_bfloat16>, cap_grid(total, 256), 256, 0, st, reinterpret_cast<const __nv_bfloat16*>(src), src_batch_stride, reinterpret_cast<__nv_bfloat16*>(dst), B, H, W, C, R);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_unpatchify(const void* src_bf16, float* dst, long long dst_batch_stride, int B, int H, int W, int C,
                               int R, int accumulate, void* stream_) {
  MVLT_CHECK_ARG(C % 8 == 0 && H % R == 0 && W % R == 0, "unpatchify: bad shape");
  const long long total = (long long)B * H * W * (C / 8);
  mvlt_launch(unpatchify_kernel, cap_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), reinterpret_cast<const __nv_bfloat16*>(src_bf16), dst, dst_batch_stride, B, H, W, C, R, accumulate);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_patchify_nchw(const float* img, void* dst_bf16, int B, int Cin, int H, int W, int P, int Kpad,
                                  void* stream_) {
  MVLT_CHECK_ARG(H % P == 0 && W % P == 0 && Kpad >= Cin * P * P, "patchify_nchw: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(Kpad % 8 == 0 && Kpad >= Cin * P * P && W % P == 0 && H % P == 0 && (W / P) * Kpad * 2 <= 96 * 1024,
                 "patchify_nchw: unsupported geometry (W=%d P=%d Kpad=%d)", W, P, Kpad);
  if (P == 4 && Cin == 3 && Kpad == Cin * 16 && W % 4 == 0 && (H / 4) % 2 == 0 && ((uintptr_t)img & 15) == 0 && ((uintptr_t)dst_bf16 & 7) == 0) {
    int grid4 = B * (H / 8);
    const int cap4 = mvlt_num_sms() * 6;
    if (grid4 > cap4) grid4 = cap4;
    const size_t smem4 = (size_t)2 * (W / 4) * (Cin * 4 + 1) * 8;
    MVLT_CHECK_ARG(smem4 <= 48 * 1024, "patchify_nchw: image too wide (W=%d)", W);
    mvlt_launch(patchify_nchw4_kernel<3>, grid4, 256, smem4, st, img, reinterpret_cast<uint2*>(dst_bf16), B, H, W);
    MVLT_CHECK_LAUNCH();
    return 0;
  }
  const size_t smem = (size_t)(W / P) * Kpad * 2;
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaFuncSetAttribute(patchify_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  int grid = B * (H / P);
  const int cap = mvlt_num_sms() * 8;
  if (grid > cap) grid = cap;
  mvlt_launch(patchify_nchw_kernel, grid, 256, smem, st, img, reinterpret_cast<__nv_bfloat16*>(dst_bf16), B, Cin, H, W, P, Kpad);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_copy_rows(const void* src, int src_f32, const int* smap, long long lds, void* dst, int dst_f32,
                              const int* dmap, long long ldd, long long rows, int C, int accumulate, void* stream_) {
  MVLT_CHECK_ARG(C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "copy_rows: C/ld must be multiples of 4");
  auto mk = [&](const int* m) {
    const int big = 1 << 30;
    return RowMap2{m && m[0] > 0 ? m[0] : big, m && m[0] > 0 ? m[1] : big, m && m[0] > 0 ? m[2] : 0};
  };
  RowMap2 sm = mk(smap), dm = mk(dmap);
  const long long total = rows * (C / 4);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = cap_grid(total, 256);
#define LAUNCH(TI, TO)                                                                                          \
  mvlt_launch(copy_rows_kernel<TI, TO>, grid, 256, 0, st, reinterpret_cast<const TI*>(src), sm, lds,                     \
                                                 reinterpret_cast<TO*>(dst), dm, ldd, rows, C, accumulate)
  if (src_f32 && dst_f32) LAUNCH(float, float);
  else if (src_f32) LAUNCH(float, __nv_bfloat16);
  else if (dst_f32) LAUNCH(__nv_bfloat16, float);
  else LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef LAUNCH
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_batch_reduce(const float* x, long long batch_stride, int B, long long n, float* out, int accumulate,
                                 void* stream_) {
  MVLT_CHECK_ARG(n % 4 == 0 && batch_stride % 4 == 0, "batch_reduce: n must be a multiple of 4");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  // enough (column-block x batch-chunk) CTAs to fill the machine: split the batch when the row is short
  const long long xblocks = (n / 4 + 255) / 256;
  int chunks = 1;
  while (xblocks * chunks < 4LL * mvlt_num_sms() && chunks * 16 <= B) chunks *= 2;
  const int bchunk = (B + chunks - 1) / chunks;
  chunks = (B + bchunk - 1) / bchunk;
  if (chunks > 1 && !accumulate) cudaMemsetAsync(out, 0, (size_t)n * sizeof(float), st);
  dim3 grid((unsigned)(xblocks < 65535 ? xblocks : 65535), (unsigned)chunks);
  mvlt_launch(batch_reduce_kernel, grid, 256, 0, st, x, batch_stride, B, n / 4, out, accumulate, bchunk);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_pos_resize_fwd(const float* table, float* out, int h, int w, int H, int W, int C, void* stream_) {
  const long long total = (long long)H * W * C;
  mvlt_launch(pos_resize_fwd_kernel, cap_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), table, out, h, w, H, W, C);
  MVLT_CHECK_LAUNCH();
  return 0;
}
extern "C" int mvlt_pos_resize_bwd(const float* dout, float* dtable, int h, int w, int H, int W, int C, void* stream_) {
  const long long total = (long long)H * W * C;
  mvlt_launch(pos_resize_bwd_kernel, cap_grid(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), dout, dtable, h, w, H, W, C);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_cast_weight(const float* src, void* dst_bf16, long long n, void* stream_) {
  mvlt_launch(cast_weight_kernel, cap_grid(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), n);
  MVLT_CHECK_LAUNCH();
  return 0;
}
extern "C" int mvlt_cast_conv_weight(const float* src, void* dst_bf16, int Co, int Ci, int KK, int dst_ld, void* stream_) {
  mvlt_launch(cast_conv_weight_kernel, cap_grid((long long)Co * Ci * KK, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), Co, Ci, KK, dst_ld);
  MVLT_CHECK_LAUNCH();
  return 0;
}
extern "C" int mvlt_cast_conv_weight_t(const float* src, void* dst_bf16, int Co, int Ci, int KK, void* stream_) {
  mvlt_launch(cast_conv_weight_t_kernel, cap_grid((long long)Co * Ci * KK, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), Co, Ci, KK);
  MVLT_CHECK_LAUNCH();
  return 0;
}
// ---- the same fold for up to 24 convolution weights in one launch (blockIdx.y = tensor) --------------------------
struct UncastEntry {
  const float* dwp;
  float* dw;
  int Co, Ci, KK, src_ld;
};
struct UncastBatch {
  int n;
  int pad;
  UncastEntry t[24];
};
__global__ void uncast_conv_wgrad_multi_kernel(const UncastBatch b) {
  pdl_prologue();
  const UncastEntry e = b.t[blockIdx.y];
  const long long n = (long long)e.Co * e.Ci * e.KK;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % e.KK);
    const int ci = (int)((i / e.KK) % e.Ci);
    const int co = (int)(i / ((long long)e.Ci * e.KK));
    e.dw[i] += e.dwp[(long long)co * e.src_ld + (long long)kk * e.Ci + ci];
  }
}
// entries: host array of n records {const float* dwp; float* dw; int Co, Ci, KK, src_ld;} (32 bytes each), n <= 24
extern "C" int mvlt_uncast_conv_wgrad_multi(const void* entries, int n, void* stream_) {
  MVLT_CHECK_ARG(entries != nullptr && n > 0 && n <= 24, "uncast_conv_wgrad_multi: 1..24 tensors per launch");
  UncastBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n;
  memcpy(b.t, entries, (size_t)n * sizeof(UncastEntry));
  long long mx = 0;
  for (int i = 0; i < n; ++i) {
    const long long e = (long long)b.t[i].Co * b.t[i].Ci * b.t[i].KK;
    if (e > mx) mx = e;
  }
  dim3 grid((unsigned)cap_grid(mx, 256, 2), (unsigned)n);
  mvlt_launch(uncast_conv_wgrad_multi_kernel, grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_), b);
  MVLT_CHECK_LAUNCH();
  return 0;
}
extern "C" int mvlt_uncast_conv_wgrad(const float* dwp, float* dw, int Co, int Ci, int KK, int src_ld, void* stream_) {
  mvlt_launch(uncast_conv_wgrad_kernel, cap_grid((long long)Co * Ci * KK, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream_), dwp, dw, Co, Ci, KK, src_ld);
  MVLT_CHECK_LAUNCH();
  return 0;
}
