// Shared device/host helpers for the mvlt_b200 sm_100a kernels.
// Everything here is hand-written for Blackwell (B200): mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// 128-bit vector I/O and warp-shuffle reductions.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

// ---------------------------------------------------------------------------------------------
// Error plumbing for the C-ABI: 0 = ok, <0 = argument/arch error, >0 = cudaError_t.
// ---------------------------------------------------------------------------------------------
void mvlt_set_error(const char* fmt, ...);
#define MVLT_ERR_ARG (-1)
#define MVLT_ERR_ALIGN (-2)
#define MVLT_ERR_DRIVER (-3)
#ifndef MVLT_PDL_DEFAULT
#define MVLT_PDL_DEFAULT 1
#endif

#define MVLT_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      mvlt_set_error(__VA_ARGS__);                \
      return MVLT_ERR_ARG;                        \
    }                                             \
  } while (0)

#define MVLT_CHECK_LAUNCH()                                                   \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      mvlt_set_error("%s:%d CUDA launch failed: %s", __FILE__, __LINE__,      \
                     cudaGetErrorString(e__));                                \
      return (int)e__;                                                        \
    }                                                                         \
  } while (0)

static inline int mvlt_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// Kernel launch with programmatic dependent launch (PDL). Every kernel of this library starts with
// pdl_prologue() (or pdl_trigger() ... pdl_wait() around a prologue that touches no global memory):
//   griddepcontrol.launch_dependents  - the NEXT kernel of the stream may be scheduled as soon as every CTA of this
//                                       grid has started, i.e. its CTAs fill the SMs this grid's tail leaves idle;
//   griddepcontrol.wait               - blocks until the PREVIOUS kernel has completed and its writes are visible;
//                                       no global memory is read or written before it.
// Launch latency, CTA scheduling and (GEMM / attention) barrier + TMEM set-up of kernel i+1 thus overlap the tail of
// kernel i (measured: 17.17 -> 16.55 ms per training step). Both instructions are no-ops when the launch attribute is
// absent (MVLT_PDL=0 in the environment switches it off for A/B runs).
// ---------------------------------------------------------------------------------------------
#include <stdlib.h>
static inline bool mvlt_pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MVLT_PDL");
    return e == nullptr ? (MVLT_PDL_DEFAULT != 0) : (e[0] != '0');
  }();
  return on;
}
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t mvlt_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mvlt_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// same, as thread-block clusters of `cluster_x` CTAs (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
static inline cudaError_t mvlt_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                              unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mvlt_pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// Small device utilities
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `sh` needs 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// d/dx [ x * Phi(x) ] = Phi(x) + x * phi(x)
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): one MUFU.RCP, one MUFU.EX2 and a degree-5 Horner chain
// instead of erff()'s ~60 instructions; the outputs below are rounded to bf16 anyway.
__device__ __forceinline__ float erf_poly_tail(float ax, float e) {  // returns (1 - erf(ax)) given e = exp(-ax^2)
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  return poly * t * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  const float e = __expf(-az * az);
  const float tail = erf_poly_tail(az, e);           // 1 - erf(|z|)
  const float erfz = copysignf(1.0f - tail, z);
  return 0.5f * x * (1.0f + erfz);
}
__device__ __forceinline__ float dgelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float az = fabsf(z);
  const float e = __expf(-az * az);                   // = exp(-x^2 / 2)
  const float tail = erf_poly_tail(az, e);
  const float cdf = 0.5f * (1.0f + copysignf(1.0f - tail, z));
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// gelu(x) and gelu'(x) together: they share the exponential and the erfc tail (16 FP32 ops + 2 MUFU per element).
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  const float w = fabsf(x) * 0.84932180028801904272f;        // |x|/sqrt(2) * sqrt(log2 e)
  const float e = exp2f(-w * w);                               // exp(-x^2/2)
  const float t = __frcp_rn(fmaf(0.2727374767f, w, 1.0f));       // 1 / (1 + 0.3275911 |x|/sqrt(2))
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);         // A&S 7.1.26 coefficients, pre-multiplied by 0.5
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  const float h = poly * t * e;                                // 0.5 * erfc(|x|/sqrt(2))
  const float cdf = x >= 0.f ? 1.0f - h : h;
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2/FMUL2/FADD2: one issue slot for two fp32 lanes) ------------------
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_splat(float c) { return f2_pack(c, c); }
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// gelu_erf(x) and gelu_erf'(x) for two values at once. Same Abramowitz-Stegun 7.1.26 erfc tail as gelu_and_grad
// (|error| <= 1.5e-7) but evaluated with packed fp32x2 FMAs and approximate MUFU ex2 / rcp:
// 15 packed FP ops + 4 MUFU + 2 LOP3 per PAIR instead of ~38 instructions per ELEMENT.
__device__ __forceinline__ void gelu_and_grad2(f32x2_t x, f32x2_t& g, f32x2_t& dg) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  const f32x2_t ax = f2_pack(fabsf(x0), fabsf(x1));
  const f32x2_t nw = f2_mul(f2_mul(x, x), f2_splat(-0.72134752044448170368f));   // -x^2/2 * log2(e)
  float n0, n1;
  f2_unpack(nw, n0, n1);
  const float e0 = ex2_approx(n0), e1 = ex2_approx(n1);                           // exp(-x^2/2)
  const f32x2_t e = f2_pack(e0, e1);
  const f32x2_t den = f2_fma(ax, f2_splat(0.23164189f), f2_splat(1.0f));         // 1 + 0.3275911 |x|/sqrt(2)
  float d0, d1;
  f2_unpack(den, d0, d1);
  const f32x2_t t = f2_pack(rcp_approx(d0), rcp_approx(d1));
  f32x2_t poly = f2_fma(f2_splat(0.5307027145f), t, f2_splat(-0.7265760135f));   // coefficients pre-multiplied by 0.5
  poly = f2_fma(poly, t, f2_splat(0.7107068705f));
  poly = f2_fma(poly, t, f2_splat(-0.142248368f));
  poly = f2_fma(poly, t, f2_splat(0.127414796f));
  const f32x2_t h = f2_mul(f2_mul(poly, t), e);                                  // 0.5 * erfc(|x|/sqrt(2))
  // Phi(x) = 0.5 + sign(x) * (0.5 - h): the sign is OR-ed in (0.5 - h >= 0), no compare / select, no register shuffling
  const f32x2_t q = f2_fma(h, f2_splat(-1.0f), f2_splat(0.5f));
  float q0, q1;
  f2_unpack(q, q0, q1);
  q0 = __uint_as_float(__float_as_uint(q0) | (__float_as_uint(x0) & 0x80000000u));
  q1 = __uint_as_float(__float_as_uint(q1) | (__float_as_uint(x1) & 0x80000000u));
  const f32x2_t cdf = f2_add(f2_pack(q0, q1), f2_splat(0.5f));
  g = f2_mul(x, cdf);
  dg = f2_fma(f2_mul(x, e), f2_splat(0.39894228040143267794f), cdf);
}
// gelu_erf(x) alone: gelu(x) = relu(x) - |x| * h with h = 0.5 * erfc(|x| / sqrt(2)) (h is the lower tail on both sides), which
// needs no sign handling: 12 packed FP ops + 4 MUFU + 4 scalar ops per PAIR. Same A&S 7.1.26 tail as gelu_and_grad2, with
// the polynomial coefficients negated so that the last FMA subtracts.
__device__ __forceinline__ f32x2_t gelu2(f32x2_t x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  const f32x2_t ax = f2_pack(fabsf(x0), fabsf(x1));
  const f32x2_t nw = f2_mul(f2_mul(x, x), f2_splat(-0.72134752044448170368f));   // -x^2/2 * log2(e)
  float n0, n1;
  f2_unpack(nw, n0, n1);
  const f32x2_t e = f2_pack(ex2_approx(n0), ex2_approx(n1));                      // exp(-x^2/2)
  const f32x2_t den = f2_fma(ax, f2_splat(0.23164189f), f2_splat(1.0f));         // 1 + 0.3275911 |x|/sqrt(2)
  float d0, d1;
  f2_unpack(den, d0, d1);
  const f32x2_t t = f2_pack(rcp_approx(d0), rcp_approx(d1));
  f32x2_t poly = f2_fma(f2_splat(-0.5307027145f), t, f2_splat(0.7265760135f));   // -(0.5 * A&S coefficients)
  poly = f2_fma(poly, t, f2_splat(-0.7107068705f));
  poly = f2_fma(poly, t, f2_splat(0.142248368f));
  poly = f2_fma(poly, t, f2_splat(-0.127414796f));
  const f32x2_t nh = f2_mul(f2_mul(poly, t), e);                                 // -0.5 * erfc(|x|/sqrt(2))
  return f2_fma(ax, nh, f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}

// 128-bit streaming loads/stores (data touched once: keep it out of L1)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

// counter-based RNG (splitmix-style hash) used for dropout / drop-path; NOT torch's Philox stream
// (parity tests run with p = 0, production draws its own stream).
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
  return (hash_u32(seed, idx) >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------------------------
// mbarrier / TMA / tcgen05 PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 4-D tiled TMA load, completion on an mbarrier (bytes). c0 is the innermost coordinate.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// 5-D tiled TMA load (strided patch views: gemm_desc.h MVLT_CONV_PATCH_A)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// 5-D tiled TMA store (the input-gradient side of the strided patch view: gemm_desc.h patch_store)
__device__ __forceinline__ void tma_store_5d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(tmap), "r"(smem_src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// 4-D tiled TMA store (shared -> global, bulk async-group completion). c0 is the innermost coordinate.
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(smem_src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// same, as an element-wise fp32 add into global memory (the tensor map's data type selects the arithmetic)
__device__ __forceinline__ void tma_reduce_add_4d(const void* tmap, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the TMA engine has finished READING the shared-memory source of all committed bulk stores
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Whole-warp TMEM allocation of `ncols` columns (power of two >= 32); base address lands in *dst_smem.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two CTAs of one cluster on the two SMs of a TPC issue ONE 256-row UMMA ---------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {   // remote (or local) arrive through the cluster window
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; its bytes complete on the mbarrier at `bar_cluster_addr` (the leader's)
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {   // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs, 256 rows] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]; issued by one thread of the LEADER CTA
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this shared-memory offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#endif  // __CUDACC__
