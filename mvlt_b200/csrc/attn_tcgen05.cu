// Fused spatial-reduction attention forward for sm_100a (B200): O = softmax(scale * Q K^T) V in ONE kernel.
//
// PVLT's reduced key/value sequence is short (Nk = (H/32)(W/32) + T = 192 at 256x256 / 128 tokens, every stage), so a
// whole score row fits one TMEM accumulator: the softmax is a single tile ("online" softmax degenerates to one
// block) and the probabilities never visit HBM unless the caller asks for them (training keeps P for the backward).
//
// One CTA = 128 threads = one 128-row query tile at a time, two CTAs per SM (each allocates 256 of the 512 TMEM
// columns and ~112 KB of shared memory) so that the TMA / tensor-core phases of one CTA overlap the softmax of the other:
//
//   thread 0   : TMA loads (Q tile; K and V tiles only when the (batch, head) pair changes: tiles are handed out in
//                contiguous chunks, so a CTA walks the query tiles of one pair before moving on) and both tcgen05.mma
//                groups:  S[128 x Nk] = Q K^T  (K-major A and B, 4 k-steps of 16)   -> TMEM columns [0, Nk)
//                         O[128 x 64] = P V    (A = P from shared memory, B = V MN-major, Nk/16 k-steps) -> [192, 256)
//   all threads: thread r owns score row r (tcgen05.ld 32x32b): row max + sum in registers, P = exp2(a2 s - max) / sum
//                written as bf16 straight into the SWIZZLE_128B K-major layout the second MMA reads (and, for
//                training, the same tiles leave as TMA bulk stores into P[B, h, N, Nk]); then the O epilogue:
//                TMEM -> bf16 -> swizzled tile (the Q buffer, free by then) -> one TMA store.
//
// Replaces /root/reference/libs/pvlt.py:113-117 (attn = (q @ k^T) * scale; softmax; attn @ v; head merge).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_set>
#include "common.cuh"

int mvlt_tensor_map_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_b[3],
                       const uint32_t box[4], int f32, int swizzle64);   // gemm_tcgen05.cu

namespace {

constexpr int BM = 128;          // query rows per tile = TMEM lanes
constexpr int HD = 64;           // head dimension (PVLT: C / heads = 64 at every stage): one SWIZZLE_128B row of bf16
constexpr int NK_MAX = 192;      // keys per (batch, head)
constexpr int Q_BYTES = BM * HD * 2;           // 16 KB, also the O staging tile
constexpr int KV_BYTES = NK_MAX * HD * 2;      // 24 KB each
constexpr int P_ATOM = BM * 128;               // 128 rows x 64 bf16 columns
constexpr int P_BYTES = (NK_MAX / 64) * P_ATOM;
constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + KV_BYTES, OFF_P = OFF_V + KV_BYTES, OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_USED = OFF_BAR + 64;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 192;       // accumulator columns of O (S occupies [0, Nk <= 192))
constexpr int THREADS = 128;

struct AttnParams {
  int B, heads, N, Nk, C;
  int num_m, total_tiles;
  float a2;        // softmax scale * log2(e)
};

// smem matrix descriptor (SWIZZLE_128B) and kind::f16 instruction descriptor: same encodings as gemm_tcgen05.cu
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ uint32_t instr_desc(int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;    // fp32 accumulator
  d |= 1u << 7;    // A = bf16
  d |= 1u << 10;   // B = bf16
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

// ---- v1 (kept selectable with MVLT_ATTN_V1=1 for A/B runs): runtime Nk, exponentials evaluated twice (running max/sum
// pass + normalising pass), Q loaded at the top of each tile, O staged in the Q buffer.
template <bool kStoreP>
__global__ void __launch_bounds__(THREADS, 2)
sr_attention_fwd_kernel_v1(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP,
                        const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((uint32_t)(smem - smem_raw) + (uint32_t)SMEM_USED > dyn) __trap();   // base not aligned as declared
  }
  pdl_trigger();   // the next kernel of the stream may be scheduled behind this grid's tail (common.cuh)
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sP = smem + OFF_P;
  uint64_t* q_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_bar = q_bar + 1;
  uint64_t* s_bar = q_bar + 2;
  uint64_t* o_bar = q_bar + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 4);

  if (tid == 0) {
    mbar_init(q_bar, 1);
    mbar_init(kv_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    if (kStoreP) tma_prefetch_desc(&tmP);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's TMEM lane quarter; lane = row
  pdl_wait();      // barrier / TMEM set-up above overlapped the previous kernel; global memory is only touched below

  const uint32_t sQ_s = smem_u32(sQ), sK_s = smem_u32(sK), sV_s = smem_u32(sV), sP_s = smem_u32(sP);
  const uint32_t idesc_s = instr_desc(p.Nk, 0, 0);
  const uint32_t idesc_o = instr_desc(HD, 0, 1);
  const int nchunk32 = p.Nk >> 5, nchunk16 = p.Nk >> 4;
  const uint32_t kv_bytes = (uint32_t)p.Nk * HD * 2;
  const uint32_t row_s = (uint32_t)tid * 128u, row_x = (uint32_t)(tid & 7);

  const int t_begin = (int)((long long)blockIdx.x * p.total_tiles / gridDim.x);
  const int t_end = (int)((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x);
  int cur_bh = -1;
  uint32_t kv_phase = 0;

  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const uint32_t ph = (uint32_t)it & 1u;
    const int bh = t / p.num_m, m0 = (t - bh * p.num_m) * BM;
    const int b = bh / p.heads, h = bh - b * p.heads;

    if (tid == 0) {
      const bool new_kv = bh != cur_bh;
      if (new_kv) {   // every MMA that read sK / sV has retired: this thread waited on the previous tile's o_bar
        mbar_arrive_expect_tx(kv_bar, 2u * kv_bytes);
        tma_load_4d(sK, &tmKV, kv_bar, h * HD, 0, b, 0);
        tma_load_4d(sV, &tmKV, kv_bar, p.C + h * HD, 0, b, 0);
        cur_bh = bh;
      }
      mbar_arrive_expect_tx(q_bar, (uint32_t)Q_BYTES);
      tma_load_4d(sQ, &tmQ, q_bar, h * HD, m0, b, 0);
      if (new_kv) {
        mbar_wait(kv_bar, kv_phase);
        kv_phase ^= 1u;
      }
      mbar_wait(q_bar, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)   // S = Q K^T: both operands K-major, 8-row groups 1024 B apart
        umma_bf16(tmem_base, smem_desc(sQ_s + k * 32, 0u, 1024u), smem_desc(sK_s + k * 32, 0u, 1024u), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(s_bar);
    }
    __syncwarp();
    mbar_wait(s_bar, ph);
    tc_fence_after();

    // ---- row softmax: pass 1 = running (max, sum) over 32-column chunks, pass 2 = normalised bf16 P -> shared memory
    float mx = -INFINITY, sum = 0.f;
    for (int ci = 0; ci < nchunk32; ++ci) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + (uint32_t)(ci * 32), r);
      tmem_ld_wait();
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(r[j]) * p.a2);
      const float mn = fmaxf(mx, cm);
      float cs = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) cs += ex2_approx(fmaf(__uint_as_float(r[j]), p.a2, -mn));
      sum = fmaf(sum, ex2_approx(mx - mn), cs);
      mx = mn;
    }
    const float inv = 1.f / sum;
    for (int i = 0; i < nchunk16; ++i) {
      uint32_t r[16];
      tmem_ld_32x16(taddr + (uint32_t)(i * 16), r);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = pack_bf16x2(ex2_approx(fmaf(__uint_as_float(r[2 * j]), p.a2, -mx)) * inv,
                            ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), p.a2, -mx)) * inv);
      // K-major SWIZZLE_128B atom (128 rows x 64 columns): 16-byte piece q of row r at r * 128 + ((q ^ (r & 7)) << 4)
      const uint32_t base = sP_s + (uint32_t)(i >> 2) * (uint32_t)P_ATOM + row_s;
      const uint32_t q0 = (uint32_t)(i & 3) * 2u;
      st_shared_v4(base + (((q0) ^ row_x) << 4), pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(base + (((q0 + 1u) ^ row_x) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();     // P (generic-proxy stores) -> visible to the tensor core and the TMA engine
    tc_fence_before();       // this thread's TMEM reads of S precede the next MMA that overwrites it
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      for (int kk = 0; kk < nchunk16; ++kk)   // O = P V: A K-major (one atom per 64 keys), B = V[key, d] MN-major (2 KB per 16 keys)
        umma_bf16(tmem_base + O_COL, smem_desc(sP_s + (uint32_t)(kk >> 2) * (uint32_t)P_ATOM + (uint32_t)(kk & 3) * 32u, 0u, 1024u),
                  smem_desc(sV_s + (uint32_t)kk * 2048u, 8192u, 1024u), idesc_o, kk > 0 ? 1u : 0u);
      umma_commit(o_bar);
      if (kStoreP) {   // training keeps the probabilities for the backward: the same tiles, as bulk tensor stores
        for (int a = 0; a * 64 < p.Nk; ++a) tma_store_4d(&tmP, sP_s + (uint32_t)a * (uint32_t)P_ATOM, a * 64, m0, h, b);
        tma_store_commit();
      }
    }
    __syncwarp();
    mbar_wait(o_bar, ph);
    tc_fence_after();

    // ---- O epilogue: 64 fp32 columns -> bf16 -> swizzled tile in the Q buffer (Q was consumed by the first MMA)
#pragma unroll
    for (int i = 0; i < HD / 16; ++i) {
      uint32_t r[16];
      tmem_ld_32x16(taddr + (uint32_t)(O_COL + i * 16), r);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      const uint32_t base = sQ_s + row_s;
      st_shared_v4(base + (((uint32_t)(2 * i) ^ row_x) << 4), pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(base + (((uint32_t)(2 * i + 1) ^ row_x) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&tmO, sQ_s, h * HD, m0, b, 0);   // rows past N are clipped by the tensor map
      tma_store_commit();
      tma_store_wait_read();   // sQ and sP may be overwritten (next tile's TMA load / softmax)
    }
  }

  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- v2 (default). Same tile / barrier protocol as v1, with three changes:
//   * the key count is a template parameter, so the whole score row is walked by fully unrolled loops and every
//     exponential is evaluated ONCE: pass 1 = row max of the raw scores (FMNMX only), pass 2 = e = exp2(a2 s - max),
//     its running sum and a register cache of e as packed bf16 pairs, pass 3 = P = e / sum from the cache into the
//     swizzled shared-memory operand. MUFU work per row halves (the kernel is MUFU/issue-bound, not HBM-bound);
//   * TMEM loads are double-buffered in registers (the next 32-column chunk is in flight while the current one is used);
//   * O is staged in P atom 0 (free once the second MMA has retired and the P stores have been read), which frees the Q
//     buffer right after the first MMA: thread 0 prefetches the NEXT tile's Q during the softmax of the current one.
template <int NK, bool kStoreP>
__global__ void __launch_bounds__(THREADS, 2)
sr_attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmP,
                        const __grid_constant__ AttnParams p) {
  static_assert(NK % 32 == 0 && NK >= 32 && NK <= NK_MAX, "key count: multiple of 32, <= 192");
  constexpr int NC32 = NK / 32, NC16 = NK / 16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if ((uint32_t)(smem - smem_raw) + (uint32_t)SMEM_USED > dyn) __trap();   // base not aligned as declared
  }
  pdl_trigger();   // the next kernel of the stream may be scheduled behind this grid's tail (common.cuh)
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint64_t* q_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_bar = q_bar + 1;
  uint64_t* s_bar = q_bar + 2;
  uint64_t* o_bar = q_bar + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_bar + 4);

  if (tid == 0) {
    mbar_init(q_bar, 1);
    mbar_init(kv_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
    if (kStoreP) tma_prefetch_desc(&tmP);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's TMEM lane quarter; lane = row
  pdl_wait();      // barrier / TMEM set-up above overlapped the previous kernel; global memory is only touched below

  const uint32_t sQ_s = smem_u32(sQ), sK_s = smem_u32(sK), sV_s = smem_u32(sV), sP_s = smem_u32(smem + OFF_P);
  const uint32_t idesc_s = instr_desc(NK, 0, 0);
  const uint32_t idesc_o = instr_desc(HD, 0, 1);
  constexpr uint32_t kv_bytes = (uint32_t)NK * HD * 2;
  const uint32_t row_s = (uint32_t)tid * 128u, row_x = (uint32_t)(tid & 7);
  const float a2 = p.a2;

  const int t_begin = (int)((long long)blockIdx.x * p.total_tiles / gridDim.x);
  const int t_end = (int)((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x);
  int cur_bh = -1;
  uint32_t kv_phase = 0;

  if (tid == 0) {   // Q of the first tile; later tiles are prefetched one tile ahead
    const int bh = t_begin / p.num_m, m0 = (t_begin - bh * p.num_m) * BM;
    const int b = bh / p.heads, h = bh - b * p.heads;
    mbar_arrive_expect_tx(q_bar, (uint32_t)Q_BYTES);
    tma_load_4d(sQ, &tmQ, q_bar, h * HD, m0, b, 0);
  }

  for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
    const uint32_t ph = (uint32_t)it & 1u;
    const int bh = t / p.num_m, m0 = (t - bh * p.num_m) * BM;
    const int b = bh / p.heads, h = bh - b * p.heads;

    if (tid == 0) {
      if (bh != cur_bh) {   // every MMA that read sK / sV has retired: this thread waited on the previous tile's o_bar
        mbar_arrive_expect_tx(kv_bar, 2u * kv_bytes);
        tma_load_4d(sK, &tmKV, kv_bar, h * HD, 0, b, 0);
        tma_load_4d(sV, &tmKV, kv_bar, p.C + h * HD, 0, b, 0);
        cur_bh = bh;
        mbar_wait(kv_bar, kv_phase);
        kv_phase ^= 1u;
      }
      mbar_wait(q_bar, ph);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)   // S = Q K^T: both operands K-major, 8-row groups 1024 B apart
        umma_bf16(tmem_base, smem_desc(sQ_s + k * 32, 0u, 1024u), smem_desc(sK_s + k * 32, 0u, 1024u), idesc_s, k > 0 ? 1u : 0u);
      // the previous tile's O store must have finished reading P atom 0 before anyone rewrites it (pass 3 below): waiting
      // here, between the MMA issue and the commit every thread waits on, overlaps that read with the first MMA
      tma_store_wait_read();
      umma_commit(s_bar);
    }
    __syncwarp();
    mbar_wait(s_bar, ph);
    tc_fence_after();
    if (tid == 0 && t + 1 < t_end) {   // the first MMA has retired: the Q buffer is free for the next tile
      const int bh1 = (t + 1) / p.num_m, m1 = ((t + 1) - bh1 * p.num_m) * BM;
      const int b1 = bh1 / p.heads, h1 = bh1 - b1 * p.heads;
      mbar_arrive_expect_tx(q_bar, (uint32_t)Q_BYTES);
      tma_load_4d(sQ, &tmQ, q_bar, h1 * HD, m1, b1, 0);
    }
    __syncwarp();

    // ---- row softmax, one exponential per score
    uint32_t r[2][32];
    float mraw = -INFINITY;
    tmem_ld_32x32(taddr, r[0]);
#pragma unroll
    for (int ci = 0; ci < NC32; ++ci) {   // pass 1: row max of the raw scores (a2 > 0)
      tmem_ld_wait();
      if (ci + 1 < NC32) tmem_ld_32x32(taddr + (uint32_t)((ci + 1) * 32), r[(ci + 1) & 1]);
      else tmem_ld_32x32(taddr, r[(ci + 1) & 1]);   // first chunk of pass 2
#pragma unroll
      for (int j = 0; j < 32; ++j) mraw = fmaxf(mraw, __uint_as_float(r[ci & 1][j]));
    }
    const float nmx = -mraw * a2;
    uint32_t e[NK / 2];     // exp2(a2 s - max) as packed bf16 pairs (unnormalised, in (0, 1])
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int ci = 0; ci < NC32; ++ci) {   // pass 2 (its chunk 0 sits in r[NC32 & 1])
      tmem_ld_wait();
      if (ci + 1 < NC32) tmem_ld_32x32(taddr + (uint32_t)((ci + 1) * 32), r[(NC32 + ci + 1) & 1]);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = ex2_approx(fmaf(__uint_as_float(r[(NC32 + ci) & 1][2 * j]), a2, nmx));
        const float y = ex2_approx(fmaf(__uint_as_float(r[(NC32 + ci) & 1][2 * j + 1]), a2, nmx));
        s0 += x;
        s1 += y;
        e[16 * ci + j] = pack_bf16x2(x, y);
      }
    }
    const float inv = 1.f / (s0 + s1);
#pragma unroll
    for (int i = 0; i < NC16; ++i) {      // pass 3: P = e / sum -> K-major SWIZZLE_128B atoms (128 rows x 64 columns)
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = unpack_bf16x2(e[8 * i + j]);
        pk[j] = pack_bf16x2(f.x * inv, f.y * inv);
      }
      const uint32_t base = sP_s + (uint32_t)(i >> 2) * (uint32_t)P_ATOM + row_s;
      const uint32_t q0 = (uint32_t)(i & 3) * 2u;
      st_shared_v4(base + (((q0) ^ row_x) << 4), pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(base + (((q0 + 1u) ^ row_x) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();     // P (generic-proxy stores) -> visible to the tensor core and the TMA engine
    tc_fence_before();       // this thread's TMEM reads of S precede the next MMA that overwrites it
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < NC16; ++kk)   // O = P V: A K-major (one atom per 64 keys), B = V[key, d] MN-major (2 KB per 16 keys)
        umma_bf16(tmem_base + O_COL, smem_desc(sP_s + (uint32_t)(kk >> 2) * (uint32_t)P_ATOM + (uint32_t)(kk & 3) * 32u, 0u, 1024u),
                  smem_desc(sV_s + (uint32_t)kk * 2048u, 8192u, 1024u), idesc_o, kk > 0 ? 1u : 0u);
      umma_commit(o_bar);
      if (kStoreP) {   // training keeps the probabilities for the backward: the same tiles, as bulk tensor stores
#pragma unroll
        for (int a = 0; a * 64 < NK; ++a) tma_store_4d(&tmP, sP_s + (uint32_t)a * (uint32_t)P_ATOM, a * 64, m0, h, b);
        tma_store_commit();
      }
    }
    __syncwarp();
    mbar_wait(o_bar, ph);
    tc_fence_after();
    if (kStoreP) {           // O is staged in P atom 0: the P stores must have finished reading it
      if (tid == 0) tma_store_wait_read();
      __syncthreads();
    }

    // ---- O epilogue: 64 fp32 columns -> bf16 -> swizzled tile in P atom 0 -> one TMA store
    tmem_ld_32x32(taddr + (uint32_t)O_COL, r[0]);
    tmem_ld_32x32(taddr + (uint32_t)(O_COL + 32), r[1]);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[i >> 1][16 * (i & 1) + 2 * j]), __uint_as_float(r[i >> 1][16 * (i & 1) + 2 * j + 1]));
      const uint32_t base = sP_s + row_s;
      st_shared_v4(base + (((uint32_t)(2 * i) ^ row_x) << 4), pk[0], pk[1], pk[2], pk[3]);
      st_shared_v4(base + (((uint32_t)(2 * i + 1) ^ row_x) << 4), pk[4], pk[5], pk[6], pk[7]);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&tmO, sP_s, h * HD, m0, b, 0);   // rows past N are clipped by the tensor map
      tma_store_commit();      // its shared-memory read is awaited after the next tile's first MMA has been issued
    }
  }

  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

typedef void (*AttnKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);
template <int NK>
AttnKernelFn attn_variant(bool store_p) {
  return store_p ? sr_attention_fwd_kernel<NK, true> : sr_attention_fwd_kernel<NK, false>;
}
AttnKernelFn pick_attn_kernel(int Nk, bool store_p) {
  switch (Nk) {
    case 32: return attn_variant<32>(store_p);
    case 64: return attn_variant<64>(store_p);
    case 96: return attn_variant<96>(store_p);
    case 128: return attn_variant<128>(store_p);
    case 160: return attn_variant<160>(store_p);
    default: return attn_variant<192>(store_p);
  }
}

}  // namespace

// O[B*N, C] = merge_heads(softmax(scale * Q_h K_h^T) V_h); q: [B*N, C] bf16, kv: [B*Nk, 2C] bf16 (K columns [0, C), V
// columns [C, 2C)), heads = C / 64, o: [B*N, C] bf16, p_out: optional [B, heads, N, Nk] bf16 (nullptr: not stored).
// Needs Nk % 32 == 0 and Nk <= 192; all tensors contiguous and 16-byte aligned.
extern "C" int mvlt_sr_attention_fwd(const void* q_bf16, const void* kv_bf16, void* o_bf16, void* p_out_bf16, int B, int N,
                                     int Nk, int heads, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(q_bf16 && kv_bf16 && o_bf16, "sr_attention_fwd: null operand");
  MVLT_CHECK_ARG(B > 0 && N > 0 && heads > 0, "sr_attention_fwd: bad shape B=%d N=%d heads=%d", B, N, heads);
  MVLT_CHECK_ARG(Nk >= 32 && Nk <= NK_MAX && Nk % 32 == 0, "sr_attention_fwd: Nk=%d unsupported (multiple of 32, <= %d)", Nk, NK_MAX);
  MVLT_CHECK_ARG(((((uintptr_t)q_bf16) | ((uintptr_t)kv_bf16) | ((uintptr_t)o_bf16) | ((uintptr_t)p_out_bf16)) & 15) == 0,
                 "sr_attention_fwd: operands must be 16-byte aligned");
  const int C = heads * HD;
  AttnParams p;
  p.B = B; p.heads = heads; p.N = N; p.Nk = Nk; p.C = C;
  p.num_m = (N + BM - 1) / BM;
  const long long tiles = (long long)B * heads * p.num_m;
  MVLT_CHECK_ARG(tiles < (1ll << 31), "sr_attention_fwd: too many tiles");
  p.total_tiles = (int)tiles;
  p.a2 = scale * 1.4426950408889634f;

  CUtensorMap tmQ, tmKV, tmO, tmP;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)N, (uint64_t)B, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)N * C * 2, (uint64_t)B * N * C * 2};
    const uint32_t box[4] = {HD, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmQ, q_bf16, dims, str, box, 0, 0)) != 0) return rc;
    if ((rc = mvlt_tensor_map_4d(&tmO, o_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)2 * C, (uint64_t)Nk, (uint64_t)B, 1};
    const uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)Nk * C * 4, (uint64_t)B * Nk * C * 4};
    const uint32_t box[4] = {HD, (uint32_t)Nk, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmKV, kv_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  tmP = tmQ;
  if (p_out_bf16 != nullptr) {
    const uint64_t dims[4] = {(uint64_t)Nk, (uint64_t)N, (uint64_t)heads, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Nk * 2, (uint64_t)N * Nk * 2, (uint64_t)heads * N * Nk * 2};
    const uint32_t box[4] = {64, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmP, p_out_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }

  MVLT_CHECK_ARG(scale > 0.f, "sr_attention_fwd: the softmax scale must be positive");
  static const bool use_v1 = [] { const char* e = getenv("MVLT_ATTN_V1"); return e != nullptr && e[0] == '1'; }();
  AttnKernelFn fn = use_v1 ? (p_out_bf16 != nullptr ? sr_attention_fwd_kernel_v1<true> : sr_attention_fwd_kernel_v1<false>)
                           : pick_attn_kernel(Nk, p_out_bf16 != nullptr);
  {
    static std::mutex mu;
    static std::unordered_set<const void*> configured;
    std::lock_guard<std::mutex> g(mu);
    if (configured.insert(reinterpret_cast<const void*>(fn)).second) {
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_USED + 1024);
      cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
  }
  int grid = 2 * mvlt_num_sms();
  if (tiles < grid) grid = (int)tiles;
  // no alignment slack: the dynamic shared memory base is 1024-byte aligned (declared; the kernel traps otherwise), which
  // is what lets two CTAs (2 x 112 KB + 2 x 1 KB reserved) share one SM
  const size_t smem = SMEM_USED;
  mvlt_launch(fn, grid, THREADS, smem, stream, tmQ, tmKV, tmO, tmP, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}
