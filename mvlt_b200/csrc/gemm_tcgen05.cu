// Persistent, warp-specialised tcgen05 GEMM for sm_100a (B200).
//
//   warp 0      : TMA producer (cp.async.bulk.tensor.4d, SWIZZLE_128B boxes, mbarrier complete_tx)
//   warp 1      : TMEM owner + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, bf16 -> fp32)
//   warps 2..9  : epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> 128-bit global stores)
//
// Accumulators live in TMEM and are double-buffered (2 x BLOCK_N <= 512 columns) so the epilogue of tile
// i overlaps the mainloop of tile i+1. Tiles are scheduled round-robin over a grid of <= #SM CTAs.
// Operands may be K-major or MN-major (see gemm_desc.h); MN-major tiles are fetched as 64x64 swizzle
// atoms and described to the tensor core with the MN-major canonical layout (LBO = atom stride).
//
// Replaces every cuBLAS/cuDNN call the reference issues through nn.Linear / conv(k=s) / bmm:
// /root/reference/libs/pvlt.py:66-70,98,104,108-118,168 and libs/vl_heads.py:31,67.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>
#include "common.cuh"
#include "gemm_desc.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_THREADS = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 190 * 1024;  // pipeline stages; + 32 KB epilogue store staging + barriers <= 227 KB
constexpr int STAGING_BYTES = 8 * 4096;

struct KParams {
  int M, N, K;
  int block_n, stages;
  int num_m_blocks, num_n_blocks, num_k_blocks, kb_per_split, split_k;
  int batch1, batch2;
  int a_mn, b_mn;
  int a_b1, a_b2, b_b1, b_b2;  // 1 if the batch coordinate is used for that operand, 0 if broadcast
  void* D;
  void* D2;
  const float* bias;
  const __nv_bfloat16* aux;
  const float* residual;
  const float* rowscale;
  long long ldd, sD1, sD2;
  float alpha;
  int act, out_f32, atomic_add, rows_per_scale;
};

// ---- UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp, restated) -----------------------------
// smem matrix descriptor: [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) swizzle
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version for Blackwell
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: c=f32 (bit4), a=bf16 (bit7), b=bf16 (bit10), a_major bit15, b_major bit16,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_instr_desc(int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BLOCK_M >> 4) << 24;
  return d;
}

struct TileCoord {
  int b1, b2, m_blk, n_blk, split;
};
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
  TileCoord c;
  c.n_blk = t % p.num_n_blocks;
  t /= p.num_n_blocks;
  c.m_blk = t % p.num_m_blocks;
  t /= p.num_m_blocks;
  c.split = t % p.split_k;
  t /= p.split_k;
  c.b2 = t % p.batch2;
  c.b1 = t / p.batch2;
  return c;
}

// Fused epilogue for one thread's 32 consecutive output columns of one row. Every loop is fully unrolled so that
// v[] stays in registers (a dynamically indexed tail loop would demote it to local memory).
// ex[] holds this chunk's residual (32 fp32 words) or aux (16 words of bf16 pairs), loaded ahead of time by
// issue_extra() so that their DRAM latency overlaps the MMA wait / the previous chunk.
__device__ __forceinline__ bool chunk_vec_ok(const KParams& p, int col0, bool aligned) {
  return (col0 + 32 <= p.N) && aligned && ((col0 & 7) == 0);
}
__device__ __forceinline__ void issue_extra(const KParams& p, long long row_off, int col0, bool use, uint32_t (&ex)[32]) {
  if (!use) return;
  if (p.residual != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(p.residual + row_off + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 t = rp[j];
      ex[4 * j] = __float_as_uint(t.x); ex[4 * j + 1] = __float_as_uint(t.y);
      ex[4 * j + 2] = __float_as_uint(t.z); ex[4 * j + 3] = __float_as_uint(t.w);
    }
  } else if (p.aux != nullptr) {
    const uint4* ap = reinterpret_cast<const uint4*>(p.aux + row_off + col0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 t = ap[j];
      ex[4 * j] = t.x; ex[4 * j + 1] = t.y; ex[4 * j + 2] = t.z; ex[4 * j + 3] = t.w;
    }
  }
}

// Warp-cooperative, fully coalesced store of a 32-row x (P*16)-byte chunk: every lane holds one row in registers;
// the rows go through a swizzled (bank-conflict-free) per-warp smem tile so that each global store instruction
// writes whole contiguous row segments (64 B for bf16, 128 B for fp32) instead of 16 B slivers of 32 different rows.
template <int P>
__device__ __forceinline__ void staged_store(uint8_t* stage, const uint4 (&pieces)[P], uint8_t* gbase, long long ld_bytes,
                                             int lane, int rows_valid) {
  constexpr int W = P * 16;
  const int swz_w = (P == 8) ? (lane & 7) : ((lane >> 1) & 3);
  uint8_t* wrow = stage + lane * W;
#pragma unroll
  for (int q = 0; q < P; ++q) *reinterpret_cast<uint4*>(wrow + ((q ^ swz_w) << 4)) = pieces[q];
  __syncwarp();
  constexpr int RPI = 32 / P;  // rows covered by one store instruction
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const int row = q * RPI + lane / P, pc = lane % P;
    const int swz_r = (P == 8) ? (row & 7) : ((row >> 1) & 3);
    const uint4 val = *reinterpret_cast<const uint4*>(stage + row * W + ((pc ^ swz_r) << 4));
    if (row < rows_valid) *reinterpret_cast<uint4*>(gbase + (long long)row * ld_bytes + pc * 16) = val;
  }
  __syncwarp();
}

__device__ __forceinline__ void pack_bf16_row(const float (&v)[32], uint4 (&out)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    out[j].x = pack_bf16x2(v[8 * j], v[8 * j + 1]); out[j].y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    out[j].z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]); out[j].w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
  }
}

// Fused epilogue for one warp's 32 rows x 32 columns (one row per lane). Called warp-uniformly; every loop is fully
// unrolled so that v[] stays in registers.
template <bool kHasExtra>
__device__ __forceinline__ void epilogue_chunk(const KParams& p, const uint32_t (&r)[32], const uint32_t (&ex)[32],
                                               long long row_base_off, int lane, int rows_valid, int col0, float rs,
                                               bool aligned, uint8_t* stage) {
  const bool row_ok = lane < rows_valid;
  const long long row_off = row_base_off + (long long)lane * p.ldd;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
  const bool full = (col0 + 32 <= p.N);
  const bool vec_ok = full && aligned && ((col0 & 7) == 0);   // warp-uniform
  if (p.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
    }
  }
  if (!kHasExtra && (p.act == MVLT_ACT_GELU || p.act == MVLT_ACT_GELU_SAVE_GRAD)) {
    float d2v[32];
    if (p.act == MVLT_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { d2v[j] = v[j]; v[j] = gelu_fast(v[j]); }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) gelu_and_grad(v[j], v[j], d2v[j]);
    }
    if (p.D2 != nullptr) {
      __nv_bfloat16* d2base = reinterpret_cast<__nv_bfloat16*>(p.D2) + row_base_off + col0;
      if (vec_ok) {
        uint4 pk[4];
        pack_bf16_row(d2v, pk);
        staged_store<4>(stage, pk, reinterpret_cast<uint8_t*>(d2base), p.ldd * 2, lane, rows_valid);
      } else if (row_ok) {
        __nv_bfloat16* d2 = d2base + (long long)lane * p.ldd;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) d2[j] = __float2bfloat16(d2v[j]);
      }
    }
  } else if (kHasExtra && p.act == MVLT_ACT_MUL_AUX) {
    if (vec_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 f = unpack_bf16x2(ex[j]);
        v[2 * j] *= f.x; v[2 * j + 1] *= f.y;
      }
    } else if (row_ok) {
      const __nv_bfloat16* ax = p.aux + row_off + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] *= __bfloat162float(ax[j]);
    }
  } else if (kHasExtra && p.act == MVLT_ACT_DGELU) {
    if (vec_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 f = unpack_bf16x2(ex[j]);
        v[2 * j] *= dgelu_fast(f.x); v[2 * j + 1] *= dgelu_fast(f.y);
      }
    } else if (row_ok) {
      const __nv_bfloat16* ax = p.aux + row_off + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] *= dgelu_fast(__bfloat162float(ax[j]));
    }
  }
  if (kHasExtra && p.residual != nullptr) {
    if (vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(rs, v[j], __uint_as_float(ex[j]));
    } else if (row_ok) {
      const float* rp = p.residual + row_off + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) v[j] = rp[j] + rs * v[j];
    }
  } else if (p.rowscale != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rs;
  }
  if (p.atomic_add) {
    if (row_ok) {
      float* d = reinterpret_cast<float*>(p.D) + row_off + col0;
      if (vec_ok) {   // 128-bit vector reductions: 4x fewer L2 atomic operations for the split-K dW GEMMs
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + j), "f"(v[j]), "f"(v[j + 1]),
                       "f"(v[j + 2]), "f"(v[j + 3])
                       : "memory");
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) atomicAdd(d + j, v[j]);
      }
    }
  } else if (p.out_f32) {
    float* dbase = reinterpret_cast<float*>(p.D) + row_base_off + col0;
    if (vec_ok) {
      uint4 pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                           __float_as_uint(v[4 * j + 3]));
      staged_store<8>(stage, pk, reinterpret_cast<uint8_t*>(dbase), p.ldd * 4, lane, rows_valid);
    } else if (row_ok) {
      float* d = dbase + (long long)lane * p.ldd;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) d[j] = v[j];
    }
  } else {
    __nv_bfloat16* dbase = reinterpret_cast<__nv_bfloat16*>(p.D) + row_base_off + col0;
    if (vec_ok) {
      uint4 pk[4];
      pack_bf16_row(v, pk);
      staged_store<4>(stage, pk, reinterpret_cast<uint8_t*>(dbase), p.ldd * 2, lane, rows_valid);
    } else if (row_ok) {
      __nv_bfloat16* d = dbase + (long long)lane * p.ldd;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < p.N) d[j] = __float2bfloat16(v[j]);
    }
  }
}

template <bool kHasExtra>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_stage_bytes = p.block_n * BLOCK_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(full_bar) + 256;   // 8 x 4 KB per-warp store staging tiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one elected lane per warp of the owning epilogue group
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.batch1 * p.batch2 * p.split_k * p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int m0 = tc.m_blk * BLOCK_M, n0 = tc.n_blk * p.block_n;
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn) {
            tma_load_4d(sa, &tmA, &full_bar[stage], k0, m0, tc.b2 * p.a_b2, tc.b1 * p.a_b1);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_4d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, k0, tc.b2 * p.a_b2,
                          tc.b1 * p.a_b1);
          }
          if (!p.b_mn) {
            tma_load_4d(sb, &tmB, &full_bar[stage], k0, n0, tc.b2 * p.b_b2, tc.b1 * p.b_b1);
          } else {
            for (int j = 0; j < p.block_n / 64; ++j)
              tma_load_4d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, k0, tc.b2 * p.b_b2,
                          tc.b1 * p.b_b1);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(p.block_n, p.a_mn, p.b_mn);
      // K-major: 8-row groups are 1024 B apart (SBO); the single 128 B swizzle atom along K makes LBO unused.
      // MN-major: 64(mn) x 8(k) atoms; SBO = 1024 B between k-groups, LBO = 8192 B between 64-wide mn groups.
      const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int kb0 = tc.split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.block_n);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, 1024u);
            const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, 1024u);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    // Two groups of four warps (one warp per TMEM lane quarter). Group g drains accumulator buffer g, i.e. every
    // second tile of this CTA, so the (latency-bound) epilogues of consecutive tiles overlap each other as well as
    // the mainloop. Residual / aux operands of a chunk are fetched one chunk ahead (and, for the first chunk,
    // before waiting for the MMA).
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int group = (warp - 2) >> 2;     // accumulator buffer / tile parity owned by this warp
    constexpr bool has_extra = kHasExtra;
    uint32_t phase = 0;
    int local = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++local) {
      if ((local & 1) != group) continue;
      const TileCoord tc = decode_tile(p, t);
      const int row = tc.m_blk * BLOCK_M + quarter * 32 + lane;
      const int n0 = tc.n_blk * p.block_n;
      const bool row_ok = row < p.M;
      const int row_base = tc.m_blk * BLOCK_M + quarter * 32;
      const int rows_valid = min(32, p.M - row_base);          // <= 0 when the whole warp is past the M tail
      const long long batch_off = (long long)tc.b1 * p.sD1 + (long long)tc.b2 * p.sD2;
      const long long row_base_off = batch_off + (long long)row_base * p.ldd;
      const long long row_off = row_base_off + (long long)lane * p.ldd;
      float rs = 1.f;
      if (p.rowscale != nullptr && row_ok) rs = p.rowscale[row / p.rows_per_scale];
      const bool aligned = ((p.ldd & 7) == 0) && ((batch_off & 7) == 0);
      uint8_t* stage = stage_base + (warp - 2) * 4096;

      uint32_t exA[32], exB[32];
      if (kHasExtra) issue_extra(p, row_off, n0, row_ok && chunk_vec_ok(p, n0, aligned), exA);
      mbar_wait(&tfull_bar[group], phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(group * p.block_n);
      for (int c = 0; c < p.block_n; c += 64) {
        uint32_t r[32];
        {
          tmem_ld_32x32(taddr0 + (uint32_t)c, r);
          const int cn = n0 + c + 32;
          if (kHasExtra && c + 32 < p.block_n) issue_extra(p, row_off, cn, row_ok && chunk_vec_ok(p, cn, aligned), exB);
          tmem_ld_wait();
          const int col0 = n0 + c;
          if (rows_valid > 0 && col0 < p.N) epilogue_chunk<kHasExtra>(p, r, exA, row_base_off, lane, rows_valid, col0, rs, aligned, stage);
        }
        if (c + 32 < p.block_n) {
          tmem_ld_32x32(taddr0 + (uint32_t)(c + 32), r);
          const int cn = n0 + c + 64;
          if (kHasExtra && c + 64 < p.block_n) issue_extra(p, row_off, cn, row_ok && chunk_vec_ok(p, cn, aligned), exA);
          tmem_ld_wait();
          const int col0 = n0 + c + 32;
          if (rows_valid > 0 && col0 < p.N) epilogue_chunk<kHasExtra>(p, r, exB, row_base_off, lane, rows_valid, col0, rs, aligned, stage);
        }
      }
      // all of this warp's TMEM reads are complete (wait::ld above): release the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[group]);
      phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// Host side: tensor-map cache + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  uint64_t v[12];
  bool operator==(const MapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 12; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};

// Builds (or fetches) a 4-D bf16 tensor map: dims innermost-first, strides in BYTES for dims 1..3.
int get_tensor_map(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_b[3],
                   const uint32_t box[4]) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  key.v[0] = (uint64_t)base;
  for (int i = 0; i < 4; ++i) key.v[1 + i] = dims[i];
  for (int i = 0; i < 3; ++i) key.v[5 + i] = strides_b[i];
  for (int i = 0; i < 4; ++i) key.v[8 + i] = box[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    mvlt_set_error("cuTensorMapEncodeTiled not available from the driver");
    return MVLT_ERR_DRIVER;
  }
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_b[0], strides_b[1], strides_b[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mvlt_set_error("cuTensorMapEncodeTiled failed (%d): base=%p dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] "
                   "box=[%u,%u,%u,%u]",
                   (int)r, base, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)strides_b[0],
                   (unsigned long long)strides_b[1], (unsigned long long)strides_b[2], box[0], box[1], box[2],
                   box[3]);
    return MVLT_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

// operand = rows x K matrix (rows = M for A, N for B)
int make_operand_map(CUtensorMap* out, const void* base, int rows, int K, int mn_major, long long ld,
                     int batch1, int batch2, long long s1, long long s2, int box_rows, int* use_b1,
                     int* use_b2) {
  if (((uintptr_t)base & 15) != 0) {
    mvlt_set_error("gemm operand base %p not 16-byte aligned", base);
    return MVLT_ERR_ALIGN;
  }
  if ((ld * 2) % 16 != 0 || (s1 * 2) % 16 != 0 || (s2 * 2) % 16 != 0) {
    mvlt_set_error("gemm operand strides must be multiples of 8 elements (ld=%lld s1=%lld s2=%lld)", ld, s1, s2);
    return MVLT_ERR_ALIGN;
  }
  *use_b1 = (batch1 > 1 && s1 != 0) ? 1 : 0;
  *use_b2 = (batch2 > 1 && s2 != 0) ? 1 : 0;
  uint64_t dims[4], str[3];
  uint32_t box[4];
  if (!mn_major) {
    dims[0] = (uint64_t)K;
    dims[1] = (uint64_t)rows;
    box[0] = BLOCK_K;
    box[1] = (uint32_t)box_rows;
  } else {
    dims[0] = (uint64_t)rows;
    dims[1] = (uint64_t)K;
    box[0] = 64;
    box[1] = BLOCK_K;
  }
  str[0] = (uint64_t)ld * 2;
  dims[2] = *use_b2 ? (uint64_t)batch2 : 1;
  dims[3] = *use_b1 ? (uint64_t)batch1 : 1;
  const uint64_t dflt = ((dims[1] * str[0] + 15) / 16) * 16;
  str[1] = *use_b2 ? (uint64_t)s2 * 2 : dflt;
  str[2] = *use_b1 ? (uint64_t)s1 * 2 : dflt;
  box[2] = 1;
  box[3] = 1;
  return get_tensor_map(out, base, dims, str, box);
}

int pick_block_n(const mvlt_gemm_desc* g) {
  const int N = g->N;
  const int step = g->b_mn ? 64 : 32;
  if (g->block_n > 0) return g->block_n;
  if (N <= 256) return ((N + step - 1) / step) * step;
  // prefer the largest block that divides N exactly, else 256 (tail handled by TMA zero fill + predicates)
  for (int bn = 256; bn >= 128; bn -= step)
    if (N % bn == 0) return bn;
  if (!g->b_mn)
    for (int bn = 224; bn >= 96; bn -= 32)
      if (N % bn == 0) return bn;
  return g->b_mn ? 128 : 256;
}

}  // namespace

extern "C" int mvlt_gemm(const mvlt_gemm_desc* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(g != nullptr, "mvlt_gemm: null descriptor");
  MVLT_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0, "mvlt_gemm: bad shape M=%d N=%d K=%d", g->M, g->N, g->K);
  MVLT_CHECK_ARG(g->batch1 >= 1 && g->batch2 >= 1, "mvlt_gemm: bad batch %d x %d", g->batch1, g->batch2);
  MVLT_CHECK_ARG(g->A && g->B && g->D, "mvlt_gemm: null operand");
  MVLT_CHECK_ARG(!(g->atomic_add && !g->out_f32), "mvlt_gemm: atomic_add needs an fp32 output");
  MVLT_CHECK_ARG(!(g->split_k > 1 && !g->atomic_add), "mvlt_gemm: split_k > 1 needs atomic_add");
  MVLT_CHECK_ARG(!((g->act == MVLT_ACT_DGELU || g->act == MVLT_ACT_MUL_AUX) && g->aux == nullptr),
                 "mvlt_gemm: dgelu / mul_aux epilogue needs aux");
  MVLT_CHECK_ARG(!(g->act == MVLT_ACT_GELU_SAVE_GRAD && g->D2 == nullptr), "mvlt_gemm: gelu_save_grad needs D2");
  MVLT_CHECK_ARG(!((g->act == MVLT_ACT_GELU || g->act == MVLT_ACT_GELU_SAVE_GRAD) && (g->residual || g->aux)),
                 "mvlt_gemm: the GELU epilogues do not combine with residual / aux operands");
  MVLT_CHECK_ARG(!(g->rowscale && g->rows_per_scale <= 0), "mvlt_gemm: rowscale needs rows_per_scale");

  KParams p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.block_n = pick_block_n(g);
  MVLT_CHECK_ARG(p.block_n >= 32 && p.block_n <= 256 && p.block_n % (g->b_mn ? 64 : 32) == 0,
                 "mvlt_gemm: unsupported block_n %d", p.block_n);
  const int stage_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;
  p.stages = SMEM_BUDGET / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.num_m_blocks = (g->M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (g->N + p.block_n - 1) / p.block_n;
  p.num_k_blocks = (g->K + BLOCK_K - 1) / BLOCK_K;
  int split = g->split_k > 1 ? g->split_k : 1;
  if (split > p.num_k_blocks) split = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + split - 1) / split;
  p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.batch1 = g->batch1; p.batch2 = g->batch2;
  p.a_mn = g->a_mn ? 1 : 0; p.b_mn = g->b_mn ? 1 : 0;
  p.D = g->D; p.D2 = g->D2; p.bias = g->bias;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(g->aux);
  p.residual = g->residual; p.rowscale = g->rowscale;
  p.ldd = g->ldd; p.sD1 = g->sD1; p.sD2 = g->sD2;
  p.alpha = g->alpha;
  p.act = g->act; p.out_f32 = g->out_f32; p.atomic_add = g->atomic_add;
  p.rows_per_scale = g->rows_per_scale > 0 ? g->rows_per_scale : 1;

  CUtensorMap tmA, tmB;
  int rc = make_operand_map(&tmA, g->A, g->M, g->K, p.a_mn, g->lda, g->batch1, g->batch2, g->sA1, g->sA2, BLOCK_M,
                            &p.a_b1, &p.a_b2);
  if (rc) return rc;
  rc = make_operand_map(&tmB, g->B, g->N, g->K, p.b_mn, g->ldb, g->batch1, g->batch2, g->sB1, g->sB2, p.block_n,
                        &p.b_b1, &p.b_b2);
  if (rc) return rc;

  // > half of the SM's shared memory so two CTAs (each wanting all 512 TMEM columns) never share an SM
  size_t smem = (size_t)p.stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/ + STAGING_BYTES;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  const long long total_tiles =
      (long long)p.batch1 * p.batch2 * p.split_k * p.num_m_blocks * p.num_n_blocks;
  MVLT_CHECK_ARG(total_tiles < (1ll << 31), "mvlt_gemm: too many tiles");
  int grid = mvlt_num_sms();
  if (total_tiles < grid) grid = (int)total_tiles;
  if (p.residual != nullptr || p.aux != nullptr) gemm_tcgen05_kernel<true><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, p);
  else gemm_tcgen05_kernel<false><<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}
