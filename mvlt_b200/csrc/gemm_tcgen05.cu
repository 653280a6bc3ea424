// Persistent, warp-specialised tcgen05 GEMM for sm_100a (B200).
//
//   warp 0      : TMA producer (cp.async.bulk.tensor.4d, SWIZZLE_128B boxes, mbarrier complete_tx)
//   warp 1      : TMEM owner + single-thread tcgen05.mma issuer (UMMA 128 x BLOCK_N x 16, bf16 -> fp32)
//   warps 2, 3  : idle (they only pad warpgroup 0 so that setmaxnreg can hand its registers to the epilogue)
//   warps 4..19 : epilogue (tcgen05.ld 32x32b -> registers -> fused epilogue -> smem-staged coalesced global I/O)
//
// Accumulators live in TMEM and are multi-buffered (2 x BLOCK_N, or 4 x BLOCK_N when BLOCK_N <= 128, of the 512
// columns) so the epilogue of tile i overlaps the mainloop of tiles i+1.. . Tiles are scheduled round-robin over a
// grid of <= #SM CTAs. Most PVLT-tiny GEMMs have K = 64..512 and are bound by their epilogue / HBM traffic, not by
// the tensor pipe: every accumulator is therefore drained by EIGHT warps (4 TMEM lane quarters x 2 column halves),
// 16 epilogue warps = 4 per SM sub-partition, and the kernel is specialised per epilogue kind so that the inner
// loops carry no runtime dispatch.
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
#include <stdlib.h>
#include "common.cuh"
#include "gemm_desc.h"

extern "C" int mvlt_colsum(const void* x, int x_f32, long long rows, int C, long long ld, float* out, void* stream_);

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int NUM_EPI_WARPS = 16;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS = 32 * (FIRST_EPI_WARP + NUM_EPI_WARPS);  // warpgroup 0 (TMA, MMA, 2 idle) + 16 epilogue warps
// Register budget: 640 threads launch with 96 registers each (61440 = the CTA's pool; setmaxnreg can only move
// registers inside that pool, an over-subscribed .inc blocks forever). Warpgroup 0 shrinks to 32 and the four
// epilogue warpgroups grow to 112: 128 * 32 + 512 * 112 = 61440.
constexpr int LAUNCH_REGS = 96;
constexpr int PRODUCER_REGS = 32;
constexpr int EPILOGUE_REGS = 112;
static_assert(128 * PRODUCER_REGS + (NUM_THREADS - 128) * EPILOGUE_REGS <= NUM_THREADS * LAUNCH_REGS,
              "setmaxnreg budget exceeds the registers the CTA is launched with");
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 8;
constexpr int MAX_ACC = 4;
constexpr int SMEM_BUDGET = 160 * 1024;  // pipeline stages; + 1 KB ones tile + barriers + 64 KB epilogue staging <= 227 KB
constexpr int ONES_BYTES = 1024;         // 8 rows x 64 bf16 of 1.0: B operand of the row-sum MMA (both 8-row groups alias it)
constexpr int ONES_N = 16;
constexpr int STAGING_BYTES = NUM_EPI_WARPS * 4096;

// epilogue kinds the kernel is specialised on
enum { EPI_PLAIN = 0, EPI_GELU = 1, EPI_AUX = 2, EPI_RESID = 3, EPI_SOFTMAX = 4, EPI_SOFTMAX_BWD = 5 };

// n / d for 0 <= n < 2^31 with a host-computed magic number (no integer division in the tile loops)
struct FastDiv {
  uint32_t d, m, sh;
};
__device__ __forceinline__ void fast_divmod(const FastDiv& f, uint32_t n, uint32_t& q, uint32_t& r) {
  q = (f.d == 1u) ? n : (__umulhi(n, f.m) >> f.sh);
  r = n - q * f.d;
}

struct KParams {
  int M, N, K;
  int block_n, stages, num_acc;
  int num_m_blocks, num_n_blocks, num_k_blocks, kb_per_split, split_k;
  int batch1, batch2;
  FastDiv div_n, div_m, div_s, div_b2;
  int a_mn, b_mn;
  int a_b1, a_b2, b_b1, b_b2;  // 1 if the batch coordinate is used for that operand, 0 if broadcast
  void* D;
  void* D2;
  const float* bias;
  const __nv_bfloat16* aux;
  const float* residual;
  const float* rowscale;
  float* rowsum;
  long long ldd, sD1, sD2;
  float alpha;
  int act, out_f32, atomic_add, rows_per_scale;
  int pair;       // CTA-pair mode (cta_group::2): clusters of 2 CTAs share one 256-row UMMA; num_m_blocks counts row PAIRS
  int patch_store;   // fp32 D is the 5-D patch view of an NHWC tensor (gemm_desc.h): units leave through cp.async.bulk.tensor.5d
  int patch_rc;      // R * C: columns per kernel row of that view
  int tma_store;  // D (and D2) leave the staging tiles through TMA bulk stores (tmD / tmD2) instead of LDS + STG
  int prefetch;   // EPI_AUX only: two staging tiles per epilogue warp, the aux unit of the next tile is fetched with cp.async
  int warp_stage_bytes;   // staging bytes per epilogue warp: 4096, or 8192 (prefetch; fused LayerNorm with 2 units per warp)
  // fused LayerNorm of the output rows (EPI_RESID, fp32 out, N == block_n == 64 | 128): see gemm_desc.h
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  float ln_eps;
  // implicit 3x3 convolution operand (gemm_desc.h): tile/k-block index -> (b, y) pixel coordinates and (tap, c0)
  int conv_mode, conv_W, conv_C;
  FastDiv div_hw, div_w, div_cb, div_c;   // pixels / (H*W), / W; k-block / (C/64); column / C
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
__device__ __forceinline__ uint32_t make_instr_desc(int n, int a_mn, int b_mn, int m = BLOCK_M) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

struct TileCoord {
  int b1, b2, m_blk, n_blk, split;
};
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
  TileCoord c;
  uint32_t q = (uint32_t)t, r;
  fast_divmod(p.div_n, q, q, r);
  c.n_blk = (int)r;
  fast_divmod(p.div_m, q, q, r);
  c.m_blk = (int)r;
  fast_divmod(p.div_s, q, q, r);
  c.split = (int)r;
  fast_divmod(p.div_b2, q, q, r);
  c.b2 = (int)r;
  c.b1 = (int)q;
  return c;
}

// ---- epilogue building blocks -------------------------------------------------------------------------------------
// Every epilogue warp owns a 4 KB staging tile in shared memory: 32 rows x 128 B (P = 8 sixteen-byte pieces per row:
// 32 fp32 or 64 bf16 columns) or 32 rows x 64 B (P = 4: 32 bf16 columns). A lane reads/writes ITS row (one TMEM lane)
// as 16-byte pieces, XOR-swizzled by row so that both the row-per-lane side and the coalesced side are bank-conflict
// free. Global memory is only ever touched on the coalesced side: one instruction moves whole contiguous row
// segments (4 rows x 128 B), never 16-byte slivers of 32 different rows. On B200 the L1/L2 request rate, not bytes,
// bounds these thin-K GEMMs, so every operand of the epilogue (D, D2, residual, aux) goes through the tile.
//
// piece(row, q) lives at tile + row * 16P + ((q ^ swz(row)) << 4), swz = row & 7 (P = 8) or (row >> 1) & 3 (P = 4).
// All lane-dependent parts of these addresses are computed once per warp (StageAddr); the per-piece part is an
// immediate.
struct StageAddr {
  uint32_t w8;    // row-per-lane side, P = 8: tile + lane * 128 (128-byte aligned)
  uint32_t wx8;   // (lane & 7) << 4
  uint32_t f8e;   // coalesced side, P = 8, even q: tile + (lane >> 3) * 128 + (((lane & 7) ^ (lane >> 3)) << 4)
  uint32_t f8o;   // odd q: f8e ^ 64
  uint32_t w4;    // row-per-lane side, P = 4: tile + lane * 64 (64-byte aligned)
  uint32_t wx4;   // ((lane >> 1) & 3) << 4
  uint32_t f4;    // coalesced side, P = 4: tile + (lane >> 2) * 64 + (((lane & 3) ^ ((lane >> 3) & 3)) << 4)
};
__device__ __forceinline__ StageAddr make_stage_addr(uint32_t tile_s, int lane) {
  StageAddr s;
  s.w8 = tile_s + (uint32_t)lane * 128u;
  s.wx8 = (uint32_t)(lane & 7) << 4;
  s.f8e = tile_s + (uint32_t)(lane >> 3) * 128u + ((uint32_t)((lane & 7) ^ (lane >> 3)) << 4);
  s.f8o = s.f8e ^ 64u;
  s.w4 = tile_s + (uint32_t)lane * 64u;
  s.wx4 = (uint32_t)((lane >> 1) & 3) << 4;
  s.f4 = tile_s + (uint32_t)(lane >> 2) * 64u + ((uint32_t)((lane & 3) ^ ((lane >> 3) & 3)) << 4);
  return s;
}
template <int P>
__device__ __forceinline__ uint32_t own_piece(const StageAddr& s, int q) {   // piece q of this lane's row
  return (P == 8) ? (s.w8 | (((uint32_t)q << 4) ^ s.wx8)) : (s.w4 | (((uint32_t)q << 4) ^ s.wx4));
}
template <int P>
__device__ __forceinline__ uint32_t co_piece(const StageAddr& s, int q) {    // piece handled by this lane in coalesced pass q
  return (P == 8) ? (((q & 1) ? s.f8o : s.f8e) + (uint32_t)q * 512u) : (s.f4 + (uint32_t)q * 512u);
}
// this lane's global byte offset inside a 32-row unit for coalesced pass 0; pass q adds q * (32 / P) rows
template <int P>
__device__ __forceinline__ long long co_goff(int lane, long long ld_bytes) {
  return (P == 8) ? ((long long)(lane >> 3) * ld_bytes + (lane & 7) * 16) : ((long long)(lane >> 2) * ld_bytes + (lane & 3) * 16);
}

// staging tile -> global (plain stores, or vector fp32 reductions for the split-K dW GEMMs)
template <int P, bool kRed>
__device__ __forceinline__ void stage_flush(const StageAddr& s, uint8_t* gbase, long long ld_bytes, int lane, int rows_valid) {
  constexpr int RPI = 32 / P;  // rows covered by one instruction
  uint8_t* g = gbase + co_goff<P>(lane, ld_bytes);
  const int row0 = (P == 8) ? (lane >> 3) : (lane >> 2);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const uint4 val = ld_shared_v4(co_piece<P>(s, q));
    if (rows_valid == 32 || q * RPI + row0 < rows_valid) {
      uint8_t* gq = g + (long long)(q * RPI) * ld_bytes;
      if (kRed)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gq), "f"(__uint_as_float(val.x)),
                     "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)), "f"(__uint_as_float(val.w))
                     : "memory");
      else
        *reinterpret_cast<uint4*>(gq) = val;
    }
  }
  __syncwarp();
}
// global -> registers in the coalesced pattern (issue only), then staging tile -> this lane's row in registers
template <int P>
__device__ __forceinline__ void load_rows_issue(const uint8_t* gbase, long long ld_bytes, int lane, int rows_valid, uint4 (&t)[P]) {
  constexpr int RPI = 32 / P;
  const uint8_t* g = gbase + co_goff<P>(lane, ld_bytes);
  const int row0 = (P == 8) ? (lane >> 3) : (lane >> 2);
#pragma unroll
  for (int q = 0; q < P; ++q) {
    t[q] = make_uint4(0u, 0u, 0u, 0u);
    if (rows_valid == 32 || q * RPI + row0 < rows_valid)
      t[q] = *reinterpret_cast<const uint4*>(g + (long long)(q * RPI) * ld_bytes);
  }
}
template <int P>
__device__ __forceinline__ void stage_transpose(const StageAddr& s, const uint4 (&t)[P], uint32_t (&ex)[P * 4]) {
#pragma unroll
  for (int q = 0; q < P; ++q) st_shared_v4(co_piece<P>(s, q), t[q].x, t[q].y, t[q].z, t[q].w);
  __syncwarp();
#pragma unroll
  for (int q = 0; q < P; ++q) {
    const uint4 v = ld_shared_v4(own_piece<P>(s, q));
    ex[4 * q] = v.x; ex[4 * q + 1] = v.y; ex[4 * q + 2] = v.z; ex[4 * q + 3] = v.w;
  }
  __syncwarp();
}
template <int P>
__device__ __forceinline__ void load_rows_via_stage(const StageAddr& s, const uint8_t* gbase, long long ld_bytes, int lane,
                                                    int rows_valid, uint32_t (&ex)[P * 4]) {
  uint4 t[P];
  load_rows_issue<P>(gbase, ld_bytes, lane, rows_valid, t);
  stage_transpose<P>(s, t, ex);
}
__device__ __forceinline__ uint32_t pack_bf16x2_f2(f32x2_t v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return pack_bf16x2(lo, hi);
}
// 16 bf16 columns (8 packed pairs) of this lane's row -> pieces (2*i16, 2*i16+1) of its staging row
template <int P>
__device__ __forceinline__ void stage_put_bf16(const StageAddr& s, int i16, const uint32_t (&pk)[8]) {
  st_shared_v4(own_piece<P>(s, 2 * i16), pk[0], pk[1], pk[2], pk[3]);
  st_shared_v4(own_piece<P>(s, 2 * i16 + 1), pk[4], pk[5], pk[6], pk[7]);
}

// alpha * acc + bias for 16 columns of this lane's row, as 8 packed fp32 pairs (FFMA2)
template <bool kFull>
__device__ __forceinline__ void scale_bias16(const KParams& p, const uint32_t (&r)[16], int c0, f32x2_t (&v2)[8]) {
  const f32x2_t alpha2 = f2_splat(p.alpha);
  if (p.bias == nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v2[j] = f2_mul(f2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), alpha2);
  } else if (kFull) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 4 * j));
      v2[2 * j] = f2_fma(f2_pack(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), alpha2, f2_pack(b4.x, b4.y));
      v2[2 * j + 1] = f2_fma(f2_pack(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), alpha2, f2_pack(b4.z, b4.w));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // N tail: predicated scalar bias loads
      float b0 = 0.f, b1 = 0.f;
      if (c0 + 2 * j < p.N) b0 = __ldg(p.bias + c0 + 2 * j);
      if (c0 + 2 * j + 1 < p.N) b1 = __ldg(p.bias + c0 + 2 * j + 1);
      v2[j] = f2_fma(f2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), alpha2, f2_pack(b0, b1));
    }
  }
}
// multiply by aux (MUL_AUX) or by gelu'(aux) (DGELU); ax = 8 bf16 pairs
__device__ __forceinline__ void apply_aux16(int act, const uint32_t* ax, f32x2_t (&v2)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 f = unpack_bf16x2(ax[j]);
    f32x2_t m = f2_pack(f.x, f.y);
    if (act == MVLT_ACT_DGELU) {
      f32x2_t g;
      gelu_and_grad2(m, g, m);
    }
    v2[j] = f2_mul(v2[j], m);
  }
}

// Vector path of the fused epilogue for one "unit" of this warp: 32 rows x (32 fp32 | 64 bf16 | 32 bf16) columns,
// i.e. one staging tile with P pieces per row. All columns of the unit are inside N and 16-byte aligned.
// The TMEM accumulator is read 16 columns at a time to keep the register footprint small (no spills at 112
// registers: local memory has almost no L1 behind it in this kernel). Warp-uniform; loops fully unrolled.
// The staging tile of a 128-byte-row unit has exactly the SWIZZLE_128B layout of a [32 rows x 128 B] TMA box, so the
// unit leaves shared memory with ONE bulk tensor store issued by one lane (rows past M are clipped by the tensor map)
// instead of 8 x (LDS + predicated STG) per lane; the warp only waits until the TMA engine has read the tile.
struct StoreCtx {
  const CUtensorMap* tmD;     // 128-byte-row boxes (64 bf16 / 32 fp32 columns)
  const CUtensorMap* tmD2;
  const CUtensorMap* tmDh;    // 64-byte-row boxes (32 bf16 columns), SWIZZLE_64B == the P = 4 staging layout
  const CUtensorMap* tmD2h;
  int row, b2, b1;    // tile-row coordinate of this warp's first row and the batch coordinates
};
template <bool kAdd = false>
__device__ __forceinline__ void tma_flush(const CUtensorMap* tm, uint32_t tile_s, int col0, const StoreCtx& sc, int lane) {
  fence_proxy_async();          // this lane's st.shared writes -> visible to the async proxy
  __syncwarp();
  if (lane == 0) {
    if (kAdd) tma_reduce_add_4d(tm, tile_s, col0, sc.row, sc.b2, sc.b1);   // split-K / gradient accumulation (fp32)
    else tma_store_4d(tm, tile_s, col0, sc.row, sc.b2, sc.b1);
    tma_store_commit();
    tma_store_wait_read();
  }
  __syncwarp();
}

// one 32-row x 32-fp32-column unit of the product into the patch view of dX: rows = 4 oy x 8 ox of one image, columns = 32
// consecutive (kx, c) of one kernel row ky
__device__ __forceinline__ void tma_flush_patch(const KParams& p, const CUtensorMap* tm, uint32_t tile_s, int col0, int row, int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    const int ky = col0 / p.patch_rc;
    tma_store_5d(tm, tile_s, col0 - ky * p.patch_rc, 0, ky, (row & 63) >> 3, row >> 6);
    tma_store_commit();
    tma_store_wait_read();
  }
  __syncwarp();
}

template <int kEpi, bool kOutF32, int P>
__device__ __forceinline__ void epilogue_unit_compute(const KParams& p, uint32_t taddr, long long row_base_off, int lane,
                                                      int rows_valid, int col0, float rs, const StageAddr& s,
                                                      const uint32_t (&ex)[P * 4], const StoreCtx& sc) {
  constexpr int COLS = kOutF32 ? 32 : P * 8;
  constexpr int N16 = COLS / 16;
  static_assert(!kOutF32 || P == 8, "fp32 units are 32 columns = 128-byte rows");
  const long long ld_bytes = p.ldd * (kOutF32 ? 4 : 2);
  uint32_t d2pk[P * 4];   // second bf16 output of the GELU epilogues, flushed after D (dead otherwise)
#pragma unroll
  for (int i = 0; i < N16; ++i) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + (uint32_t)(16 * i), r);
    tmem_ld_wait();
    f32x2_t v2[8];
    scale_bias16<true>(p, r, col0 + 16 * i, v2);
    if constexpr (kEpi == EPI_GELU) {
      if (p.act == MVLT_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { d2pk[8 * i + j] = pack_bf16x2_f2(v2[j]); v2[j] = gelu2(v2[j]); }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f32x2_t dg;
          gelu_and_grad2(v2[j], v2[j], dg);
          d2pk[8 * i + j] = pack_bf16x2_f2(dg);
        }
      }
    } else if constexpr (kEpi == EPI_PLAIN && kOutF32) {
      if (p.act == MVLT_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v2[j] = gelu2(v2[j]);
      }
    }
    if constexpr (kEpi == EPI_AUX) apply_aux16(p.act, &ex[8 * i], v2);
    if constexpr (kEpi == EPI_RESID) {
      const f32x2_t rs2 = f2_splat(rs);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v2[j] = f2_fma(rs2, v2[j], f2_pack(__uint_as_float(ex[16 * i + 2 * j]), __uint_as_float(ex[16 * i + 2 * j + 1])));
    } else if (p.rowscale != nullptr) {
      const f32x2_t rs2 = f2_splat(rs);
#pragma unroll
      for (int j = 0; j < 8; ++j) v2[j] = f2_mul(v2[j], rs2);
    }
    if constexpr (kOutF32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a, b, c, d;
        f2_unpack(v2[2 * j], a, b);
        f2_unpack(v2[2 * j + 1], c, d);
        st_shared_v4(own_piece<8>(s, 4 * i + j), __float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
      }
    } else {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2_f2(v2[j]);
      stage_put_bf16<P>(s, i, pk);
    }
  }
  if constexpr (kOutF32) {
    uint8_t* g = reinterpret_cast<uint8_t*>(reinterpret_cast<float*>(p.D) + row_base_off + col0);
    if (kEpi == EPI_PLAIN && p.patch_store) {
      tma_flush_patch(p, sc.tmD, s.w8 - (uint32_t)lane * 128u, col0, sc.row, lane);
    } else if (kEpi == EPI_PLAIN && p.atomic_add) {
      if (p.tma_store) tma_flush<true>(sc.tmD, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
      else stage_flush<8, true>(s, g, ld_bytes, lane, rows_valid);
    } else if (p.tma_store) tma_flush(sc.tmD, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
    else stage_flush<8, false>(s, g, ld_bytes, lane, rows_valid);
  } else {
    if (p.tma_store) tma_flush(P == 8 ? sc.tmD : sc.tmDh, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
    else
    stage_flush<P, false>(s, reinterpret_cast<uint8_t*>(reinterpret_cast<__nv_bfloat16*>(p.D) + row_base_off + col0), ld_bytes,
                          lane, rows_valid);
    if constexpr (kEpi == EPI_GELU) {
      if (p.D2 != nullptr) {
#pragma unroll
        for (int i = 0; i < N16; ++i) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = d2pk[8 * i + j];
          stage_put_bf16<P>(s, i, pk);
        }
        if (p.tma_store) tma_flush(P == 8 ? sc.tmD2 : sc.tmD2h, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
        else
        stage_flush<P, false>(s, reinterpret_cast<uint8_t*>(reinterpret_cast<__nv_bfloat16*>(p.D2) + row_base_off + col0),
                              ld_bytes, lane, rows_valid);
      }
    }
  }
}

template <int kEpi, bool kOutF32, int P>
__device__ __forceinline__ void epilogue_unit_vec(const KParams& p, uint32_t taddr, long long row_base_off, int lane,
                                                  int rows_valid, int col0, float rs, const StageAddr& s, const StoreCtx& sc) {
  uint32_t ex[P * 4];
  if constexpr (kEpi == EPI_RESID || kEpi == EPI_AUX) {   // residual (fp32 out) / aux (bf16 out): geometry of the output unit
    const long long ld_bytes = p.ldd * (kOutF32 ? 4 : 2);
    const uint8_t* g = (kEpi == EPI_RESID) ? reinterpret_cast<const uint8_t*>(p.residual + row_base_off + col0)
                                           : reinterpret_cast<const uint8_t*>(p.aux + row_base_off + col0);
    load_rows_via_stage<P>(s, g, ld_bytes, lane, rows_valid, ex);
  }
  epilogue_unit_compute<kEpi, kOutF32, P>(p, taddr, row_base_off, lane, rows_valid, col0, rs, s, ex, sc);
}

// Asynchronous (register-free) fetch of a 32-row x 128-byte operand unit into a staging tile: cp.async in the
// coalesced pattern, zero-filled past rows_valid. Completion: cp.async.wait_group + __syncwarp by the caller.
__device__ __forceinline__ void prefetch_rows_async(const StageAddr& s, const uint8_t* gbase, long long ld_bytes, int lane,
                                                    int rows_valid) {
  const uint8_t* g = gbase + co_goff<8>(lane, ld_bytes);
  const int row0 = lane >> 3;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const bool ok = q * 4 + row0 < rows_valid;
    const uint8_t* src = ok ? g + (long long)(q * 4) * ld_bytes : gbase;   // never form an out-of-range address
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(co_piece<8>(s, q)), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Tail path (N not a multiple of 32 / unaligned rows): one 32-column chunk with per-element predicates and
// row-per-lane global accesses. Only the last column block of odd-sized problems (vocabulary 30522) gets here.
__device__ __noinline__ void epilogue_chunk_tail(const KParams& p, uint32_t taddr, long long row_base_off, int lane,
                                                 int rows_valid, int col0, float rs) {
  const bool row_ok = lane < rows_valid;
  const long long row_off = row_base_off + (long long)lane * p.ldd;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const int c0 = col0 + 16 * h;
    uint32_t r[16];
    tmem_ld_32x16(taddr + (uint32_t)(16 * h), r);
    tmem_ld_wait();
    f32x2_t v2[8];
    scale_bias16<false>(p, r, c0, v2);
    if (p.act == MVLT_ACT_GELU || p.act == MVLT_ACT_GELU_SAVE_GRAD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f32x2_t g, dg;
        gelu_and_grad2(v2[j], g, dg);
        const uint32_t second = pack_bf16x2_f2(p.act == MVLT_ACT_GELU ? v2[j] : dg);
        v2[j] = g;
        if (p.D2 != nullptr && row_ok) {
          __nv_bfloat16* d2 = reinterpret_cast<__nv_bfloat16*>(p.D2) + row_off + c0;
          const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(&second);
          if (c0 + 2 * j < p.N) d2[2 * j] = hv.x;
          if (c0 + 2 * j + 1 < p.N) d2[2 * j + 1] = hv.y;
        }
      }
    }
    if (p.aux != nullptr) {
      uint32_t ax[8];
      const __nv_bfloat16* ap = p.aux + row_off + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a0 = 0.f, a1 = 0.f;
        if (row_ok && c0 + 2 * j < p.N) a0 = __bfloat162float(ap[2 * j]);
        if (row_ok && c0 + 2 * j + 1 < p.N) a1 = __bfloat162float(ap[2 * j + 1]);
        ax[j] = pack_bf16x2(a0, a1);
      }
      apply_aux16(p.act, ax, v2);
    }
    float v[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) f2_unpack(v2[j], v[2 * j], v[2 * j + 1]);
    if (p.residual != nullptr) {
      const float* rp = p.residual + row_off + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (row_ok && c0 + j < p.N) v[j] = fmaf(rs, v[j], rp[j]);
    } else if (p.rowscale != nullptr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= rs;
    }
    if (row_ok) {
      if (p.out_f32) {
        float* d = reinterpret_cast<float*>(p.D) + row_off + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < p.N) {
            if (p.atomic_add) atomicAdd(d + j, v[j]);
            else d[j] = v[j];
          }
        }
      } else {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(p.D) + row_off + c0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.N) d[j] = __float2bfloat16(v[j]);
      }
    }
  }
}

// Exchange of per-row partial results between the two column-half warps of one (accumulator, lane-quarter) pair:
// each writes its value(s) to the head of its own staging tile, the pair meets on a named barrier, each reads the
// partner's, and a second barrier protects the staging tile before it is reused.
__device__ __forceinline__ float2 pair_exchange(uint32_t my_stage_s, uint32_t partner_stage_s, int bar_id, int lane, float2 mine) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(my_stage_s + lane * 8), "f"(mine.x), "f"(mine.y) : "memory");
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  float2 other;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(other.x), "=f"(other.y) : "r"(partner_stage_s + lane * 8) : "memory");
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
  return other;
}

// ---- fused row softmax (forward): pass 2 for one unit of P*8 bf16 columns: P = exp2(a2 * acc - mx) * inv
template <int P>
__device__ __forceinline__ void softmax_unit(const KParams& p, uint32_t taddr, long long row_base_off, int lane, int rows_valid,
                                             int col0, float a2, float mx, float inv, const StageAddr& s, const StoreCtx& sc) {
#pragma unroll
  for (int i = 0; i < P / 2; ++i) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + (uint32_t)(16 * i), r);
    tmem_ld_wait();
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      pk[j] = pack_bf16x2(ex2_approx(fmaf(__uint_as_float(r[2 * j]), a2, -mx)) * inv,
                          ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), a2, -mx)) * inv);
    stage_put_bf16<P>(s, i, pk);
  }
  if (p.tma_store) tma_flush(P == 8 ? sc.tmD : sc.tmDh, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
  else
  stage_flush<P, false>(s, reinterpret_cast<uint8_t*>(reinterpret_cast<__nv_bfloat16*>(p.D) + row_base_off + col0), p.ldd * 2,
                        lane, rows_valid);
}
// ---- fused softmax backward for one unit: kPass = 0 accumulates dot += sum_n P * dP; kPass = 1 writes
// dS = alpha * P * (dP - dot). P (aux) is fetched through the staging tile both times (DRAM, then L2).
template <int P, int kPass>
__device__ __forceinline__ float softmax_bwd_unit(const KParams& p, uint32_t taddr, long long row_base_off, int lane,
                                                  int rows_valid, int col0, float dot, const StageAddr& s, const StoreCtx& sc) {
  uint32_t ex[P * 4];
  load_rows_via_stage<P>(s, reinterpret_cast<const uint8_t*>(p.aux + row_base_off + col0), p.ldd * 2, lane, rows_valid, ex);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < P / 2; ++i) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + (uint32_t)(16 * i), r);
    tmem_ld_wait();
    if (kPass == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = unpack_bf16x2(ex[8 * i + j]);
        acc = fmaf(f.x, __uint_as_float(r[2 * j]), acc);
        acc = fmaf(f.y, __uint_as_float(r[2 * j + 1]), acc);
      }
    } else {
      uint32_t pk[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = unpack_bf16x2(ex[8 * i + j]);
        pk[j] = pack_bf16x2(p.alpha * f.x * (__uint_as_float(r[2 * j]) - dot),
                            p.alpha * f.y * (__uint_as_float(r[2 * j + 1]) - dot));
      }
      stage_put_bf16<P>(s, i, pk);
    }
  }
  if (kPass == 1) {
    if (p.tma_store) tma_flush(P == 8 ? sc.tmD : sc.tmDh, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
    else
      stage_flush<P, false>(s, reinterpret_cast<uint8_t*>(reinterpret_cast<__nv_bfloat16*>(p.D) + row_base_off + col0), p.ldd * 2,
                            lane, rows_valid);
  }
  return acc;
}

// ---- fused LayerNorm of the output rows (residual epilogue, fp32 out, the whole row in this tile) ------------------------------
// Pass 1, one unit (32 fp32 columns) of this lane's row: v = residual + rs * (alpha * acc + bias) -> D through the staging tile
// (which KEEPS v afterwards: one tile per unit of the warp), row sum and sum of squares accumulated on the way.
__device__ __forceinline__ void resid_ln_pass1(const KParams& p, uint32_t taddr, long long row_base_off, int lane, int rows_valid,
                                               int col0, float rs, const StageAddr& s, const StoreCtx& sc, float& sum, float& sq) {
  uint32_t ex[32];
  load_rows_via_stage<8>(s, reinterpret_cast<const uint8_t*>(p.residual + row_base_off + col0), p.ldd * 4, lane, rows_valid, ex);
  const f32x2_t rs2 = f2_splat(rs);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + (uint32_t)(16 * i), r);
    tmem_ld_wait();
    f32x2_t v2[8];
    scale_bias16<true>(p, r, col0 + 16 * i, v2);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      v2[j] = f2_fma(rs2, v2[j], f2_pack(__uint_as_float(ex[16 * i + 2 * j]), __uint_as_float(ex[16 * i + 2 * j + 1])));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a, b, c, d;
      f2_unpack(v2[2 * j], a, b);
      f2_unpack(v2[2 * j + 1], c, d);
      sum += (a + b) + (c + d);
      sq = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, sq))));
      st_shared_v4(own_piece<8>(s, 4 * i + j), __float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
    }
  }
  tma_flush(sc.tmD, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);   // returns once the TMA engine has READ the tile
}
__device__ __forceinline__ float4 ld_nc_v4_ordered(const float* ptr) {
  float4 r;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr));
  return r;
}
// Pass 2, same unit: v back from the staging tile, normalised, bf16 -> the tile again in the 64-byte-row layout -> D2.
__device__ __forceinline__ void resid_ln_pass2(const KParams& p, int lane, int col0, float mean, float rstd, const StageAddr& s,
                                               const StoreCtx& sc) {
  float v[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 t = ld_shared_v4(own_piece<8>(s, q));
    v[4 * q] = __uint_as_float(t.x); v[4 * q + 1] = __uint_as_float(t.y);
    v[4 * q + 2] = __uint_as_float(t.z); v[4 * q + 3] = __uint_as_float(t.w);
  }
  __syncwarp();   // every lane holds its fp32 row before the tile is rewritten as 32 rows x 64 bytes
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // (volatile loads: they stay behind the previous half's staging stores instead of all 64 values being hoisted and spilled)
      const float4 g4 = ld_nc_v4_ordered(p.ln_gamma + col0 + 16 * h + 4 * j);
      const float4 b4 = ld_nc_v4_ordered(p.ln_beta + col0 + 16 * h + 4 * j);
      const float* vv = &v[16 * h + 4 * j];
      pk[2 * j] = pack_bf16x2(fmaf((vv[0] - mean) * rstd, g4.x, b4.x), fmaf((vv[1] - mean) * rstd, g4.y, b4.y));
      pk[2 * j + 1] = pack_bf16x2(fmaf((vv[2] - mean) * rstd, g4.z, b4.z), fmaf((vv[3] - mean) * rstd, g4.w, b4.w));
    }
    stage_put_bf16<4>(s, h, pk);
  }
  tma_flush(sc.tmD2h, s.w8 - (uint32_t)lane * 128u, col0, sc, lane);
}

// kPair: CTA-pair build (cta_group::2). A kernel that contains cta_group::2 instructions can only be launched as clusters
// of two, so the pair path is a separate instantiation and the single-CTA build carries none of its code or registers.
template <int kEpi, bool kOutF32, bool kPair = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
                    const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmD2h,
                    const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();   // the next kernel of the stream may be scheduled behind this grid's tail (common.cuh)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA-pair mode: the two CTAs of a cluster (ranks 0 = leader, 1) each stage their own 128 A rows and HALF of the B tile;
  // the leader issues tcgen05.mma.cta_group::2 (M = 256), each CTA's TMEM receives its 128 rows x block_n accumulator
  constexpr int pair = kPair ? 1 : 0;
  constexpr int mb_mul = kPair ? 2 : 1;
  // cluster (or CTA) index / count of the persistent grid and this CTA's rank in its pair: re-materialised at every use
  // (special registers / constant bank) instead of living in registers through the register-starved epilogue
#define MVLT_CID (kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)
#define MVLT_NCL (kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x)
#define MVLT_RANK (kPair ? cluster_ctarank() : 0u)
#define MVLT_BROWS (kPair ? (p.block_n >> 1) : p.block_n)   /* B rows (or MN-major columns) staged by this CTA */
  const int b_stage_bytes = MVLT_BROWS * BLOCK_K * 2;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;

  uint8_t* ones_tile = smem + (size_t)p.stages * stage_bytes;   // 1024-byte aligned (stage sizes are multiples of 4 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones_tile + ONES_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + MAX_ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + MAX_ACC);
  uint8_t* stage_base = ones_tile + 2048;   // 16 x 4 KB per-warp staging tiles, 1024-byte aligned (TMA 128B-swizzle atoms)

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < MAX_ACC; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], (NUM_EPI_WARPS / 2) * mb_mul);  // one elected lane per warp of the owning epilogue group (both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) {
      tma_prefetch_desc(&tmD);
      if (p.D2 != nullptr) tma_prefetch_desc(&tmD2);
    }
  }
  if (p.rowsum != nullptr) {   // bf16 1.0 everywhere: layout / swizzle of this operand tile is irrelevant
    if (threadIdx.x < ONES_BYTES / 4) reinterpret_cast<uint32_t*>(ones_tile)[threadIdx.x] = 0x3F803F80u;
    fence_proxy_async();       // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers are initialised before anything of this CTA can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // barrier / TMEM set-up above overlapped the previous kernel; global memory is only touched below

  const int total_tiles = p.batch1 * p.batch2 * p.split_k * p.num_m_blocks * p.num_n_blocks;
  const int acc_mask = p.num_acc - 1;                 // num_acc is 2 or 4
  const int acc_shift = (p.num_acc == 4) ? 2 : 1;

  if (warp < FIRST_EPI_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = MVLT_CID; t < total_tiles; t += MVLT_NCL) {
          const TileCoord tc = decode_tile(p, t);
          const int m0 = (tc.m_blk * mb_mul + (int)MVLT_RANK) * BLOCK_M, n0 = tc.n_blk * p.block_n;
          const int kb0 = tc.split * p.kb_per_split;
          const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            uint8_t* sb = sa + A_STAGE_BYTES;
            const int k0 = kb * BLOCK_K;
            if constexpr (kPair) {
              // both CTAs' bytes complete on the LEADER's full barrier, which the leader arms for the two halves
              const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0u);
              if (MVLT_RANK == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)stage_bytes);
              if (!p.a_mn) {
                tma_load_4d_2sm(sa, &tmA, lbar, k0, m0, tc.b2 * p.a_b2, tc.b1 * p.a_b1);
              } else {
#pragma unroll
                for (int j = 0; j < BLOCK_M / 64; ++j)
                  tma_load_4d_2sm(sa + j * 8192, &tmA, lbar, m0 + j * 64, k0, tc.b2 * p.a_b2, tc.b1 * p.a_b1);
              }
              const int nb0 = n0 + (int)MVLT_RANK * MVLT_BROWS;     // this CTA's half of the B tile
              if (!p.b_mn) {
                tma_load_4d_2sm(sb, &tmB, lbar, k0, nb0, tc.b2 * p.b_b2, tc.b1 * p.b_b1);
              } else {
                for (int j = 0; j < MVLT_BROWS / 64; ++j)
                  tma_load_4d_2sm(sb + j * 8192, &tmB, lbar, nb0 + j * 64, k0, tc.b2 * p.b_b2, tc.b1 * p.b_b1);
              }
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
              continue;
            }
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
            if (p.conv_mode == MVLT_CONV_PATCH_A) {
              // A tile = the 2 x 64 output pixels of two images x 64 elements of kernel row ky = kb / (R*C/64): one 5-D box
              uint32_t ky, cb;
              fast_divmod(p.div_cb, (uint32_t)kb, ky, cb);
              tma_load_5d(sa, &tmA, &full_bar[stage], (int)cb * 64, 0, (int)ky, 0, m0 >> 6);
            } else if (p.conv_mode == MVLT_CONV_A) {
              // A tile = 128 consecutive pixels x 64 channels of tap (kb / (C/64)): one shifted NHWC box
              uint32_t tap, cb, b0, rem, y0, xr;
              fast_divmod(p.div_cb, (uint32_t)kb, tap, cb);
              fast_divmod(p.div_hw, (uint32_t)m0, b0, rem);
              fast_divmod(p.div_w, rem, y0, xr);
              const int ty = (int)tap / 3;
              tma_load_4d(sa, &tmA, &full_bar[stage], (int)cb * 64, (int)tap - 3 * ty - 1, (int)y0 + ty - 1, (int)b0);
            } else if (!p.a_mn) {
              tma_load_4d(sa, &tmA, &full_bar[stage], k0, m0, tc.b2 * p.a_b2, tc.b1 * p.a_b1);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_4d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, k0, tc.b2 * p.a_b2, tc.b1 * p.a_b1);
            }
            if (p.conv_mode == MVLT_CONV_BT) {
              // B tile (MN-major) = 64 consecutive pixels (k) x 64 channels of tap (n / C) per 64-column atom
              uint32_t b0, rem, y0, xr;
              fast_divmod(p.div_hw, (uint32_t)k0, b0, rem);
              fast_divmod(p.div_w, rem, y0, xr);
              for (int j = 0; j < p.block_n / 64; ++j) {
                uint32_t tap, c0;
                fast_divmod(p.div_c, (uint32_t)(n0 + j * 64), tap, c0);
                const int ty = (int)tap / 3;
                if (tap < 9u)
                  tma_load_4d(sb + j * 8192, &tmB, &full_bar[stage], (int)c0, (int)tap - 3 * ty - 1, (int)y0 + ty - 1, (int)b0);
                else   // columns past 9*C (N tail of the last block): any in-bounds-free box = zeros, keeps the tx count
                  tma_load_4d(sb + j * 8192, &tmB, &full_bar[stage], 0, 0, 0, 0x40000000);
              }
            } else if (!p.b_mn) {
              tma_load_4d(sb, &tmB, &full_bar[stage], k0, n0, tc.b2 * p.b_b2, tc.b1 * p.b_b1);
            } else {
              for (int j = 0; j < p.block_n / 64; ++j)
                tma_load_4d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, k0, tc.b2 * p.b_b2, tc.b1 * p.b_b1);
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
      if (lane == 0 && MVLT_RANK == 0) {     // (pair mode: the leader CTA issues for both)
        const uint32_t idesc = make_instr_desc(p.block_n, p.a_mn, p.b_mn, BLOCK_M * mb_mul);
        const uint32_t idesc_ones = make_instr_desc(ONES_N, p.a_mn, 0);
        const uint64_t ones_desc = make_smem_desc(smem_u32(ones_tile), 0u, 0u);   // SBO = 0: rows 8..15 re-read rows 0..7
        // K-major: 8-row groups are 1024 B apart (SBO); the single 128 B swizzle atom along K makes LBO unused.
        // MN-major: 64(mn) x 8(k) atoms; SBO = 1024 B between k-groups, LBO = 8192 B between 64-wide mn groups.
        const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
        const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
        int stage = 0;
        uint32_t phase = 0;
        int local = 0;
        for (int t = MVLT_CID; t < total_tiles; t += MVLT_NCL, ++local) {
          const TileCoord tc = decode_tile(p, t);
          const int kb0 = tc.split * p.kb_per_split;
          const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
          const int acc = local & acc_mask;
          const uint32_t acc_phase = (uint32_t)(local >> acc_shift) & 1u;
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.block_n);
          const bool do_rowsum = (p.rowsum != nullptr) && tc.n_blk == 0;
          const uint32_t tmem_ones = tmem_base + (uint32_t)(p.num_acc * p.block_n + acc * ONES_N);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t adesc = make_smem_desc(sa + k * a_kstep, a_lbo, 1024u);
              const uint64_t bdesc = make_smem_desc(sb + k * b_kstep, b_lbo, 1024u);
              if constexpr (kPair) umma_bf16_2sm(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              if (do_rowsum)   // row sums of the A tile: 128 x 16 x 16 MMA against the ones tile
                umma_bf16(tmem_ones, adesc, ones_desc, idesc_ones, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if constexpr (kPair) umma_commit_2sm(&empty_bar[stage]);   // frees the slot in BOTH CTAs
            else umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if constexpr (kPair) umma_commit_2sm(&tfull_bar[acc]);   // accumulator halves complete -> the epilogue warps of both CTAs
          else umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPILOGUE_REGS));
    // ===================== epilogue warps =====================
    // 16 warps = 2 tile parities x 4 TMEM lane quarters x 2 column halves. The warps of parity g drain every second
    // tile of this CTA, so the epilogues of consecutive tiles overlap each other as well as the mainloop. A warp
    // walks its column range in "units" of one staging tile: 64 bf16 / 32 fp32 columns (128-byte row segments), a
    // 32-column bf16 unit for an odd remainder, and a predicated tail path for ragged N.
    const int ew = warp - FIRST_EPI_WARP;
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int group = (ew >> 2) & 1;       // tile parity owned by this warp
    const int half = ew >> 3;              // column half of the tile
    const int nchunks = p.block_n >> 5;    // 32-column chunks per tile
    const int split = (nchunks + 1) >> 1;
    const int c_begin = half ? split : 0;
    const int c_end = half ? nchunks : split;
    const uint32_t stage = smem_u32(stage_base) + (uint32_t)ew * (uint32_t)p.warp_stage_bytes;
    const uint32_t partner_stage = smem_u32(stage_base) + (uint32_t)(ew ^ 8) * 4096u;
    const StageAddr sa = make_stage_addr(stage, lane);
    const int pair_bar = 1 + group * 4 + quarter;   // named barriers 1..8 (0 is __syncthreads)
    const bool ld_aligned = (p.ldd & 7) == 0;
    bool total_tiles_done = false;
    if constexpr (kEpi == EPI_AUX) {
      if (p.prefetch) {
        // ---- thin-K aux epilogues (dH = (dY W2) * gelu'): block_n = 128, so every warp owns exactly one 64-column unit
        // per tile. The unit's aux operand is the only DRAM read on the warp's critical path: it is fetched with
        // cp.async into a second staging tile one tile AHEAD, while the current unit is computed and stored.
        const StageAddr se = make_stage_addr(stage + 4096u, lane);
        const long long ld_bytes = p.ldd * 2;
        long long rbo = 0, rbo_n = 0;
        int rv = 0, rv_n = 0, c0 = 0, c0_n = 0, rb = 0, rb_n = 0, tb1 = 0, tb2 = 0, tb1_n = 0, tb2_n = 0;
        auto coords = [&](int tt, long long& o, int& v, int& c, int& r, int& cb1, int& cb2) {
          const TileCoord tc = decode_tile(p, tt);
          cb1 = tc.b1;
          cb2 = tc.b2;
          r = (tc.m_blk * mb_mul + (int)MVLT_RANK) * BLOCK_M + quarter * 32;
          v = min(32, p.M - r);
          o = (long long)tc.b1 * p.sD1 + (long long)tc.b2 * p.sD2 + (long long)r * p.ldd;
          c = tc.n_blk * p.block_n + half * 64;
        };
        int t = MVLT_CID + group * MVLT_NCL;
        if (t < total_tiles) {
          coords(t, rbo, rv, c0, rb, tb1, tb2);
          prefetch_rows_async(se, reinterpret_cast<const uint8_t*>(p.aux + rbo + c0), ld_bytes, lane, rv);
        }
        for (int local = group; t < total_tiles; t += 2 * MVLT_NCL, local += 2) {
          const int tn = t + 2 * MVLT_NCL;
          if (tn < total_tiles) coords(tn, rbo_n, rv_n, c0_n, rb_n, tb1_n, tb2_n);
          float rs = 1.f;
          if (p.rowscale != nullptr && lane < rv) rs = p.rowscale[(rb + lane) / p.rows_per_scale];
          const int acc = local & acc_mask;
          mbar_wait(&tfull_bar[acc], (uint32_t)(local >> acc_shift) & 1u);
          tc_fence_after();
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          uint32_t ex[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint4 v = ld_shared_v4(own_piece<8>(se, q));
            ex[4 * q] = v.x; ex[4 * q + 1] = v.y; ex[4 * q + 2] = v.z; ex[4 * q + 3] = v.w;
          }
          __syncwarp();
          if (tn < total_tiles)
            prefetch_rows_async(se, reinterpret_cast<const uint8_t*>(p.aux + rbo_n + c0_n), ld_bytes, lane, rv_n);
          if (rv > 0) {
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.block_n + half * 64);
            const StoreCtx sc{&tmD, &tmD2, &tmDh, &tmD2h, rb, tb2, tb1};
            epilogue_unit_compute<EPI_AUX, false, 8>(p, taddr, rbo, lane, rv, c0, rs, sa, ex, sc);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          rbo = rbo_n; rv = rv_n; c0 = c0_n; rb = rb_n; tb1 = tb1_n; tb2 = tb2_n;
        }
        total_tiles_done = true;
      }
    }
    uint32_t tempty_leader = 0u;   // the leader's MMA thread owns the accumulators of both CTAs
    if constexpr (kPair) tempty_leader = mapa_shared(smem_u32(&tempty_bar[0]), 0u);
    for (int local = group, t = MVLT_CID + group * MVLT_NCL; !total_tiles_done && t < total_tiles; t += 2 * MVLT_NCL, local += 2) {
      const TileCoord tc = decode_tile(p, t);
      const int n0 = tc.n_blk * p.block_n;
      const int row_base = (tc.m_blk * mb_mul + (int)MVLT_RANK) * BLOCK_M + quarter * 32;
      const int rows_valid = min(32, p.M - row_base);          // <= 0 when the whole warp is past the M tail
      const long long batch_off = (long long)tc.b1 * p.sD1 + (long long)tc.b2 * p.sD2;
      const long long row_base_off = batch_off + (long long)row_base * p.ldd;
      float rs = 1.f;
      if (p.rowscale != nullptr && lane < rows_valid) rs = p.rowscale[(row_base + lane) / p.rows_per_scale];
      const bool aligned = ld_aligned && ((batch_off & 7) == 0);
      const int acc = local & acc_mask;

      mbar_wait(&tfull_bar[acc], (uint32_t)(local >> acc_shift) & 1u);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * p.block_n);
      const StoreCtx ssc{&tmD, &tmD2, &tmDh, &tmD2h, row_base, tc.b2, tc.b1};
      if constexpr (kEpi == EPI_SOFTMAX) {
        // ---- fused row softmax (whole row in this tile): pass 1 = this warp's partial (max, sum) over its column
        // half, exchanged with the partner warp; pass 2 = normalise + store. The accumulator is read twice from TMEM.
        const float a2 = p.alpha * 1.4426950408889634f;   // exp(alpha*x) = exp2(a2*x)
        float mx = -INFINITY, sum = 0.f;
        for (int ci = c_begin; ci < c_end; ++ci) {
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + (uint32_t)(ci * 32), r);
          tmem_ld_wait();
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(r[j]) * a2);
          const float mn = fmaxf(mx, cm);
          float cs = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) cs += ex2_approx(fmaf(__uint_as_float(r[j]), a2, -mn));
          sum = fmaf(sum, ex2_approx(mx - mn), cs);
          mx = mn;
        }
        {
          const float2 o = pair_exchange(stage, partner_stage, pair_bar, lane, make_float2(mx, sum));
          const float mn = fmaxf(mx, o.x);
          // a warp without columns (block_n = 32) contributes (-inf, 0): guard the inf - inf
          const float s_me = (sum > 0.f) ? sum * ex2_approx(mx - mn) : 0.f;
          const float s_ot = (o.y > 0.f) ? o.y * ex2_approx(o.x - mn) : 0.f;
          sum = s_me + s_ot;
          mx = mn;
        }
        const float inv = 1.f / sum;
        if (rows_valid > 0) {
          for (int ci = c_begin; ci < c_end;) {
            if (ci + 2 <= c_end) {
              softmax_unit<8>(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, a2, mx, inv, sa, ssc);
              ci += 2;
            } else {
              softmax_unit<4>(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, a2, mx, inv, sa, ssc);
              ci += 1;
            }
          }
        }
      } else if constexpr (kEpi == EPI_SOFTMAX_BWD) {
        // ---- fused softmax backward: dS = alpha * P * (dP - sum_n P*dP)
        float dot = 0.f;
        for (int ci = c_begin; ci < c_end;) {
          if (ci + 2 <= c_end) {
            dot += softmax_bwd_unit<8, 0>(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, 0.f, sa, ssc);
            ci += 2;
          } else {
            dot += softmax_bwd_unit<4, 0>(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, 0.f, sa, ssc);
            ci += 1;
          }
        }
        dot += pair_exchange(stage, partner_stage, pair_bar, lane, make_float2(dot, 0.f)).x;
        for (int ci = c_begin; ci < c_end;) {
          if (ci + 2 <= c_end) {
            softmax_bwd_unit<8, 1>(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, dot, sa, ssc);
            ci += 2;
          } else {
            softmax_bwd_unit<4, 1>(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, dot, sa, ssc);
            ci += 1;
          }
        }
      } else if (kEpi == EPI_RESID && kOutF32 && p.ln_gamma != nullptr) {
        // ---- residual epilogue + LayerNorm of the finished row (the whole row is in this tile: n0 == 0). Pass 1 per unit:
        // D + partial (sum, sum of squares), exchanged with the partner warp of the other column half; pass 2: normalise
        // from the staging tiles (one per unit, they still hold the fp32 values) and store the bf16 operand of the next GEMM.
        // Both warps of a pair always meet on the named barrier, rows past M included (their stores are clipped by the maps).
        if constexpr (kEpi == EPI_RESID && kOutF32) {
          float sum = 0.f, sq = 0.f;
          for (int ci = c_begin, u = 0; ci < c_end; ++ci, ++u) {
            const StageAddr su = make_stage_addr(stage + (uint32_t)u * 4096u, lane);
            resid_ln_pass1(p, taddr0 + (uint32_t)(ci * 32), row_base_off, lane, rows_valid, ci * 32, rs, su, ssc, sum, sq);
          }
          const uint32_t xchg = smem_u32(stage_base) + (uint32_t)NUM_EPI_WARPS * (uint32_t)p.warp_stage_bytes;
          const float2 o = pair_exchange(xchg + (uint32_t)ew * 256u, xchg + (uint32_t)(ew ^ 8) * 256u, pair_bar, lane, make_float2(sum, sq));
          const float inv_n = 1.f / (float)p.N;
          const float mean = (sum + o.x) * inv_n;
          const float var = fmaxf(fmaf(-mean, mean, (sq + o.y) * inv_n), 0.f);
          const float rstd = rsqrtf(var + p.ln_eps);
          if (half == 0 && lane < rows_valid) {
            if (p.ln_mean != nullptr) p.ln_mean[row_base + lane] = mean;
            if (p.ln_rstd != nullptr) p.ln_rstd[row_base + lane] = rstd;
          }
          for (int ci = c_begin, u = 0; ci < c_end; ++ci, ++u) {
            const StageAddr su = make_stage_addr(stage + (uint32_t)u * 4096u, lane);
            resid_ln_pass2(p, lane, ci * 32, mean, rstd, su, ssc);
          }
        }
      } else if (rows_valid > 0) {
        const StoreCtx sc{&tmD, &tmD2, &tmDh, &tmD2h, row_base, tc.b2, tc.b1};
        for (int ci = c_begin; ci < c_end;) {
          const int col0 = n0 + ci * 32;
          const uint32_t taddr = taddr0 + (uint32_t)(ci * 32);
          if (col0 >= p.N) break;
          if (!aligned || col0 + 32 > p.N) {
            epilogue_chunk_tail(p, taddr, row_base_off, lane, rows_valid, col0, rs);
            ci += 1;
          } else if constexpr (kOutF32) {
            epilogue_unit_vec<kEpi, true, 8>(p, taddr, row_base_off, lane, rows_valid, col0, rs, sa, sc);
            ci += 1;
          } else {
            if (ci + 2 <= c_end && col0 + 64 <= p.N) {
              epilogue_unit_vec<kEpi, false, 8>(p, taddr, row_base_off, lane, rows_valid, col0, rs, sa, sc);
              ci += 2;
            } else {
              epilogue_unit_vec<kEpi, false, 4>(p, taddr, row_base_off, lane, rows_valid, col0, rs, sa, sc);
              ci += 1;
            }
          }
        }
      }
      if constexpr (kEpi == EPI_PLAIN && kOutF32) {
        if (p.rowsum != nullptr && tc.n_blk == 0 && half == 0) {   // row sums of A (bias gradient), once per (m, k-split)
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p.num_acc * p.block_n + acc * ONES_N), r);
          tmem_ld_wait();
          if (lane < rows_valid) atomicAdd(p.rowsum + row_base + lane, __uint_as_float(r[0]) * p.alpha);
        }
      }
      // all of this warp's TMEM reads are complete (wait::ld above): release the accumulator buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_cluster(tempty_leader + (uint32_t)acc * 8u);
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
  }

  if (p.tma_store && warp >= FIRST_EPI_WARP && lane == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this lane's bulk stores have fully completed
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // both CTAs have drained their accumulators and no remote arrive is in flight
  if (warp == 1) {
    __syncwarp();
    if constexpr (kPair) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}
#undef MVLT_CID
#undef MVLT_NCL
#undef MVLT_RANK
#undef MVLT_BROWS

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
                   const uint32_t box[4], int f32 = 0, int swizzle64 = 0) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  key.v[0] = (uint64_t)base;
  for (int i = 0; i < 4; ++i) key.v[1 + i] = dims[i];
  for (int i = 0; i < 3; ++i) key.v[5 + i] = strides_b[i];
  for (int i = 0; i < 4; ++i) key.v[8 + i] = box[i];
  key.v[11] |= (uint64_t)((f32 ? 1 : 0) | (swizzle64 ? 2 : 0)) << 32;
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
  CUresult r = fn(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.m = 0;
  f.sh = 0;
  if (d > 1) {
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                      // l = ceil(log2 d) >= 1
    f.m = (uint32_t)((((unsigned long long)1 << (31 + l)) + d - 1) / d);   // ceil(2^(31+l) / d) < 2^32 ... <= 2^32 - 1 for d > 2^(l-1)
    f.sh = l - 1;
  }
  return f;
}

typedef void (*GemmKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const CUtensorMap, const KParams);
struct KernelVariant {
  GemmKernelFn fn;
  const char* name;
};
const KernelVariant kVariants[] = {
    {gemm_tcgen05_kernel<EPI_PLAIN, false>, "plain/bf16"},       {gemm_tcgen05_kernel<EPI_PLAIN, true>, "plain/f32"},
    {gemm_tcgen05_kernel<EPI_GELU, false>, "gelu/bf16"},         {gemm_tcgen05_kernel<EPI_AUX, false>, "aux/bf16"},
    {gemm_tcgen05_kernel<EPI_RESID, true>, "residual/f32"},      {gemm_tcgen05_kernel<EPI_SOFTMAX, false>, "softmax/bf16"},
    {gemm_tcgen05_kernel<EPI_SOFTMAX_BWD, false>, "softmax_bwd/bf16"},
    // CTA-pair builds of the first five (index + 7)
    {gemm_tcgen05_kernel<EPI_PLAIN, false, true>, "plain/bf16/pair"},   {gemm_tcgen05_kernel<EPI_PLAIN, true, true>, "plain/f32/pair"},
    {gemm_tcgen05_kernel<EPI_GELU, false, true>, "gelu/bf16/pair"},     {gemm_tcgen05_kernel<EPI_AUX, false, true>, "aux/bf16/pair"},
    {gemm_tcgen05_kernel<EPI_RESID, true, true>, "residual/f32/pair"},
};
constexpr int kPairVariantOffset = 7;
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

// NHWC bf16 tensor X[b, y, x, c] as a 4-D map (c, x, y, b) whose box covers `pixels` consecutive pixels x 64 channels
int make_conv_map(CUtensorMap* out, const mvlt_gemm_desc* g, const void* base, int pixels) {
  const int W = g->conv_W, H = g->conv_H;
  if (((uintptr_t)base & 15) != 0 || (g->conv_pix_stride % 8) != 0 || (g->conv_batch_stride % 8) != 0) {
    mvlt_set_error("conv operand: base / strides must be 16-byte aligned (base=%p pix_stride=%lld batch_stride=%lld)", base,
                   (long long)g->conv_pix_stride, (long long)g->conv_batch_stride);
    return MVLT_ERR_ALIGN;
  }
  const int by = (pixels / W < H) ? pixels / W : H;   // rows of one image in the box
  const int bb = pixels / (W * by);                    // images in the box (small feature maps)
  uint64_t dims[4] = {(uint64_t)g->conv_C, (uint64_t)W, (uint64_t)H, (uint64_t)g->conv_B};
  uint64_t str[3] = {(uint64_t)g->conv_pix_stride * 2, (uint64_t)g->conv_pix_stride * 2 * W, (uint64_t)g->conv_batch_stride * 2};
  uint32_t box[4] = {64, (uint32_t)W, (uint32_t)by, (uint32_t)bb};
  return get_tensor_map(out, base, dims, str, box);
}

// NHWC bf16 tensor X as the 5-D patch view (kx*C + c, ox, ky, oy, b) of a kernel = stride = R convolution; box = 64 elements x
// all ox x one ky x all oy x 2 images = 128 rows x 128 bytes (not cached: a handful of launches per forward use it)
int make_patch_map(CUtensorMap* out, const mvlt_gemm_desc* g, const void* base) {
  const int R = g->conv_R, C = g->conv_C, ow = g->conv_W / R, oh = g->conv_H / R;
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    mvlt_set_error("cuTensorMapEncodeTiled not available from the driver");
    return MVLT_ERR_DRIVER;
  }
  cuuint64_t gdim[5] = {(cuuint64_t)R * C, (cuuint64_t)ow, (cuuint64_t)R, (cuuint64_t)oh, (cuuint64_t)g->conv_B};
  cuuint64_t gstr[4] = {(cuuint64_t)R * C * 2, (cuuint64_t)g->conv_W * C * 2, (cuuint64_t)R * g->conv_W * C * 2,
                        (cuuint64_t)g->conv_batch_stride * 2};
  cuuint32_t bx[5] = {64, (cuuint32_t)ow, 1, (cuuint32_t)oh, 2};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mvlt_set_error("cuTensorMapEncodeTiled (5-D patch view) failed (%d): B=%d H=%d W=%d C=%d R=%d", (int)r, g->conv_B, g->conv_H, g->conv_W, C, R);
    return MVLT_ERR_DRIVER;
  }
  return 0;
}

// fp32 NHWC tensor dX as the 5-D patch view (kx*C + c, ox, ky, oy, b); box = one staging tile: 32 fp32 x 8 ox x 1 ky x 4 oy x 1 image
int make_patch_store_map(CUtensorMap* out, const mvlt_gemm_desc* g) {
  const int R = g->conv_R, C = g->conv_C, ow = g->conv_W / R, oh = g->conv_H / R;
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    mvlt_set_error("cuTensorMapEncodeTiled not available from the driver");
    return MVLT_ERR_DRIVER;
  }
  cuuint64_t gdim[5] = {(cuuint64_t)R * C, (cuuint64_t)ow, (cuuint64_t)R, (cuuint64_t)oh, (cuuint64_t)g->conv_B};
  cuuint64_t gstr[4] = {(cuuint64_t)R * C * 4, (cuuint64_t)g->conv_W * C * 4, (cuuint64_t)R * g->conv_W * C * 4,
                        (cuuint64_t)g->conv_batch_stride * 4};
  cuuint32_t bx[5] = {32, (cuuint32_t)ow, 1, 4, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, g->D, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mvlt_set_error("cuTensorMapEncodeTiled (5-D patch store) failed (%d): B=%d H=%d W=%d C=%d R=%d", (int)r, g->conv_B, g->conv_H, g->conv_W, C, R);
    return MVLT_ERR_DRIVER;
  }
  return 0;
}

// Output tensor D (or D2) as a 4-D map (n, m, batch2, batch1) whose box is one staging tile: 32 rows x 128 bytes
// (SWIZZLE_128B), or 32 rows x 64 bytes (SWIZZLE_64B) for the 32-column bf16 remainder units
int make_output_map(CUtensorMap* out, const mvlt_gemm_desc* g, const void* base, int f32, int row_bytes = 128) {
  const uint64_t es = f32 ? 4 : 2;
  const int ub2 = (g->batch2 > 1) ? 1 : 0, ub1 = (g->batch1 > 1) ? 1 : 0;
  uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->M, ub2 ? (uint64_t)g->batch2 : 1, ub1 ? (uint64_t)g->batch1 : 1};
  const uint64_t dflt = (((uint64_t)g->M * (uint64_t)g->ldd * es + 15) / 16) * 16;
  uint64_t str[3] = {(uint64_t)g->ldd * es, ub2 ? (uint64_t)g->sD2 * es : dflt, ub1 ? (uint64_t)g->sD1 * es : dflt};
  uint32_t box[4] = {(uint32_t)(row_bytes / es), 32, 1, 1};
  return get_tensor_map(out, base, dims, str, box, f32, row_bytes == 64);
}

int pick_variant(const mvlt_gemm_desc* g) {
  if (g->act == MVLT_ACT_SOFTMAX) return 5;
  if (g->act == MVLT_ACT_SOFTMAX_BWD) return 6;
  if (g->residual != nullptr) return 4;
  if (g->aux != nullptr) return 3;
  if (!g->out_f32 && (g->act == MVLT_ACT_GELU || g->act == MVLT_ACT_GELU_SAVE_GRAD)) return 2;
  return g->out_f32 ? 1 : 0;
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

// tensor-map cache shared with attn_tcgen05.cu (C++ linkage: not part of the C-ABI)
int mvlt_tensor_map_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_b[3],
                       const uint32_t box[4], int f32, int swizzle64) {
  return get_tensor_map(out, base, dims, strides_b, box, f32, swizzle64);
}

extern "C" int mvlt_gemm(const mvlt_gemm_desc* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(g != nullptr, "mvlt_gemm: null descriptor");
  MVLT_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0, "mvlt_gemm: bad shape M=%d N=%d K=%d", g->M, g->N, g->K);
  MVLT_CHECK_ARG(g->batch1 >= 1 && g->batch2 >= 1, "mvlt_gemm: bad batch %d x %d", g->batch1, g->batch2);
  MVLT_CHECK_ARG(g->A && g->B && g->D, "mvlt_gemm: null operand");
  MVLT_CHECK_ARG(!(g->atomic_add && !g->out_f32), "mvlt_gemm: atomic_add needs an fp32 output");
  MVLT_CHECK_ARG(!(g->split_k > 1 && !g->atomic_add), "mvlt_gemm: split_k > 1 needs atomic_add");
  MVLT_CHECK_ARG(!((g->act == MVLT_ACT_DGELU || g->act == MVLT_ACT_MUL_AUX || g->act == MVLT_ACT_SOFTMAX_BWD) && g->aux == nullptr),
                 "mvlt_gemm: dgelu / mul_aux epilogue needs aux");
  MVLT_CHECK_ARG(!(g->act == MVLT_ACT_GELU_SAVE_GRAD && g->D2 == nullptr), "mvlt_gemm: gelu_save_grad needs D2");
  if (g->act == MVLT_ACT_SOFTMAX || g->act == MVLT_ACT_SOFTMAX_BWD) {
    MVLT_CHECK_ARG(g->N <= 256 && g->N % 32 == 0 && !g->out_f32 && !g->atomic_add && !g->bias && !g->residual &&
                       !g->rowscale && (g->ldd % 8) == 0 && (g->sD1 % 8) == 0 && (g->sD2 % 8) == 0 &&
                       (g->block_n == 0 || g->block_n == g->N) && (g->b_mn == 0 || g->N % 64 == 0),
                   "mvlt_gemm: row softmax epilogues need N <= 256, N %% 32 == 0, bf16 output, 16-byte aligned rows");
    MVLT_CHECK_ARG((g->act == MVLT_ACT_SOFTMAX) == (g->aux == nullptr), "mvlt_gemm: softmax_bwd needs aux = P, softmax none");
  }
  MVLT_CHECK_ARG(!((g->act == MVLT_ACT_GELU || g->act == MVLT_ACT_GELU_SAVE_GRAD) && (g->residual || g->aux)),
                 "mvlt_gemm: the GELU epilogues do not combine with residual / aux operands");
  // the epilogue streams residual / aux / D2 through the same staging-tile geometry as D
  MVLT_CHECK_ARG(!(g->residual && !g->out_f32), "mvlt_gemm: residual (fp32) needs an fp32 output");
  MVLT_CHECK_ARG(!(g->residual && g->aux), "mvlt_gemm: residual and aux are mutually exclusive");
  MVLT_CHECK_ARG(!(g->aux && g->out_f32), "mvlt_gemm: aux (bf16) needs a bf16 output");
  const bool ln = g->ln_gamma != nullptr;
  MVLT_CHECK_ARG(!(g->D2 && g->out_f32 && !ln), "mvlt_gemm: D2 (bf16) needs a bf16 output");
  if (ln)
    MVLT_CHECK_ARG(g->ln_beta && g->D2 && g->residual && g->out_f32 && (g->N == 64 || g->N == 128) && g->act == MVLT_ACT_NONE &&
                       !g->atomic_add && g->split_k <= 1 && g->batch1 == 1 && g->batch2 == 1 && g->conv_mode == MVLT_CONV_NONE &&
                       (g->block_n == 0 || g->block_n == g->N) && g->ldd % 8 == 0 && ((uintptr_t)g->ln_gamma & 15) == 0 &&
                       ((uintptr_t)g->ln_beta & 15) == 0 && g->rowsum == nullptr,
                   "mvlt_gemm: the fused LayerNorm epilogue needs residual + fp32 D + bf16 D2, N = 64 | 128 in one tile, no batch / split-K");
  MVLT_CHECK_ARG(!(g->act == MVLT_ACT_GELU_SAVE_GRAD && g->out_f32), "mvlt_gemm: gelu_save_grad needs a bf16 output");
  MVLT_CHECK_ARG(!(g->atomic_add && (g->residual || g->aux || g->act != MVLT_ACT_NONE)),
                 "mvlt_gemm: atomic_add only combines with alpha / bias / rowscale");
  MVLT_CHECK_ARG(!(g->aux && g->act != MVLT_ACT_MUL_AUX && g->act != MVLT_ACT_DGELU && g->act != MVLT_ACT_SOFTMAX_BWD),
                 "mvlt_gemm: aux is only consumed by the mul_aux / dgelu / softmax_bwd epilogues");
  MVLT_CHECK_ARG(!(g->rowscale && g->rows_per_scale <= 0), "mvlt_gemm: rowscale needs rows_per_scale");

  if (g->patch_store) {
    const int R = g->conv_R;
    MVLT_CHECK_ARG(g->out_f32 && !g->atomic_add && g->split_k <= 1 && !g->bias && !g->residual && !g->aux && !g->D2 && !g->rowscale &&
                       !g->rowsum && !ln && g->act == MVLT_ACT_NONE && g->batch1 == 1 && g->batch2 == 1 && g->conv_mode == MVLT_CONV_NONE,
                   "mvlt_gemm: patch_store takes a plain fp32 product (no epilogue operands, no batch, no split-K)");
    MVLT_CHECK_ARG(R > 1 && g->conv_B > 0 && g->conv_C > 0 && g->conv_H % R == 0 && g->conv_W % R == 0 && g->conv_W / R == 8 &&
                       g->conv_H / R == 8 && ((long long)R * g->conv_C) % 64 == 0 && g->conv_pix_stride == g->conv_C &&
                       g->conv_batch_stride % 4 == 0 && ((uintptr_t)g->D & 15) == 0 && g->M == g->conv_B * 64 &&
                       g->N == R * R * g->conv_C,
                   "mvlt_gemm: unsupported patch_store geometry B=%d H=%d W=%d C=%d R=%d M=%d N=%d (need an 8 x 8 reduced map)",
                   g->conv_B, g->conv_H, g->conv_W, g->conv_C, R, g->M, g->N);
  }
  KParams p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.patch_store = g->patch_store ? 1 : 0;
  p.patch_rc = g->patch_store ? g->conv_R * g->conv_C : 1;
  p.block_n = (g->act == MVLT_ACT_SOFTMAX || g->act == MVLT_ACT_SOFTMAX_BWD) ? g->N : pick_block_n(g);
  MVLT_CHECK_ARG(p.block_n >= 32 && p.block_n <= 256 && p.block_n % (g->b_mn ? 64 : 32) == 0,
                 "mvlt_gemm: unsupported block_n %d", p.block_n);
  // thin-K aux epilogues: 128-wide tiles, a second staging tile per epilogue warp (cp.async prefetch of the aux operand)
  const bool aux_kind = g->aux != nullptr && (g->act == MVLT_ACT_MUL_AUX || g->act == MVLT_ACT_DGELU);
  p.prefetch = (aux_kind && g->K <= 2 * BLOCK_K && g->N % 128 == 0 && (g->ldd % 8) == 0 && (g->sD1 % 8) == 0 && (g->sD2 % 8) == 0 &&
                (g->block_n == 0 || g->block_n == 128)) ? 1 : 0;
  if (p.prefetch) p.block_n = 128;
  if (ln) p.block_n = g->N;
  p.warp_stage_bytes = p.prefetch ? 8192 : (ln ? 4096 * (g->N / 64) : 4096);
  const int staging_total = NUM_EPI_WARPS * p.warp_stage_bytes + (ln ? 4096 : 0);   // (+ the LayerNorm partial-sum exchange area)
  p.ln_gamma = g->ln_gamma; p.ln_beta = g->ln_beta; p.ln_mean = g->ln_mean; p.ln_rstd = g->ln_rstd; p.ln_eps = g->ln_eps;
  // CTA-pair mode (cta_group::2) for the tensor-bound shapes: a 256 x block_n tile per cluster of two CTAs halves the B bytes
  // each SM stages per k-block (32 KB instead of 48 KB at block_n = 256: 5 instead of 3 pipeline stages). Validated
  // (tests/test_gemm_gpu.py passes with MVLT_GEMM_PAIR=1) but measured SLOWER than the single-CTA build on every PVLT shape
  // (stage-4 MLP fc1 84.1 -> 90.8 us, fc2 72.0 -> 76.3 us, profiles/r2m_gemm_pair_ab.txt): these launches are bound by their
  // GELU / residual epilogues and by wave quantisation, not by the operand fill. Opt-in: MVLT_GEMM_PAIR=1.
  static const int pair_env = [] { const char* e = getenv("MVLT_GEMM_PAIR"); return e ? atoi(e) : 0; }();
  p.pair = (pair_env && !ln && !g->patch_store && g->conv_mode == MVLT_CONV_NONE && g->rowsum == nullptr && !p.prefetch && g->split_k <= 1 && !g->atomic_add &&
            g->act != MVLT_ACT_SOFTMAX && g->act != MVLT_ACT_SOFTMAX_BWD && g->K >= 4 * BLOCK_K && g->M >= 16 * BLOCK_M &&
            p.block_n >= 128 && p.block_n % 128 == 0 && (long long)g->batch1 * g->batch2 == 1) ? 1 : 0;
  const int stage_bytes = A_STAGE_BYTES + (p.pair ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;
  p.stages = (SMEM_BUDGET + STAGING_BYTES - staging_total) / stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.num_m_blocks = (g->M + BLOCK_M - 1) / BLOCK_M;
  if (p.pair) p.num_m_blocks = (p.num_m_blocks + 1) / 2;      // 256-row pairs; an odd last block is zero-filled / clipped
  p.num_n_blocks = (g->N + p.block_n - 1) / p.block_n;
  p.num_k_blocks = (g->K + BLOCK_K - 1) / BLOCK_K;
  int split = g->split_k > 1 ? g->split_k : 1;
  if (split > p.num_k_blocks) split = p.num_k_blocks;
  p.kb_per_split = (p.num_k_blocks + split - 1) / split;
  p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
  p.batch1 = g->batch1; p.batch2 = g->batch2;
  p.num_acc = (p.block_n <= 128) ? 4 : 2;
  p.rowsum = g->rowsum;
  if (g->rowsum != nullptr) {
    MVLT_CHECK_ARG(g->out_f32 && !g->residual && !g->aux && g->act == MVLT_ACT_NONE && g->batch1 == 1 && g->batch2 == 1,
                   "mvlt_gemm: rowsum needs a plain fp32 (atomic) output and no batch");
    if (p.block_n > 240) {
      // 2 x 256 accumulator columns fill TMEM: no room for the row-sum columns. Narrower tiles would cost more than
      // the fusion saves on these (tensor-bound) shapes, so the sums come from the column-sum kernel instead.
      MVLT_CHECK_ARG(g->a_mn && g->alpha == 1.0f, "mvlt_gemm: rowsum with 256-wide tiles needs an MN-major A and alpha = 1");
      const int rc2 = mvlt_colsum(g->A, 0, g->K, g->M, g->lda, g->rowsum, stream_);
      if (rc2) return rc2;
      p.rowsum = nullptr;
    } else {
      p.num_acc = 2;   // TMEM: 2 x block_n accumulator columns + 2 x 16 row-sum columns <= 512
    }
  }
  p.div_n = make_fastdiv((uint32_t)p.num_n_blocks);
  p.div_m = make_fastdiv((uint32_t)p.num_m_blocks);
  p.div_s = make_fastdiv((uint32_t)p.split_k);
  p.div_b2 = make_fastdiv((uint32_t)p.batch2);
  p.a_mn = g->a_mn ? 1 : 0; p.b_mn = g->b_mn ? 1 : 0;
  p.D = g->D; p.D2 = g->D2; p.bias = g->bias;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(g->aux);
  p.residual = g->residual; p.rowscale = g->rowscale;
  p.ldd = g->patch_store ? g->N : g->ldd; p.sD1 = g->sD1; p.sD2 = g->sD2;   // (patch_store: D is addressed by its tensor map only)
  p.alpha = g->alpha;
  p.act = g->act; p.out_f32 = g->out_f32; p.atomic_add = g->atomic_add;
  p.rows_per_scale = g->rows_per_scale > 0 ? g->rows_per_scale : 1;

  CUtensorMap tmA, tmB;
  int rc = 0;
  p.conv_mode = g->conv_mode;
  if (g->conv_mode == MVLT_CONV_PATCH_A) {
    const int R = g->conv_R;
    MVLT_CHECK_ARG(g->batch1 == 1 && g->batch2 == 1 && !g->a_mn && !p.pair, "mvlt_gemm: the patch view is a plain K-major A operand");
    MVLT_CHECK_ARG(R > 1 && g->conv_B > 0 && g->conv_C > 0 && g->conv_H % R == 0 && g->conv_W % R == 0 &&
                       (g->conv_H / R) * (g->conv_W / R) == 64 && ((long long)R * g->conv_C) % 64 == 0 && g->conv_pix_stride == g->conv_C &&
                       g->conv_batch_stride % 8 == 0 && ((uintptr_t)g->A & 15) == 0 && g->conv_W / R <= 256 && g->conv_H / R <= 256,
                   "mvlt_gemm: unsupported patch geometry B=%d H=%d W=%d C=%d R=%d (need 64 output pixels per image, R*C %% 64 == 0)",
                   g->conv_B, g->conv_H, g->conv_W, g->conv_C, R);
    MVLT_CHECK_ARG(g->M == g->conv_B * 64 && g->K == R * R * g->conv_C, "mvlt_gemm: patch view needs M = B*64, K = R*R*C");
    p.div_cb = make_fastdiv((uint32_t)(R * g->conv_C / 64));
  } else if (g->conv_mode != MVLT_CONV_NONE) {
    const long long HW = (long long)g->conv_H * g->conv_W, pix = HW * g->conv_B;
    MVLT_CHECK_ARG(g->conv_mode == MVLT_CONV_A || g->conv_mode == MVLT_CONV_BT, "mvlt_gemm: bad conv_mode %d", g->conv_mode);
    MVLT_CHECK_ARG(g->batch1 == 1 && g->batch2 == 1, "mvlt_gemm: implicit convolution is not batched");
    MVLT_CHECK_ARG(g->conv_B > 0 && g->conv_H > 0 && g->conv_W > 0 && g->conv_C > 0 && g->conv_C % 64 == 0 &&
                       g->conv_W <= 64 && 64 % g->conv_W == 0 && HW % 64 == 0 && (HW >= 128 ? HW % 128 == 0 : 128 % HW == 0),
                   "mvlt_gemm: unsupported conv geometry B=%d H=%d W=%d C=%d", g->conv_B, g->conv_H, g->conv_W, g->conv_C);
    p.conv_W = g->conv_W;
    p.conv_C = g->conv_C;
    p.div_hw = make_fastdiv((uint32_t)HW);
    p.div_w = make_fastdiv((uint32_t)g->conv_W);
    p.div_cb = make_fastdiv((uint32_t)(g->conv_C / 64));
    p.div_c = make_fastdiv((uint32_t)g->conv_C);
    if (g->conv_mode == MVLT_CONV_A)
      MVLT_CHECK_ARG(g->M == pix && g->K == 9 * g->conv_C && !g->a_mn, "mvlt_gemm: conv A needs M = B*H*W, K = 9*C, K-major");
    else
      MVLT_CHECK_ARG(g->K == pix && g->N == 9 * g->conv_C && g->b_mn && p.block_n % 64 == 0,
                     "mvlt_gemm: conv B^T needs K = B*H*W, N = 9*C, MN-major");
  }
  if (g->conv_mode == MVLT_CONV_PATCH_A) rc = make_patch_map(&tmA, g, g->A);
  else if (g->conv_mode == MVLT_CONV_A) rc = make_conv_map(&tmA, g, g->A, BLOCK_M);
  else rc = make_operand_map(&tmA, g->A, g->M, g->K, p.a_mn, g->lda, g->batch1, g->batch2, g->sA1, g->sA2, BLOCK_M, &p.a_b1, &p.a_b2);
  if (rc) return rc;
  if (g->conv_mode == MVLT_CONV_BT) rc = make_conv_map(&tmB, g, g->B, 64);
  else rc = make_operand_map(&tmB, g->B, g->N, g->K, p.b_mn, g->ldb, g->batch1, g->batch2, g->sB1, g->sB2, p.pair ? p.block_n / 2 : p.block_n, &p.b_b1, &p.b_b2);
  if (rc) return rc;

  // > half of the SM's shared memory so two CTAs (each wanting all 512 TMEM columns) never share an SM
  size_t smem = (size_t)p.stages * stage_bytes + 1024 /*align slack*/ + 2048 /*ones tile + barriers, keeps staging 1 KB aligned*/ +
                (size_t)staging_total;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static std::once_flag attr_once;
  static int launch_regs_ok = 1;
  std::call_once(attr_once, [] {
    for (int v = 0; v < kNumVariants; ++v) {
      cudaFuncSetAttribute(kVariants[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      // the setmaxnreg budget assumes the launch allocation; a smaller one would make setmaxnreg.inc block forever
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, kVariants[v].fn) != cudaSuccess || fa.numRegs != LAUNCH_REGS) launch_regs_ok = 0;
    }
  });
  MVLT_CHECK_ARG(launch_regs_ok, "mvlt_gemm: a kernel variant was not built with exactly %d registers per thread", LAUNCH_REGS);
  const long long total_tiles =
      (long long)p.batch1 * p.batch2 * p.split_k * p.num_m_blocks * p.num_n_blocks;
  MVLT_CHECK_ARG(total_tiles < (1ll << 31), "mvlt_gemm: too many tiles");
  int grid = mvlt_num_sms();
  if (p.pair) {
    grid &= ~1;
    if (2 * total_tiles < grid) grid = (int)(2 * total_tiles);
  } else if (total_tiles < grid) {
    grid = (int)total_tiles;
  }
  // TMA-store epilogue (bulk tensor stores, or bulk tensor fp32 reductions for atomic_add): 16-byte aligned output pitch /
  // batch strides / base
  CUtensorMap tmD = tmA, tmD2 = tmA, tmDh = tmA, tmD2h = tmA;
  {
    const long long es = g->out_f32 ? 4 : 2;
    const bool ok = ((uintptr_t)g->D & 15) == 0 && (g->ldd * es) % 16 == 0 && (g->sD1 * es) % 16 == 0 && (g->sD2 * es) % 16 == 0 &&
                    (g->batch1 == 1 || g->sD1 != 0) && (g->batch2 == 1 || g->sD2 != 0) &&
                    (g->D2 == nullptr || ((uintptr_t)g->D2 & 15) == 0);
    if (g->patch_store) {
      rc = make_patch_store_map(&tmD, g);
      if (rc) return rc;
      p.tma_store = 1;
    } else if (ok) {
      rc = make_output_map(&tmD, g, g->D, g->out_f32);
      if (rc) return rc;
      if (g->D2 != nullptr) {
        rc = make_output_map(&tmD2, g, g->D2, 0);
        if (rc) return rc;
      }
      if (ln) {            // the bf16 normalised rows leave as 32-column units
        rc = make_output_map(&tmD2h, g, g->D2, 0, 64);
        if (rc) return rc;
      }
      if (!g->out_f32) {   // 32-column remainder units of bf16 outputs
        rc = make_output_map(&tmDh, g, g->D, 0, 64);
        if (rc) return rc;
        if (g->D2 != nullptr) {
          rc = make_output_map(&tmD2h, g, g->D2, 0, 64);
          if (rc) return rc;
        }
      }
      p.tma_store = 1;
    }
  }
  MVLT_CHECK_ARG(!ln || p.tma_store, "mvlt_gemm: the fused LayerNorm epilogue needs 16-byte aligned D / D2 (TMA stores)");
  if (p.pair) mvlt_launch_cluster(kVariants[pick_variant(g) + kPairVariantOffset].fn, grid, NUM_THREADS, smem, stream, 2u, tmA, tmB, tmD, tmD2, tmDh, tmD2h, p);
  else mvlt_launch(kVariants[pick_variant(g)].fn, grid, NUM_THREADS, smem, stream, tmA, tmB, tmD, tmD2, tmDh, tmD2h, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}
