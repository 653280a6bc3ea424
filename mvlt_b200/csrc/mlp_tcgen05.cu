// Fused transformer MLP for sm_100a (B200): out = residual + rowscale * (GELU(X W1^T + b1) W2^T + b2) in ONE kernel --
// the hidden activation [M x Hd] never visits HBM.
//
// PVLT's stage-1/2 MLPs (C = 64 / 128, hidden = 8 C) are the most HBM-hungry part of the step when run as two GEMMs:
// the bf16 hidden tensor is 554 MB at stage 1 (B = 128) and is written once (twice in training, with gelu') and read
// once per GEMM that consumes it. Here a persistent CTA walks 128-row tiles; per tile the hidden axis is processed in
// chunks of 64 columns (one SWIZZLE_128B atom of bf16):
//
//   warp 0      : TMA producer: X tile [128 x C] (double-buffered) and, per chunk, W1[chunk rows, :] and W2[:, chunk] (ring)
//   warp 1      : tcgen05.mma issuer:  GEMM1  H_g[128 x 64]  = X W1_g^T            (K = C)   -> TMEM H buffer g % NH
//                                      GEMM2  Y  [128 x C ] += A_g W2_g^T          (K = 64)  -> TMEM Y buffer tile % 2
//                 GEMM2 runs two chunks behind GEMM1, so the tensor pipe never waits for the GELU warps
//   warps 4..19 : GELU warps (lane quarter x 16-column slice; 4 per SM sub-partition: the kernel is bound by their MUFU /
//                 issue rate): H_g (tcgen05.ld) + b1 -> exact-erf GELU (packed fp32x2) -> bf16 -> A_g in shared memory,
//                 written straight into the K-major SWIZZLE_128B layout GEMM2 reads
//   warps 20..23: output warps: Y + b2, x drop-path factor, + fp32 residual -> 4 KB swizzled staging tile -> TMA store
//
// The same kernel serves inference (retrieval) and training; the backward recomputes H from X (mlp_bwd below) instead of
// reading saved activations.
//
// Replaces /root/reference/libs/pvlt.py:65-71 (Mlp.forward) + the residual / DropPath of Block.forward (:142).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

int mvlt_tensor_map_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_b[3],
                       const uint32_t box[4], int f32, int swizzle64);   // gemm_tcgen05.cu

namespace {

constexpr int BM = 128;            // rows per tile = TMEM lanes
constexpr int CW = 64;             // hidden chunk width (bf16: one 128-byte swizzle row)
constexpr int ATOM = BM * 128;     // 16 KB: [128 rows x 64 bf16] K-major SWIZZLE_128B atom
constexpr int NTHREADS = 768;
constexpr int GELU_WARP0 = 4, NUM_GELU_WARPS = 16, OUT_WARP0 = 20, NUM_OUT_WARPS = 4;
constexpr int LAG = 2;             // GEMM2 of chunk g is issued after GEMM1 of chunk g + LAG
constexpr int TMEM_COLS = 512;
constexpr int Y_COL0 = 256;        // Y accumulators live in columns [256, 256 + 2 C); H buffers in [0, NH * 64)

template <int C>
struct Cfg {
  static constexpr int NH = (C == 64) ? 4 : 3;          // H accumulators (TMEM) = A buffers (shared memory)
  static constexpr int NS = (C == 64) ? 4 : 3;          // weight-chunk ring stages
  static constexpr int X_BYTES = BM * C * 2;            // 16 / 32 KB, C / 64 atoms
  static constexpr int W1_BYTES = CW * C * 2;           // [64 hidden rows x C]: C / 64 atoms of 8 KB
  static constexpr int W2_BYTES = C * CW * 2;           // [C rows x 64 hidden]: one box
  static constexpr int W_STAGE = W1_BYTES + W2_BYTES;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_W = OFF_X + 2 * X_BYTES;
  static constexpr int OFF_A = OFF_W + NS * W_STAGE;
  static constexpr int OFF_ST = OFF_A + NH * ATOM;      // 4 x 4 KB output staging tiles
  static constexpr int OFF_BAR = OFF_ST + NUM_OUT_WARPS * 4096;
  static constexpr int SMEM_USED = OFF_BAR + 256;
  static_assert(NH > LAG && NS > LAG, "the GEMM2 lag needs deeper rings");
  static_assert(NH * CW <= Y_COL0 && Y_COL0 + 2 * C <= TMEM_COLS, "TMEM plan");
  static_assert(SMEM_USED + 1024 <= 232448, "shared memory plan");
};

struct MlpParams {
  int M, C, HD;
  int num_tiles, nch;       // 128-row tiles, hidden chunks per tile
  const float* b1;
  const float* b2;
  const float* residual;    // fp32 [M, C]
  const float* rowscale;    // per-sample drop-path factor or nullptr
  int rows_per_scale;
  int dbg;                  // tuning experiments only (MVLT_MLP_DBG): 1 = no residual read, 2 = identity instead of GELU, 4 = no output
  // LayerNorm of the OUTPUT rows (the next block's norm1) fused into the output warps (ln_gamma != nullptr): every lane owns a
  // whole row, so the statistics need no exchange; the normalised bf16 rows leave through a second tensor map
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;           // optional fp32 [M]
  float* ln_rstd;
  float ln_eps;
};

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t sbo_bytes) {   // K-major SWIZZLE_128B operand
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ uint32_t instr_desc(int n) {   // kind::f16, bf16 x bf16 -> fp32, both operands K-major, M = 128
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmLn, const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
               const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ MlpParams p) {
  using K = Cfg<C>;
  constexpr int NH = K::NH, NS = K::NS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  uint64_t* x_full = bars;              // [2]  X tile landed (TMA tx)
  uint64_t* x_empty = bars + 2;         // [2]  every GEMM1 of the tile retired
  uint64_t* w_full = bars + 4;          // [NS] weight chunk landed
  uint64_t* w_empty = bars + 8;         // [NS] GEMM2 of the chunk retired (GEMM1 was issued before it)
  uint64_t* h_full = bars + 12;         // [NH] GEMM1 of the chunk retired: H readable
  uint64_t* a_full = bars + 16;         // [NH] the eight GELU warps have written A (and finished reading H)
  uint64_t* a_empty = bars + 20;        // [NH] GEMM2 of the chunk retired: A buffer reusable
  uint64_t* y_full = bars + 24;         // [2]  last GEMM2 of the tile retired
  uint64_t* y_empty = bars + 26;        // [2]  the four output warps have drained Y
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], NUM_OUT_WARPS);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < NH; ++i) {
      mbar_init(&h_full[i], 1);
      mbar_init(&a_full[i], NUM_GELU_WARPS);
      mbar_init(&a_empty[i], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int nch = p.nch;
  int my_tiles = 0;
  if ((int)blockIdx.x < p.num_tiles) my_tiles = (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;
  const int total = my_tiles * nch;       // chunks this CTA processes

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int g = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * BM;
        const int xb = tl & 1;
        mbar_wait(&x_empty[xb], (((uint32_t)tl >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&x_full[xb], (uint32_t)K::X_BYTES);
#pragma unroll
        for (int a = 0; a < C / 64; ++a) tma_load_4d(smem + K::OFF_X + xb * K::X_BYTES + a * ATOM, &tmX, &x_full[xb], a * 64, m0, 0, 0);
        for (int j = 0; j < nch; ++j, ++g) {
          const int s = g % NS;
          mbar_wait(&w_empty[s], (((uint32_t)(g / NS)) & 1u) ^ 1u);
          uint8_t* w = smem + K::OFF_W + s * K::W_STAGE;
          mbar_arrive_expect_tx(&w_full[s], (uint32_t)K::W_STAGE);
#pragma unroll
          for (int a = 0; a < C / 64; ++a) tma_load_4d(w + a * (CW * 128), &tmW1, &w_full[s], a * 64, j * CW, 0, 0);
          tma_load_4d(w + K::W1_BYTES, &tmW2, &w_full[s], j * CW, 0, 0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = instr_desc(CW);    // GEMM1: N = 64 hidden columns
      const uint32_t idesc2 = instr_desc(C);     // GEMM2: N = C
      const uint32_t sX = smem_u32(smem + K::OFF_X), sW = smem_u32(smem + K::OFF_W), sA = smem_u32(smem + K::OFF_A);
      int tl1 = 0, j1 = 0, tl2 = 0, j2 = 0;        // (tile, chunk-in-tile) of the GEMM1 / GEMM2 being issued: counters, no division
      for (int g = 0; g < total + LAG; ++g) {
        if (g < total) {                          // ---- GEMM1 of chunk g
          const int tl = tl1, j = j1;
          if (++j1 == nch) {
            j1 = 0;
            ++tl1;
          }
          const int xb = tl & 1, s = g % NS, hb = g % NH;
          if (j == 0) mbar_wait(&x_full[xb], ((uint32_t)tl >> 1) & 1u);
          mbar_wait(&w_full[s], ((uint32_t)(g / NS)) & 1u);
          tc_fence_after();
          // (the H buffer is free: GEMM2 of chunk g - NH, which waited for its GELU warps, was issued LAG < NH steps ago)
          const uint32_t xa = sX + (uint32_t)(xb * K::X_BYTES), wa = sW + (uint32_t)(s * K::W_STAGE);
#pragma unroll
          for (int k = 0; k < C / 16; ++k)
            umma_bf16(tmem_base + (uint32_t)(hb * CW), smem_desc(xa + (uint32_t)((k >> 2) * ATOM + (k & 3) * 32), 1024u),
                      smem_desc(wa + (uint32_t)((k >> 2) * (CW * 128) + (k & 3) * 32), 1024u), idesc1, k > 0 ? 1u : 0u);
          umma_commit(&h_full[hb]);
          if (j == nch - 1) umma_commit(&x_empty[xb]);
        }
        if (g >= LAG) {                           // ---- GEMM2 of chunk g - LAG
          const int g2 = g - LAG;
          const int tl = tl2, j = j2;
          if (++j2 == nch) {
            j2 = 0;
            ++tl2;
          }
          const int yb = tl & 1, s = g2 % NS, hb = g2 % NH;
          mbar_wait(&a_full[hb], ((uint32_t)(g2 / NH)) & 1u);
          if (j == 0) mbar_wait(&y_empty[yb], (((uint32_t)tl >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t aa = sA + (uint32_t)(hb * ATOM), wb = sW + (uint32_t)(s * K::W_STAGE + K::W1_BYTES);
#pragma unroll
          for (int k = 0; k < CW / 16; ++k)
            umma_bf16(tmem_base + (uint32_t)(Y_COL0 + yb * C), smem_desc(aa + (uint32_t)(k * 32), 1024u),
                      smem_desc(wb + (uint32_t)(k * 32), 1024u), idesc2, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(&w_empty[s]);
          umma_commit(&a_empty[hb]);
          if (j == nch - 1) umma_commit(&y_full[yb]);
        }
      }
    }
  } else if (warp >= GELU_WARP0 && warp < GELU_WARP0 + NUM_GELU_WARPS) {
    // ===================== GELU warps =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access (warp id % 4)
    const int cq = (warp - GELU_WARP0) >> 2;      // 16-column slice of the 64-column chunk
    const uint32_t tlane = ((uint32_t)(quarter * 32) << 16);
    const uint32_t row = (uint32_t)(quarter * 32 + lane);
    const uint32_t a_row = smem_u32(smem + K::OFF_A) + row * 128u;
    const uint32_t rx = row & 7u;
    int hb = 0, j = 0;                            // g % NH and g % nch, kept as counters (no integer division per chunk)
    uint32_t ph = 0;                              // (g / NH) & 1
    for (int g = 0; g < total; ++g) {
      // bias of this chunk slice: identical for every lane (L1 broadcast), fetched before the accumulator is waited for
      float4 bv[4];
      const float4* bp = reinterpret_cast<const float4*>(p.b1 + j * CW + cq * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) bv[i] = __ldg(bp + i);
      mbar_wait(&h_full[hb], ph);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(hb * CW + cq * 16), r);
      tmem_ld_wait();
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2_t v0 = f2_add(f2_pack(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])), f2_pack(bv[i].x, bv[i].y));
        const f32x2_t v1 = f2_add(f2_pack(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])), f2_pack(bv[i].z, bv[i].w));
        float a, b;
        f2_unpack(gelu2(v0), a, b);
        pk[2 * i] = pack_bf16x2(a, b);
        f2_unpack(gelu2(v1), a, b);
        pk[2 * i + 1] = pack_bf16x2(a, b);
      }
      mbar_wait(&a_empty[hb], ph ^ 1u);           // GEMM2 of chunk g - NH has finished reading this A buffer
      const uint32_t base = a_row + (uint32_t)(hb * ATOM);
#pragma unroll
      for (int q = 0; q < 2; ++q)
        st_shared_v4(base + ((((uint32_t)(cq * 2 + q)) ^ rx) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async();     // generic-proxy stores -> visible to the tensor core
      tc_fence_before();       // this warp's TMEM reads precede the MMA that will overwrite the H buffer
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[hb]);
      if (++hb == NH) {
        hb = 0;
        ph ^= 1u;
      }
      if (++j == nch) j = 0;
    }
  } else if (warp >= OUT_WARP0) {
    // ===================== output warps =====================
    const int quarter = warp & 3;
    const uint32_t tlane = ((uint32_t)(quarter * 32) << 16);
    const uint32_t tile_s = smem_u32(smem + K::OFF_ST) + (uint32_t)((warp - OUT_WARP0) * 4096);
    const uint32_t own = tile_s + (uint32_t)lane * 128u;
    const uint32_t rx = (uint32_t)(lane & 7);
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int m0 = ((int)blockIdx.x + tl * (int)gridDim.x) * BM;
      const int yb = tl & 1;
      const int row = m0 + quarter * 32 + lane;
      const bool valid = row < p.M;
      float rs = 1.f;
      if (p.rowscale != nullptr && valid) rs = p.rowscale[row / p.rows_per_scale];
      const float* res_row = p.residual + (long long)row * C;
      float ln_sum = 0.f, ln_sq = 0.f;
#pragma unroll 1
      for (int u = 0; u < C / 32; ++u) {
        if (u == 0) {
          mbar_wait(&y_full[yb], ((uint32_t)tl >> 1) & 1u);
          tc_fence_after();
        }
        if (p.dbg & 4) continue;
        if (lane == 0) tma_store_wait_read();     // the previous unit's store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {          // 16 columns at a time keeps the register footprint small
          float4 rv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rv[i] = (valid && !(p.dbg & 1)) ? __ldg(reinterpret_cast<const float4*>(res_row + u * 32 + hh * 16) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          uint32_t r[16];
          tmem_ld_32x16(tmem_base + tlane + (uint32_t)(Y_COL0 + yb * C + u * 32 + hh * 16), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + u * 32 + hh * 16) + i);
            const float o0 = fmaf(__uint_as_float(r[4 * i]) + b4.x, rs, rv[i].x);
            const float o1 = fmaf(__uint_as_float(r[4 * i + 1]) + b4.y, rs, rv[i].y);
            const float o2 = fmaf(__uint_as_float(r[4 * i + 2]) + b4.z, rs, rv[i].z);
            const float o3 = fmaf(__uint_as_float(r[4 * i + 3]) + b4.w, rs, rv[i].w);
            ln_sum += (o0 + o1) + (o2 + o3);
            ln_sq = fmaf(o0, o0, fmaf(o1, o1, fmaf(o2, o2, fmaf(o3, o3, ln_sq))));
            st_shared_v4(own + ((((uint32_t)(hh * 4 + i)) ^ rx) << 4), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2),
                         __float_as_uint(o3));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmOut, tile_s, u * 32, m0 + quarter * 32, 0, 0);   // rows past M are clipped by the tensor map
          tma_store_commit();
        }
      }
      if (p.ln_gamma != nullptr && !(p.dbg & 4)) {
        // ---- LayerNorm of the finished rows (the next block's norm1, libs/pvlt.py:140): second pass over the accumulator
        // (still in TMEM) and the residual (L1 / L2 now), same arithmetic -> the same fp32 values, normalised and stored as
        // the bf16 operand of the next block's q / kv projections, 64 columns (= 128-byte rows of the staging tile) at a time
        const float mean = ln_sum * (1.f / (float)C);
        const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, ln_sq * (1.f / (float)C)), 0.f) + p.ln_eps);
        if (valid) {
          if (p.ln_mean != nullptr) p.ln_mean[row] = mean;
          if (p.ln_rstd != nullptr) p.ln_rstd[row] = rstd;
        }
#pragma unroll 1
        for (int u2 = 0; u2 < C / 64; ++u2) {
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            const int c0 = u2 * 64 + hh * 16;
            float4 rv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              rv[i] = (valid && !(p.dbg & 1)) ? __ldg(reinterpret_cast<const float4*>(res_row + c0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t r[16];
            tmem_ld_32x16(tmem_base + tlane + (uint32_t)(Y_COL0 + yb * C + c0), r);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + c0) + i);
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + c0) + i);
              const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + c0) + i);
              const float o0 = fmaf(__uint_as_float(r[4 * i]) + b4.x, rs, rv[i].x);
              const float o1 = fmaf(__uint_as_float(r[4 * i + 1]) + b4.y, rs, rv[i].y);
              const float o2 = fmaf(__uint_as_float(r[4 * i + 2]) + b4.z, rs, rv[i].z);
              const float o3 = fmaf(__uint_as_float(r[4 * i + 3]) + b4.w, rs, rv[i].w);
              pk[2 * i] = pack_bf16x2(fmaf((o0 - mean) * rstd, g4.x, e4.x), fmaf((o1 - mean) * rstd, g4.y, e4.y));
              pk[2 * i + 1] = pack_bf16x2(fmaf((o2 - mean) * rstd, g4.z, e4.z), fmaf((o3 - mean) * rstd, g4.w, e4.w));
            }
            st_shared_v4(own + ((((uint32_t)(hh * 2)) ^ rx) << 4), pk[0], pk[1], pk[2], pk[3]);
            st_shared_v4(own + ((((uint32_t)(hh * 2 + 1)) ^ rx) << 4), pk[4], pk[5], pk[6], pk[7]);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmLn, tile_s, u2 * 64, m0 + quarter * 32, 0, 0);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_empty[yb]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int C>
int launch_fwd(const void* x, const void* w1, const void* w2, void* out, void* ln_out, const MlpParams& p, cudaStream_t stream) {
  using K = Cfg<C>;
  CUtensorMap tmX, tmW1, tmW2, tmOut, tmLn;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.M, 1, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)p.M * C * 2, (uint64_t)p.M * C * 2};
    const uint32_t box[4] = {64, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmX, x, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.HD, 1, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)p.HD * C * 2, (uint64_t)p.HD * C * 2};
    const uint32_t box[4] = {64, CW, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmW1, w1, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)p.HD, (uint64_t)C, 1, 1};
    const uint64_t str[3] = {(uint64_t)p.HD * 2, (uint64_t)p.HD * C * 2, (uint64_t)p.HD * C * 2};
    const uint32_t box[4] = {CW, (uint32_t)C, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmW2, w2, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.M, 1, 1};
    const uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)p.M * C * 4, (uint64_t)p.M * C * 4};
    const uint32_t box[4] = {32, 32, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmOut, out, dims, str, box, 1, 0)) != 0) return rc;
  }
  tmLn = tmOut;
  if (ln_out != nullptr) {   // bf16 [M, C] normalised rows: boxes of 32 rows x 64 columns (128-byte rows, SWIZZLE_128B)
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.M, 1, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)p.M * C * 2, (uint64_t)p.M * C * 2};
    const uint32_t box[4] = {64, 32, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmLn, ln_out, dims, str, box, 0, 0)) != 0) return rc;
  }
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(mlp_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_USED + 1024);
  });
  int grid = mvlt_num_sms();
  if (p.num_tiles < grid) grid = p.num_tiles;
  mvlt_launch(mlp_fwd_kernel<C>, grid, NTHREADS, (size_t)K::SMEM_USED + 1024, stream, tmLn, tmX, tmW1, tmW2, tmOut, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}


// =====================================================================================================================
// Backward of the fused MLP (C = 64): recomputes H = X W1^T + b1 instead of reading saved activations.
//
// A CTA owns a block of HB = 128 hidden units (W1 rows / W2 columns stay in shared memory) and a range of 128-row tiles;
// dW1 / dW2 / db1 of its block accumulate in TMEM over ALL its tiles and are flushed once with fp32 reductions. Per tile:
//
//   warp 0      : TMA: X tile, dY tile (double-buffered)
//   warp 1      : tcgen05.mma:  H_s  [128 x 64] = X  W1_s^T        (s = 0, 1: 64-column halves of the block)  -> TMEM
//                               dH_s [128 x 64] = dY W2[:, s]                                                  -> TMEM
//                 then, once the GELU warps have produced the two operand tiles:
//                               dW2_blk [128 hid x C] += act^T dY      (A = act read MN-major, B = dY read MN-major)
//                               dW1_blk [128 hid x C] += dh'^T X
//                               db1_blk               += dh'^T 1       (N = 16 MMA against a tile of ones)
//   warp 2      : TMA store of the dh' tile into dh'[M, HD] (consumed by the dX = dh' W1 GEMM that follows)
//   warps 4..19 : (g, g') = gelu, gelu'(H + b1);  act = g,  dh' = dH * g'  -> bf16 K-major SWIZZLE_128B tiles in shared memory
//                 (lane quarter x 16-column slice: 4 warps per SM sub-partition, the kernel is MUFU / issue bound)
//
// Replaces the autograd of /root/reference/libs/pvlt.py:65-71 (fc2 dX / dW / fc1 dW, db and the GELU backward); db2 and
// dX = dh' W1 are taken by the column-sum kernel and the tcgen05 GEMM (engine.py).
// =====================================================================================================================
constexpr int BW_THREADS = 640;      // warp 0 TMA, warp 1 MMA, warp 2 dh' store, warp 3 idle, warps 4..19 GELU / flush
constexpr int HB = 128;
constexpr int BW_OFF_W1 = 0;                        // [128 hid x 64 c] K-major
constexpr int BW_OFF_W2 = ATOM;                     // 2 x [64 c x 64 hid] (MN-major B operand)
constexpr int BW_OFF_T = 2 * ATOM;                  // 2 x { X | dY } tiles
constexpr int BW_OFF_ACT = BW_OFF_T + 4 * ATOM;     // 2 atoms [128 rows x 64 hid]
constexpr int BW_OFF_DH = BW_OFF_ACT + 2 * ATOM;
constexpr int BW_OFF_ONES = BW_OFF_DH + 2 * ATOM;
constexpr int BW_OFF_BAR = BW_OFF_ONES + 1024;
constexpr int BW_SMEM = BW_OFF_BAR + 128;
constexpr int COL_H = 0, COL_DH = 128, COL_DW1 = 256, COL_DW2 = 320, COL_ONES = 384;

struct MlpBwdParams {
  int M, HD;
  int num_tiles, nb, nr;    // 128-row tiles, hidden blocks, tile ranges (grid = nb * nr)
  const float* b1;
  float* dW1;               // fp32 [HD, 64], accumulated with reductions
  float* dW2;               // fp32 [64, HD]
  float* db1;               // fp32 [HD]
};

__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {   // MN-major SWIZZLE_128B operand
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ uint32_t instr_desc_mn(int n, int a_mn, int b_mn) {
  return instr_desc(n) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16);
}

__global__ void __launch_bounds__(BW_THREADS, 1)
mlp_bwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
               const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
               const __grid_constant__ CUtensorMap tmDH, const __grid_constant__ MlpBwdParams p) {
  constexpr int C = 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_OFF_BAR);
  uint64_t* t_full = bars;            // [2] X and dY tile landed
  uint64_t* t_empty = bars + 2;       // [2] the dW MMAs that read the tile retired
  uint64_t* hd_full = bars + 4;       // [2] H_s and dH_s complete
  uint64_t* hd_empty = bars + 6;      // [2] the eight GELU warps have read H_s / dH_s
  uint64_t* a_full = bars + 8;        // act and dh' tiles written (eight GELU warps)
  uint64_t* a_empty = bars + 9;       // the dW MMAs that read act / dh' retired
  uint64_t* st_empty = bars + 10;     // the dh' TMA store has finished reading its tile
  uint64_t* w_full = bars + 11;
  uint64_t* done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 1);
      mbar_init(&hd_full[i], 1);
      mbar_init(&hd_empty[i], NUM_GELU_WARPS);
    }
    mbar_init(a_full, NUM_GELU_WARPS);
    mbar_init(a_empty, 1);
    mbar_init(st_empty, 1);
    mbar_init(w_full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmDH);
  }
  if (threadIdx.x < 256) reinterpret_cast<uint32_t*>(smem + BW_OFF_ONES)[threadIdx.x] = 0x3F803F80u;   // bf16 1.0
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int jb = (int)blockIdx.x % p.nb, rg = (int)blockIdx.x / p.nb;
  int n_tiles = 0;
  if (rg < p.num_tiles) n_tiles = (p.num_tiles - 1 - rg) / p.nr + 1;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 2u * ATOM);
      tma_load_4d(smem + BW_OFF_W1, &tmW1, w_full, 0, jb * HB, 0, 0);
      tma_load_4d(smem + BW_OFF_W2, &tmW2, w_full, jb * HB, 0, 0, 0);
      tma_load_4d(smem + BW_OFF_W2 + 8192, &tmW2, w_full, jb * HB + 64, 0, 0, 0);
      for (int i = 0; i < n_tiles; ++i) {
        const int slot = i & 1, m0 = (rg + i * p.nr) * BM;
        mbar_wait(&t_empty[slot], (((uint32_t)i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&t_full[slot], 2u * ATOM);
        tma_load_4d(smem + BW_OFF_T + slot * 2 * ATOM, &tmX, &t_full[slot], 0, m0, 0, 0);
        tma_load_4d(smem + BW_OFF_T + slot * 2 * ATOM + ATOM, &tmDY, &t_full[slot], 0, m0, 0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id_kk = instr_desc_mn(64, 0, 0), id_kn = instr_desc_mn(64, 0, 1), id_mm = instr_desc_mn(C, 1, 1),
                     id_ones = instr_desc_mn(16, 1, 0);
      const uint32_t sW1 = smem_u32(smem + BW_OFF_W1), sW2 = smem_u32(smem + BW_OFF_W2), sT = smem_u32(smem + BW_OFF_T),
                     sAct = smem_u32(smem + BW_OFF_ACT), sDh = smem_u32(smem + BW_OFF_DH);
      const uint64_t ones_desc = smem_desc(smem_u32(smem + BW_OFF_ONES), 0u);   // SBO = 0: every 8-row group re-reads the same rows
      mbar_wait(w_full, 0);
      const uint32_t id_kk128 = instr_desc_mn(HB, 0, 0), id_kn128 = instr_desc_mn(HB, 0, 1);
      auto issue_hd = [&](int i) {
        const int slot = i & 1;
        const uint32_t xa = sT + (uint32_t)(slot * 2 * ATOM), dya = xa + (uint32_t)ATOM;
        mbar_wait(&t_full[slot], ((uint32_t)i >> 1) & 1u);
        mbar_wait(&hd_empty[0], ((uint32_t)i & 1u) ^ 1u);     // the GELU warps have read tile i - 1's H / dH
        tc_fence_after();
        // both 64-column halves in ONE N = 128 instruction per k-step (a 128 x 64 x 16 UMMA costs about as much as a
        // 128 x 128 x 16 one: 40 -> 32 instructions per tile)
#pragma unroll
        for (int k = 0; k < 4; ++k)    // H = X W1_blk^T: both K-major, B = the block's 128 hidden rows
          umma_bf16(tmem_base + (uint32_t)COL_H, smem_desc(xa + (uint32_t)(k * 32), 1024u), smem_desc(sW1 + (uint32_t)(k * 32), 1024u),
                    id_kk128, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)    // dH = dY W2[:, blk]: B = W2 rows (c) x 128 hidden columns, MN-major (two 64-wide atoms)
          umma_bf16(tmem_base + (uint32_t)COL_DH, smem_desc(dya + (uint32_t)(k * 32), 1024u),
                    smem_desc_mn(sW2 + (uint32_t)(k * 2048), 8192u), id_kn128, k > 0 ? 1u : 0u);
        umma_commit(&hd_full[0]);
      };
      if (n_tiles > 0) issue_hd(0);
      for (int i = 0; i < n_tiles; ++i) {
        if (i + 1 < n_tiles) issue_hd(i + 1);
        const int slot = i & 1;
        const uint32_t xa = sT + (uint32_t)(slot * 2 * ATOM), dya = xa + (uint32_t)ATOM;
        mbar_wait(a_full, (uint32_t)i & 1u);
        tc_fence_after();
        const uint32_t acc0 = i > 0 ? 1u : 0u;
#pragma unroll
        for (int kq = 0; kq < BM / 16; ++kq) {   // K = the tile's 128 rows, 16 per step; A tiles are read MN-major (M = hidden)
          const uint32_t acc = acc0 | (kq > 0 ? 1u : 0u);
          umma_bf16(tmem_base + COL_DW2, smem_desc_mn(sAct + (uint32_t)(kq * 2048), (uint32_t)ATOM),
                    smem_desc_mn(dya + (uint32_t)(kq * 2048), 8192u), id_mm, acc);
          umma_bf16(tmem_base + COL_DW1, smem_desc_mn(sDh + (uint32_t)(kq * 2048), (uint32_t)ATOM),
                    smem_desc_mn(xa + (uint32_t)(kq * 2048), 8192u), id_mm, acc);
          umma_bf16(tmem_base + COL_ONES, smem_desc_mn(sDh + (uint32_t)(kq * 2048), (uint32_t)ATOM), ones_desc, id_ones, acc);
        }
        umma_commit(a_empty);
        umma_commit(&t_empty[slot]);
      }
      umma_commit(done);
    }
  } else if (warp == 2) {
    if (lane == 0) {
      for (int i = 0; i < n_tiles; ++i) {
        const int m0 = (rg + i * p.nr) * BM;
        mbar_wait(a_full, (uint32_t)i & 1u);     // (the writers fenced their generic-proxy stores before arriving)
        tma_store_4d(&tmDH, smem_u32(smem + BW_OFF_DH), jb * HB, m0, 0, 0);
        tma_store_4d(&tmDH, smem_u32(smem + BW_OFF_DH + ATOM), jb * HB + 64, m0, 0, 0);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(st_empty);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp >= GELU_WARP0) {
    const int quarter = warp & 3, cq = (warp - GELU_WARP0) >> 2;     // lane quarter, 16-column slice of a 64-column half
    const uint32_t tlane = ((uint32_t)(quarter * 32) << 16);
    const uint32_t row = (uint32_t)(quarter * 32 + lane), rx = row & 7u;
    const uint32_t act_row = smem_u32(smem + BW_OFF_ACT) + row * 128u, dh_row = smem_u32(smem + BW_OFF_DH) + row * 128u;
    for (int i = 0; i < n_tiles; ++i) {
      const uint32_t ph = (uint32_t)i & 1u;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        const float4* bp = reinterpret_cast<const float4*>(p.b1 + jb * HB + s * 64 + cq * 16);
        float4 bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bv[q] = __ldg(bp + q);
        if (s == 0) {
          mbar_wait(&hd_full[0], ph);
          tc_fence_after();
        }
        uint32_t rh[16], rd[16];
        tmem_ld_32x16(tmem_base + tlane + (uint32_t)(COL_H + s * 64 + cq * 16), rh);
        tmem_ld_32x16(tmem_base + tlane + (uint32_t)(COL_DH + s * 64 + cq * 16), rd);
        tmem_ld_wait();
        if (s == 1) {
          tc_fence_before();     // this warp's TMEM reads precede the MMAs of the next tile into H / dH
          __syncwarp();
          if (lane == 0) mbar_arrive(&hd_empty[0]);
        }
        uint32_t apk[8], dpk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          f32x2_t g0, d0, g1, d1;
          gelu_and_grad2(f2_add(f2_pack(__uint_as_float(rh[4 * q]), __uint_as_float(rh[4 * q + 1])), f2_pack(bv[q].x, bv[q].y)), g0, d0);
          gelu_and_grad2(f2_add(f2_pack(__uint_as_float(rh[4 * q + 2]), __uint_as_float(rh[4 * q + 3])), f2_pack(bv[q].z, bv[q].w)), g1, d1);
          d0 = f2_mul(d0, f2_pack(__uint_as_float(rd[4 * q]), __uint_as_float(rd[4 * q + 1])));
          d1 = f2_mul(d1, f2_pack(__uint_as_float(rd[4 * q + 2]), __uint_as_float(rd[4 * q + 3])));
          float a, b;
          f2_unpack(g0, a, b);
          apk[2 * q] = pack_bf16x2(a, b);
          f2_unpack(g1, a, b);
          apk[2 * q + 1] = pack_bf16x2(a, b);
          f2_unpack(d0, a, b);
          dpk[2 * q] = pack_bf16x2(a, b);
          f2_unpack(d1, a, b);
          dpk[2 * q + 1] = pack_bf16x2(a, b);
        }
        if (s == 0) {            // the previous tile's dW MMAs and dh' store have finished reading the operand tiles
          mbar_wait(a_empty, ph ^ 1u);
          mbar_wait(st_empty, ph ^ 1u);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t off = (uint32_t)(s * ATOM) + ((((uint32_t)(cq * 2 + q)) ^ rx) << 4);
          st_shared_v4(act_row + off, apk[4 * q], apk[4 * q + 1], apk[4 * q + 2], apk[4 * q + 3]);
          st_shared_v4(dh_row + off, dpk[4 * q], dpk[4 * q + 1], dpk[4 * q + 2], dpk[4 * q + 3]);
        }
      }
      fence_proxy_async();       // generic-proxy stores -> visible to the tensor core and the TMA engine
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
    }
    // ---- flush the block's accumulators: dW1[hid, c], dW2[c, hid], db1[hid] (fp32 reductions; lane = hidden unit)
    mbar_wait(done, 0);
    tc_fence_after();
    if (n_tiles > 0) {
      const int hid = jb * HB + quarter * 32 + lane;
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(COL_DW1 + cq * 16), r);
      tmem_ld_wait();
      float* g1 = p.dW1 + (long long)hid * C + cq * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g1 + 4 * q), "f"(__uint_as_float(r[4 * q])),
                     "f"(__uint_as_float(r[4 * q + 1])), "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3]))
                     : "memory");
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(COL_DW2 + cq * 16), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.dW2 + (long long)(cq * 16 + j) * p.HD + hid, __uint_as_float(r[j]));
      if (cq == 0) {
        tmem_ld_32x16(tmem_base + tlane + (uint32_t)COL_ONES, r);
        tmem_ld_wait();
        atomicAdd(p.db1 + hid, __uint_as_float(r[0]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// =====================================================================================================================
// Backward of the fused MLP for C = 128 (PVLT stage 2). Same structure as mlp_bwd_kernel, re-proportioned so that the
// accumulators fit TMEM and two tile slots fit shared memory: a CTA owns HB2 = 64 hidden units, and the weight-gradient
// accumulators are kept TRANSPOSED (rows = the 128 input channels = TMEM lanes, columns = the 64 hidden units):
//     H, dH [128 x 64]      as before (K = C = 128: 8 k-steps)
//     dW1^T [128 c x 64 hid] += X^T dh'       (A = X tile read MN-major, B = dh' tile read MN-major)
//     dW2   [128 c x 64 hid] += dY^T act      (the [C, HD] layout of fc2.weight's gradient directly)
//     db1   (every row)      += 1^T dh'       (A = a tile of ones)
// =====================================================================================================================
constexpr int HB2 = 64;
constexpr int B2_OFF_W1 = 0;                          // [64 hid x 128 c] K-major: 2 atoms of [64 rows x 128 B]
constexpr int B2_OFF_W2 = ATOM;                       // [128 c x 64 hid]: MN-major B operand of dH
constexpr int B2_OFF_T = 2 * ATOM;                    // 2 x { X (2 atoms) | dY (2 atoms) }
constexpr int B2_OFF_ACT = B2_OFF_T + 8 * ATOM;       // [128 rows x 64 hid]
constexpr int B2_OFF_DH = B2_OFF_ACT + ATOM;
constexpr int B2_OFF_ONES = B2_OFF_DH + ATOM;
constexpr int B2_OFF_BAR = B2_OFF_ONES + 1024;
constexpr int B2_SMEM = B2_OFF_BAR + 128;
constexpr int C2_H = 0, C2_DH = 64, C2_DW1 = 128, C2_DW2 = 192, C2_DB = 256;
static_assert(B2_SMEM + 1024 <= 232448, "shared memory plan (C = 128 backward)");

__global__ void __launch_bounds__(BW_THREADS, 1)
mlp_bwd128_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmDH, const __grid_constant__ MlpBwdParams p) {
  constexpr int C = 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2_OFF_BAR);
  uint64_t* t_full = bars;            // [2]
  uint64_t* t_empty = bars + 2;       // [2]
  uint64_t* hd_full = bars + 4;
  uint64_t* hd_empty = bars + 5;
  uint64_t* a_full = bars + 6;
  uint64_t* a_empty = bars + 7;
  uint64_t* st_empty = bars + 8;
  uint64_t* w_full = bars + 9;
  uint64_t* done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 1);
    }
    mbar_init(hd_full, 1);
    mbar_init(hd_empty, NUM_GELU_WARPS);
    mbar_init(a_full, NUM_GELU_WARPS);
    mbar_init(a_empty, 1);
    mbar_init(st_empty, 1);
    mbar_init(w_full, 1);
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmDH);
  }
  if (threadIdx.x < 256) reinterpret_cast<uint32_t*>(smem + B2_OFF_ONES)[threadIdx.x] = 0x3F803F80u;   // bf16 1.0
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int jb = (int)blockIdx.x % p.nb, rg = (int)blockIdx.x / p.nb;
  int n_tiles = 0;
  if (rg < p.num_tiles) n_tiles = (p.num_tiles - 1 - rg) / p.nr + 1;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, 2u * ATOM);
      tma_load_4d(smem + B2_OFF_W1, &tmW1, w_full, 0, jb * HB2, 0, 0);
      tma_load_4d(smem + B2_OFF_W1 + 8192, &tmW1, w_full, 64, jb * HB2, 0, 0);
      tma_load_4d(smem + B2_OFF_W2, &tmW2, w_full, jb * HB2, 0, 0, 0);
      for (int i = 0; i < n_tiles; ++i) {
        const int slot = i & 1, m0 = (rg + i * p.nr) * BM;
        uint8_t* t = smem + B2_OFF_T + slot * 4 * ATOM;
        mbar_wait(&t_empty[slot], (((uint32_t)i >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&t_full[slot], 4u * ATOM);
        tma_load_4d(t, &tmX, &t_full[slot], 0, m0, 0, 0);
        tma_load_4d(t + ATOM, &tmX, &t_full[slot], 64, m0, 0, 0);
        tma_load_4d(t + 2 * ATOM, &tmDY, &t_full[slot], 0, m0, 0, 0);
        tma_load_4d(t + 3 * ATOM, &tmDY, &t_full[slot], 64, m0, 0, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id_kk = instr_desc_mn(HB2, 0, 0), id_kn = instr_desc_mn(HB2, 0, 1), id_mm = instr_desc_mn(HB2, 1, 1);
      const uint32_t sW1 = smem_u32(smem + B2_OFF_W1), sW2 = smem_u32(smem + B2_OFF_W2), sT = smem_u32(smem + B2_OFF_T),
                     sAct = smem_u32(smem + B2_OFF_ACT), sDh = smem_u32(smem + B2_OFF_DH);
      const uint64_t ones_desc = smem_desc(smem_u32(smem + B2_OFF_ONES), 0u);   // LBO = SBO = 0: every atom aliases 1 KB of ones
      mbar_wait(w_full, 0);
      auto issue_hd = [&](int i) {
        const int slot = i & 1;
        const uint32_t xa = sT + (uint32_t)(slot * 4 * ATOM), dya = xa + 2u * ATOM;
        mbar_wait(&t_full[slot], ((uint32_t)i >> 1) & 1u);
        mbar_wait(hd_empty, ((uint32_t)i & 1u) ^ 1u);     // the GELU warps have read tile i - 1's H / dH
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < C / 16; ++k)    // H = X W1_blk^T: both K-major, K = 128 over two atoms
          umma_bf16(tmem_base + C2_H, smem_desc(xa + (uint32_t)((k >> 2) * ATOM + (k & 3) * 32), 1024u),
                    smem_desc(sW1 + (uint32_t)((k >> 2) * 8192 + (k & 3) * 32), 1024u), id_kk, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < C / 16; ++k)    // dH = dY W2[:, blk]: B = 128 rows (c) x 64 hidden columns, MN-major
          umma_bf16(tmem_base + C2_DH, smem_desc(dya + (uint32_t)((k >> 2) * ATOM + (k & 3) * 32), 1024u),
                    smem_desc_mn(sW2 + (uint32_t)(k * 2048), 8192u), id_kn, k > 0 ? 1u : 0u);
        umma_commit(hd_full);
      };
      if (n_tiles > 0) issue_hd(0);
      for (int i = 0; i < n_tiles; ++i) {
        if (i + 1 < n_tiles) issue_hd(i + 1);
        const int slot = i & 1;
        const uint32_t xa = sT + (uint32_t)(slot * 4 * ATOM), dya = xa + 2u * ATOM;
        mbar_wait(a_full, (uint32_t)i & 1u);
        tc_fence_after();
        const uint32_t acc0 = i > 0 ? 1u : 0u;
#pragma unroll
        for (int kq = 0; kq < BM / 16; ++kq) {   // K = the tile's 128 rows; A tiles read MN-major (M = the 128 channels)
          const uint32_t acc = acc0 | (kq > 0 ? 1u : 0u);
          const uint64_t bdh = smem_desc_mn(sDh + (uint32_t)(kq * 2048), 8192u);
          umma_bf16(tmem_base + C2_DW1, smem_desc_mn(xa + (uint32_t)(kq * 2048), (uint32_t)ATOM), bdh, id_mm, acc);
          umma_bf16(tmem_base + C2_DW2, smem_desc_mn(dya + (uint32_t)(kq * 2048), (uint32_t)ATOM),
                    smem_desc_mn(sAct + (uint32_t)(kq * 2048), 8192u), id_mm, acc);
          umma_bf16(tmem_base + C2_DB, ones_desc, bdh, id_mm, acc);
        }
        umma_commit(a_empty);
        umma_commit(&t_empty[slot]);
      }
      umma_commit(done);
    }
  } else if (warp == 2) {
    if (lane == 0) {
      for (int i = 0; i < n_tiles; ++i) {
        const int m0 = (rg + i * p.nr) * BM;
        mbar_wait(a_full, (uint32_t)i & 1u);
        tma_store_4d(&tmDH, smem_u32(smem + B2_OFF_DH), jb * HB2, m0, 0, 0);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(st_empty);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp >= GELU_WARP0) {
    const int quarter = warp & 3, cq = (warp - GELU_WARP0) >> 2;
    const uint32_t tlane = ((uint32_t)(quarter * 32) << 16);
    const uint32_t row = (uint32_t)(quarter * 32 + lane), rx = row & 7u;
    const uint32_t act_row = smem_u32(smem + B2_OFF_ACT) + row * 128u, dh_row = smem_u32(smem + B2_OFF_DH) + row * 128u;
    const float4* bp = reinterpret_cast<const float4*>(p.b1 + jb * HB2 + cq * 16);
    float4 bv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bv[q] = __ldg(bp + q);      // the block's bias slice never changes: loaded once
    for (int i = 0; i < n_tiles; ++i) {
      const uint32_t ph = (uint32_t)i & 1u;
      mbar_wait(hd_full, ph);
      tc_fence_after();
      uint32_t rh[16], rd[16];
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(C2_H + cq * 16), rh);
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(C2_DH + cq * 16), rd);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(hd_empty);
      uint32_t apk[8], dpk[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        f32x2_t g0, d0, g1, d1;
        gelu_and_grad2(f2_add(f2_pack(__uint_as_float(rh[4 * q]), __uint_as_float(rh[4 * q + 1])), f2_pack(bv[q].x, bv[q].y)), g0, d0);
        gelu_and_grad2(f2_add(f2_pack(__uint_as_float(rh[4 * q + 2]), __uint_as_float(rh[4 * q + 3])), f2_pack(bv[q].z, bv[q].w)), g1, d1);
        d0 = f2_mul(d0, f2_pack(__uint_as_float(rd[4 * q]), __uint_as_float(rd[4 * q + 1])));
        d1 = f2_mul(d1, f2_pack(__uint_as_float(rd[4 * q + 2]), __uint_as_float(rd[4 * q + 3])));
        float a, b;
        f2_unpack(g0, a, b);
        apk[2 * q] = pack_bf16x2(a, b);
        f2_unpack(g1, a, b);
        apk[2 * q + 1] = pack_bf16x2(a, b);
        f2_unpack(d0, a, b);
        dpk[2 * q] = pack_bf16x2(a, b);
        f2_unpack(d1, a, b);
        dpk[2 * q + 1] = pack_bf16x2(a, b);
      }
      mbar_wait(a_empty, ph ^ 1u);      // the previous tile's dW MMAs and dh' store have finished reading the operand tiles
      mbar_wait(st_empty, ph ^ 1u);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t off = ((((uint32_t)(cq * 2 + q)) ^ rx) << 4);
        st_shared_v4(act_row + off, apk[4 * q], apk[4 * q + 1], apk[4 * q + 2], apk[4 * q + 3]);
        st_shared_v4(dh_row + off, dpk[4 * q], dpk[4 * q + 1], dpk[4 * q + 2], dpk[4 * q + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
    }
    // ---- flush: lane = input channel c, columns = the block's hidden units
    mbar_wait(done, 0);
    tc_fence_after();
    if (n_tiles > 0) {
      const int c = quarter * 32 + lane, h0 = jb * HB2 + cq * 16;
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(C2_DW1 + cq * 16), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) atomicAdd(p.dW1 + (long long)(h0 + j) * C + c, __uint_as_float(r[j]));   // dW1[hid, c]
      tmem_ld_32x16(tmem_base + tlane + (uint32_t)(C2_DW2 + cq * 16), r);
      tmem_ld_wait();
      float* g2 = p.dW2 + (long long)c * p.HD + h0;                                                        // dW2[c, hid]
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g2 + 4 * q), "f"(__uint_as_float(r[4 * q])),
                     "f"(__uint_as_float(r[4 * q + 1])), "f"(__uint_as_float(r[4 * q + 2])), "f"(__uint_as_float(r[4 * q + 3]))
                     : "memory");
      if (quarter == 0) {     // every row of the db accumulator holds the same column sums: row 0 writes them
        tmem_ld_32x16(tmem_base + tlane + (uint32_t)(C2_DB + cq * 16), r);
        tmem_ld_wait();
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(p.db1 + h0 + j, __uint_as_float(r[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// out[M, C] (fp32) = residual[M, C] (fp32) + rowscale[row / rows_per_scale] * (GELU(x W1^T + b1) W2^T + b2)
//   x_bf16 [M, C], w1_bf16 [HD, C], b1 fp32 [HD], w2_bf16 [C, HD], b2 fp32 [C]; all contiguous, 16-byte aligned.
//   C in {64, 128}, HD a multiple of 64; rowscale_f32 may be null (no DropPath). ``out`` may alias ``residual``.
// ln_* (optional, all NULL / 0 for none): LayerNorm of the output rows fused into the same launch -- ln_out_bf16 [M, C] =
// (out - mean) * rstd * ln_gamma + ln_beta (biased variance over the C columns, rstd = rsqrt(var + ln_eps)), ln_mean / ln_rstd
// fp32 [M] (optional) for the LayerNorm backward: the next block's norm1 without re-reading the fp32 rows
extern "C" int mvlt_mlp_fwd(const void* x_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                            const float* residual_f32, float* out_f32, const float* rowscale_f32, int rows_per_scale, int M,
                            int C, int HD, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, float* ln_mean,
                            float* ln_rstd, float ln_eps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(x_bf16 && w1_bf16 && b1 && w2_bf16 && b2 && residual_f32 && out_f32, "mlp_fwd: null operand");
  MVLT_CHECK_ARG(M > 0 && (C == 64 || C == 128) && HD >= 64 && HD % 64 == 0, "mlp_fwd: unsupported shape M=%d C=%d HD=%d", M, C, HD);
  MVLT_CHECK_ARG(rowscale_f32 == nullptr || rows_per_scale > 0, "mlp_fwd: rows_per_scale must be positive");
  MVLT_CHECK_ARG(((((uintptr_t)x_bf16) | ((uintptr_t)w1_bf16) | ((uintptr_t)w2_bf16) | ((uintptr_t)residual_f32) | ((uintptr_t)out_f32) |
                   ((uintptr_t)b1) | ((uintptr_t)b2)) & 15) == 0, "mlp_fwd: operands must be 16-byte aligned");
  MlpParams p;
  p.M = M; p.C = C; p.HD = HD;
  p.num_tiles = (M + BM - 1) / BM;
  p.nch = HD / CW;
  p.b1 = b1; p.b2 = b2; p.residual = residual_f32; p.rowscale = rowscale_f32; p.rows_per_scale = rows_per_scale > 0 ? rows_per_scale : 1;
  static const int dbg = [] { const char* e = getenv("MVLT_MLP_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  MVLT_CHECK_ARG((ln_gamma == nullptr) == (ln_out_bf16 == nullptr) && (ln_gamma == nullptr) == (ln_beta == nullptr),
                 "mlp_fwd: ln_gamma, ln_beta and ln_out go together");
  MVLT_CHECK_ARG(((((uintptr_t)ln_gamma) | ((uintptr_t)ln_beta) | ((uintptr_t)ln_out_bf16)) & 15) == 0, "mlp_fwd: ln operands must be 16-byte aligned");
  MVLT_CHECK_ARG(ln_gamma == nullptr || (const void*)out_f32 != (const void*)residual_f32,
                 "mlp_fwd: the fused LayerNorm re-reads the residual after the output is stored: out must not alias it");
  p.ln_gamma = ln_gamma; p.ln_beta = ln_beta; p.ln_mean = ln_mean; p.ln_rstd = ln_rstd; p.ln_eps = ln_eps;
  return C == 64 ? launch_fwd<64>(x_bf16, w1_bf16, w2_bf16, out_f32, ln_out_bf16, p, stream)
                 : launch_fwd<128>(x_bf16, w1_bf16, w2_bf16, out_f32, ln_out_bf16, p, stream);
}


// Backward of mvlt_mlp_fwd's branch (see mlp_bwd_kernel / mlp_bwd128_kernel): given x_bf16 [M, C] (the forward's input), dy_bf16
// [M, C] (gradient of the branch output, DropPath factor already applied), w1_bf16 [HD, C], b1, w2_bf16 [C, HD], C in {64, 128}:
//   dh_bf16 [M, HD]  = (dy W2) * gelu'(x W1^T + b1)            (written; feeds dX = dh W1 and nothing else)
//   dW1 [HD, C] += dh^T x,  dW2 [C, HD] += dy^T gelu(x W1^T + b1),  db1 [HD] += column sums of dh    (fp32, accumulated)
// HD must be a multiple of 128. All pointers 16-byte aligned, tensors contiguous.
extern "C" int mvlt_mlp_bwd(const void* x_bf16, const void* dy_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16,
                            void* dh_bf16, float* dW1_f32, float* dW2_f32, float* db1_f32, int M, int C, int HD, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(x_bf16 && dy_bf16 && w1_bf16 && b1 && w2_bf16 && dh_bf16 && dW1_f32 && dW2_f32 && db1_f32, "mlp_bwd: null operand");
  MVLT_CHECK_ARG(M > 0 && (C == 64 || C == 128) && HD >= HB && HD % HB == 0,
                 "mlp_bwd: unsupported shape M=%d C=%d HD=%d (C in {64, 128}, HD %% 128 == 0)", M, C, HD);
  MVLT_CHECK_ARG(((((uintptr_t)x_bf16) | ((uintptr_t)dy_bf16) | ((uintptr_t)w1_bf16) | ((uintptr_t)w2_bf16) | ((uintptr_t)dh_bf16) |
                   ((uintptr_t)dW1_f32) | ((uintptr_t)dW2_f32) | ((uintptr_t)b1)) & 15) == 0, "mlp_bwd: operands must be 16-byte aligned");
  const int hb = C == 64 ? HB : HB2;      // hidden units per CTA
  MlpBwdParams p;
  p.M = M; p.HD = HD;
  p.num_tiles = (M + BM - 1) / BM;
  p.nb = HD / hb;
  p.nr = mvlt_num_sms() / p.nb;
  if (p.nr < 1) p.nr = 1;
  if (p.nr > p.num_tiles) p.nr = p.num_tiles;
  p.b1 = b1; p.dW1 = dW1_f32; p.dW2 = dW2_f32; p.db1 = db1_f32;
  CUtensorMap tmX, tmDY, tmW1, tmW2, tmDH;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)M, 1, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)M * C * 2, (uint64_t)M * C * 2};
    const uint32_t box[4] = {64, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmX, x_bf16, dims, str, box, 0, 0)) != 0) return rc;
    if ((rc = mvlt_tensor_map_4d(&tmDY, dy_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)HD, 1, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)HD * C * 2, (uint64_t)HD * C * 2};
    const uint32_t box[4] = {64, (uint32_t)hb, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmW1, w1_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)C, 1, 1};
    const uint64_t str[3] = {(uint64_t)HD * 2, (uint64_t)HD * C * 2, (uint64_t)HD * C * 2};
    const uint32_t box[4] = {64, (uint32_t)C, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmW2, w2_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)M, 1, 1};
    const uint64_t str[3] = {(uint64_t)HD * 2, (uint64_t)M * HD * 2, (uint64_t)M * HD * 2};
    const uint32_t box[4] = {64, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmDH, dh_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM + 1024);
    cudaFuncSetAttribute(mlp_bwd128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM + 1024);
  });
  if (C == 64) mvlt_launch(mlp_bwd_kernel, p.nb * p.nr, BW_THREADS, (size_t)BW_SMEM + 1024, stream, tmX, tmDY, tmW1, tmW2, tmDH, p);
  else mvlt_launch(mlp_bwd128_kernel, p.nb * p.nr, BW_THREADS, (size_t)B2_SMEM + 1024, stream, tmX, tmDY, tmW1, tmW2, tmDH, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}
