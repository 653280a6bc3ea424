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
//   warps 4..11 : GELU warps (lane quarter x column half): H_g (tcgen05.ld) + b1 -> exact-erf GELU (packed fp32x2) -> bf16
//                 -> A_g in shared memory, written straight into the K-major SWIZZLE_128B layout GEMM2 reads
//   warps 12..15: output warps: Y + b2, x drop-path factor, + fp32 residual -> 4 KB swizzled staging tile -> TMA store
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
constexpr int NTHREADS = 512;
constexpr int GELU_WARP0 = 4, NUM_GELU_WARPS = 8, OUT_WARP0 = 12, NUM_OUT_WARPS = 4;
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
mlp_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
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
      for (int g = 0; g < total + LAG; ++g) {
        if (g < total) {                          // ---- GEMM1 of chunk g
          const int tl = g / nch, j = g - tl * nch;
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
          const int tl = g2 / nch, j = g2 - tl * nch;
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
    const int half = (warp - GELU_WARP0) >> 2;    // 32-column half of the 64-column chunk
    const uint32_t tlane = ((uint32_t)(quarter * 32) << 16);
    const uint32_t row = (uint32_t)(quarter * 32 + lane);
    const uint32_t a_row = smem_u32(smem + K::OFF_A) + row * 128u;
    const uint32_t rx = row & 7u;
    for (int g = 0; g < total; ++g) {
      const int hb = g % NH;
      const uint32_t ph = ((uint32_t)(g / NH)) & 1u;
      const int j = g % nch;
      // bias of this chunk half: identical for every lane (L1 broadcast), fetched before the accumulator is waited for
      float4 bv[8];
      const float4* bp = reinterpret_cast<const float4*>(p.b1 + j * CW + half * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) bv[i] = __ldg(bp + i);
      mbar_wait(&h_full[hb], ph);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + tlane + (uint32_t)(hb * CW + half * 32), r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
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
      for (int q = 0; q < 4; ++q)
        st_shared_v4(base + ((((uint32_t)(half * 4 + q)) ^ rx) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async();     // generic-proxy stores -> visible to the tensor core
      tc_fence_before();       // this warp's TMEM reads precede the MMA that will overwrite the H buffer
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[hb]);
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
#pragma unroll 1
      for (int u = 0; u < C / 32; ++u) {
        float4 rv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = valid ? __ldg(reinterpret_cast<const float4*>(res_row + u * 32) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (u == 0) {
          mbar_wait(&y_full[yb], ((uint32_t)tl >> 1) & 1u);
          tc_fence_after();
        }
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tlane + (uint32_t)(Y_COL0 + yb * C + u * 32), r);
        tmem_ld_wait();
        if (lane == 0) tma_store_wait_read();     // the previous unit's store has finished reading the staging tile
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + u * 32) + i);
          const float o0 = fmaf(__uint_as_float(r[4 * i]) + b4.x, rs, rv[i].x);
          const float o1 = fmaf(__uint_as_float(r[4 * i + 1]) + b4.y, rs, rv[i].y);
          const float o2 = fmaf(__uint_as_float(r[4 * i + 2]) + b4.z, rs, rv[i].z);
          const float o3 = fmaf(__uint_as_float(r[4 * i + 3]) + b4.w, rs, rv[i].w);
          st_shared_v4(own + ((((uint32_t)i) ^ rx) << 4), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmOut, tile_s, u * 32, m0 + quarter * 32, 0, 0);   // rows past M are clipped by the tensor map
          tma_store_commit();
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
int launch_fwd(const void* x, const void* w1, const void* w2, void* out, const MlpParams& p, cudaStream_t stream) {
  using K = Cfg<C>;
  CUtensorMap tmX, tmW1, tmW2, tmOut;
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
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(mlp_fwd_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_USED + 1024);
  });
  int grid = mvlt_num_sms();
  if (p.num_tiles < grid) grid = p.num_tiles;
  mvlt_launch(mlp_fwd_kernel<C>, grid, NTHREADS, (size_t)K::SMEM_USED + 1024, stream, tmX, tmW1, tmW2, tmOut, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}

}  // namespace

// out[M, C] (fp32) = residual[M, C] (fp32) + rowscale[row / rows_per_scale] * (GELU(x W1^T + b1) W2^T + b2)
//   x_bf16 [M, C], w1_bf16 [HD, C], b1 fp32 [HD], w2_bf16 [C, HD], b2 fp32 [C]; all contiguous, 16-byte aligned.
//   C in {64, 128}, HD a multiple of 64; rowscale_f32 may be null (no DropPath). ``out`` may alias ``residual``.
extern "C" int mvlt_mlp_fwd(const void* x_bf16, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                            const float* residual_f32, float* out_f32, const float* rowscale_f32, int rows_per_scale, int M,
                            int C, int HD, void* stream_) {
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
  return C == 64 ? launch_fwd<64>(x_bf16, w1_bf16, w2_bf16, out_f32, p, stream) : launch_fwd<128>(x_bf16, w1_bf16, w2_bf16, out_f32, p, stream);
}
