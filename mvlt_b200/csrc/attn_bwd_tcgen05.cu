// Fused spatial-reduction attention BACKWARD for sm_100a (B200): dQ, dK, dV of O = softmax(scale * Q K^T) V in ONE
// kernel, from the probabilities P the forward kernel saved (attn_tcgen05.cu).
//
// Default backward of the attention core (engine.py:_block_bwd; MVLT_FUSED_ATTN_BWD=0 selects the four-GEMM path for A/B
// runs). Validated against fp32 autograd in tests/test_attention_gpu.py and through the model parity tests.
//
// One CTA (256 threads = 8 warps, 1 per SM: it owns all 512 TMEM columns) walks the (batch, head) strips assigned to it; a
// query row is shared by two threads (key-column halves), which halves the latency of the per-row softmax backward. For a
// strip, K and V ([Nk x 64] each) stay in shared memory and dK / dV accumulate in TMEM across the strip's 128-row query
// tiles; per tile (operands double-buffered, the next tile's TMA loads run under the current tile):
//
//   MMA-A  dP[128 x Nk]  = dO V^T           A = dO (K-major), B = V (K-major rows = keys)            -> TMEM [0, Nk)
//   MMA-C  dV[keys x 64] += P^T dO          A = P tile read MN-major (M = keys), B = dO MN-major     -> TMEM [192, 320)
//   threads (row r): delta = sum_j P_rj dP_rj;  dS_rj = scale P_rj (dP_rj - delta), written bf16 IN PLACE over P
//   MMA-B  dQ[128 x 64]  = dS K             A = dS (K-major atoms), B = K MN-major                    -> TMEM [0, 64)
//   MMA-D  dK[keys x 64] += dS^T Q          A = dS tile read MN-major, B = Q MN-major                 -> TMEM [320, 448)
//   dQ epilogue: TMEM -> bf16 -> swizzled tile (the dO buffer, free by then) -> one TMA store
//
// The key axis of dK / dV is covered by two M = 128 accumulator blocks: keys 0..127 and keys 64..191 (atoms 1 and 2 of
// the P / dS tile), so no M = 64 instruction shape and no out-of-range operand address is needed; rows 0..63 of the
// second block duplicate keys 64..127 and are simply stored again with the same values.
//
// Replaces the autograd of /root/reference/libs/pvlt.py:113-117 (the four GEMMs of engine.py:_block_bwd around P).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

int mvlt_tensor_map_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_b[3],
                       const uint32_t box[4], int f32, int swizzle64);   // gemm_tcgen05.cu

namespace {

constexpr int BM = 128, HD = 64, NK_MAX = 192;
constexpr int TILE = BM * HD * 2;            // 16 KB: a [128 x 64] bf16 tile (Q, dO, dQ staging, one P / dS atom)
constexpr int KV_BYTES = NK_MAX * HD * 2;    // 24 KB
constexpr int P_BYTES = 3 * TILE;            // 3 atoms of 64 keys
// shared memory: K | V | 2 x { dO | Q | P } | barriers
constexpr int OFF_K = 0, OFF_V = KV_BYTES, OFF_T = 2 * KV_BYTES;
constexpr int T_DO = 0, T_Q = TILE, T_P = 2 * TILE, T_BYTES = 2 * TILE + P_BYTES;   // 80 KB per tile slot
constexpr int OFF_BAR = OFF_T + 2 * T_BYTES;
constexpr int OFF_XCH = OFF_BAR + 128;        // [2][128] floats: partial row dot products exchanged between the column halves
constexpr int SMEM_USED = OFF_XCH + 1024;
constexpr int TMEM_COLS = 512;
constexpr int COL_DV = 192, COL_DK = 320;   // two 64-column blocks each
constexpr int THREADS = 256;   // 8 warps: (TMEM lane quarter) x (key-column half); thread 0 also issues TMA and tcgen05.mma

struct BwdParams {
  int B, heads, N, Nk, C;
  int num_m, strips;
  float scale;
};

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
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= (uint32_t)(a_mn & 1) << 15;
  d |= (uint32_t)(b_mn & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(BM >> 4) << 24;
  return d;
}

__global__ void __launch_bounds__(THREADS, 1)
sr_attention_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                        const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmP,
                        const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdKV,
                        const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_bar = bars;          // K, V of the strip have landed
  uint64_t* full_bar = bars + 1;    // [2] dO, Q, P of a tile slot have landed
  uint64_t* s_bar = bars + 3;       // dP complete (MMA-A)
  uint64_t* c_bar = bars + 4;       // dV accumulation of this tile complete (MMA-C): P may be overwritten with dS
  uint64_t* q_bar = bars + 5;       // dQ complete (MMA-B)
  uint64_t* d_bar = bars + 6;       // dK accumulation complete (MMA-D): the tile slot may be refilled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  if (tid == 0) {
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmP);
    tma_prefetch_desc(&tmdQ);
    tma_prefetch_desc(&tmdKV);
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int quarter = warp & 3, half = warp >> 2;   // TMEM lane quarter this warp may access (warp id % 4), key-column half
  const int row = quarter * 32 + (tid & 31);        // query row (TMEM lane) of this thread; two threads share a row
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
  pdl_wait();

  const uint32_t sK_s = smem_u32(smem + OFF_K), sV_s = smem_u32(smem + OFF_V), sT_s = smem_u32(smem + OFF_T);
  const int Nk = p.Nk;
  const int nchunk16 = Nk >> 4;
  const uint32_t kv_bytes = (uint32_t)Nk * HD * 2;
  const uint32_t tile_tx = 2u * TILE + (uint32_t)P_BYTES;   // dO + Q + 3 P atoms (out-of-range atoms are zero-filled)
  const uint32_t idesc_dp = instr_desc(Nk, 0, 0);   // dP = dO V^T
  const uint32_t idesc_dq = instr_desc(HD, 0, 1);   // dQ = dS K
  const uint32_t idesc_kv = instr_desc(HD, 1, 1);   // dV += P^T dO, dK += dS^T Q
  const uint32_t row_s = (uint32_t)row * 128u, row_x = (uint32_t)(row & 7);
  const int nc_half = nchunk16 >> 1, c_lo = half * nc_half, c_hi = c_lo + nc_half;   // this thread's 16-column chunks
  const float scale = p.scale;

  uint32_t it = 0;        // tiles processed by this CTA (barrier phases)
  uint32_t strip_it = 0;  // strips processed by this CTA

  auto load_tile = [&](int slot, int b, int h, int m0) {   // thread 0 only
    uint8_t* base = smem + OFF_T + slot * T_BYTES;
    mbar_arrive_expect_tx(&full_bar[slot], tile_tx);
    tma_load_4d(base + T_DO, &tmdO, &full_bar[slot], h * HD, m0, b, 0);
    tma_load_4d(base + T_Q, &tmQ, &full_bar[slot], h * HD, m0, b, 0);
#pragma unroll
    for (int a = 0; a < 3; ++a) tma_load_4d(base + T_P + a * TILE, &tmP, &full_bar[slot], a * 64, m0, h, b);
  };

  for (int strip = blockIdx.x; strip < p.strips; strip += gridDim.x, ++strip_it) {
    const int b = strip / p.heads, h = strip - b * p.heads;
    if (tid == 0) {
      // every MMA of the previous strip has retired (d_bar / c_bar were waited below) and its stores have been read
      mbar_arrive_expect_tx(kv_bar, 2u * kv_bytes);
      tma_load_4d(smem + OFF_K, &tmKV, kv_bar, h * HD, 0, b, 0);
      tma_load_4d(smem + OFF_V, &tmKV, kv_bar, p.C + h * HD, 0, b, 0);
      load_tile((int)(it & 1u), b, h, 0);
      mbar_wait(kv_bar, strip_it & 1u);
    }
    for (int mt = 0; mt < p.num_m; ++mt, ++it) {
      const int slot = (int)(it & 1u);
      const uint32_t ph = it & 1u;                 // s / c / q / d barriers complete once per tile
      const uint32_t fph = (it >> 1) & 1u;         // each full barrier completes once per two tiles
      const int m0 = mt * BM;
      const uint32_t sdO_s = sT_s + (uint32_t)(slot * T_BYTES + T_DO);
      const uint32_t sQ_s = sT_s + (uint32_t)(slot * T_BYTES + T_Q);
      const uint32_t sP_s = sT_s + (uint32_t)(slot * T_BYTES + T_P);
      const uint32_t acc = mt > 0 ? 1u : 0u;       // dK / dV accumulate across the tiles of the strip

      if (tid == 0) {
        if (mt + 1 < p.num_m) {
          // the other slot was released at the end of the previous tile (d_bar waited, dQ store read)
          load_tile(slot ^ 1, b, h, m0 + BM);
        }
        mbar_wait(&full_bar[slot], fph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)     // MMA-A: dP = dO V^T
          umma_bf16(tmem_base, smem_desc(sdO_s + k * 32, 0u, 1024u), smem_desc(sV_s + k * 32, 0u, 1024u), idesc_dp, k > 0 ? 1u : 0u);
        umma_commit(s_bar);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)     // MMA-C: dV[blk] += P^T dO over the 128 queries of the tile (8 k-steps of 16)
#pragma unroll
          for (int kq = 0; kq < BM / 16; ++kq)
            umma_bf16(tmem_base + (uint32_t)(COL_DV + 64 * blk),
                      smem_desc(sP_s + (uint32_t)(blk * TILE) + (uint32_t)kq * 2048u, (uint32_t)TILE, 1024u),
                      smem_desc(sdO_s + (uint32_t)kq * 2048u, 8192u, 1024u), idesc_kv, (acc | (kq > 0 ? 1u : 0u)));
        umma_commit(c_bar);
      }
      __syncwarp();
      mbar_wait(s_bar, ph);
      tc_fence_after();

      // ---- softmax backward for row `row`: pass 1 = delta (each column half its partial, exchanged through shared memory),
      // pass 2 = dS in place over P
      float delta = 0.f;
      for (int i = c_lo; i < c_hi; ++i) {
        uint32_t r[16];
        tmem_ld_32x16(taddr + (uint32_t)(i * 16), r);
        const uint32_t base = sP_s + (uint32_t)(i >> 2) * (uint32_t)TILE + row_s;
        const uint32_t q0 = (uint32_t)(i & 3) * 2u;
        const uint4 pa = ld_shared_v4(base + (((q0) ^ row_x) << 4)), pb = ld_shared_v4(base + (((q0 + 1u) ^ row_x) << 4));
        const uint32_t pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = unpack_bf16x2(pp[j]);
          delta = fmaf(f.x, __uint_as_float(r[2 * j]), delta);
          delta = fmaf(f.y, __uint_as_float(r[2 * j + 1]), delta);
        }
      }
      xch[half * BM + row] = delta;
      __syncthreads();
      delta += xch[(half ^ 1) * BM + row];
      mbar_wait(c_bar, ph);      // the tensor core has finished reading P (dV): it may be overwritten
      for (int i = c_lo; i < c_hi; ++i) {
        uint32_t r[16];
        tmem_ld_32x16(taddr + (uint32_t)(i * 16), r);
        const uint32_t base = sP_s + (uint32_t)(i >> 2) * (uint32_t)TILE + row_s;
        const uint32_t q0 = (uint32_t)(i & 3) * 2u;
        const uint32_t a0 = base + (((q0) ^ row_x) << 4), a1 = base + (((q0 + 1u) ^ row_x) << 4);
        const uint4 pa = ld_shared_v4(a0), pb = ld_shared_v4(a1);
        const uint32_t pp[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 f = unpack_bf16x2(pp[j]);
          pk[j] = pack_bf16x2(scale * f.x * (__uint_as_float(r[2 * j]) - delta), scale * f.y * (__uint_as_float(r[2 * j + 1]) - delta));
        }
        st_shared_v4(a0, pk[0], pk[1], pk[2], pk[3]);
        st_shared_v4(a1, pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();     // dS (generic-proxy stores) -> visible to the tensor core
      tc_fence_before();       // this thread's TMEM reads of dP precede the MMA that overwrites those columns with dQ
      __syncthreads();

      if (tid == 0) {
        tc_fence_after();
        for (int kk = 0; kk < nchunk16; ++kk)   // MMA-B: dQ = dS K (A K-major atoms of 64 keys, B = K[key, d] MN-major)
          umma_bf16(tmem_base, smem_desc(sP_s + (uint32_t)(kk >> 2) * (uint32_t)TILE + (uint32_t)(kk & 3) * 32u, 0u, 1024u),
                    smem_desc(sK_s + (uint32_t)kk * 2048u, 8192u, 1024u), idesc_dq, kk > 0 ? 1u : 0u);
        umma_commit(q_bar);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)       // MMA-D: dK[blk] += dS^T Q
#pragma unroll
          for (int kq = 0; kq < BM / 16; ++kq)
            umma_bf16(tmem_base + (uint32_t)(COL_DK + 64 * blk),
                      smem_desc(sP_s + (uint32_t)(blk * TILE) + (uint32_t)kq * 2048u, (uint32_t)TILE, 1024u),
                      smem_desc(sQ_s + (uint32_t)kq * 2048u, 8192u, 1024u), idesc_kv, (acc | (kq > 0 ? 1u : 0u)));
        umma_commit(d_bar);
      }
      __syncwarp();
      mbar_wait(q_bar, ph);
      tc_fence_after();

      // ---- dQ epilogue: 64 fp32 columns -> bf16 -> swizzled tile in the dO buffer (MMA-A / MMA-C have retired)
#pragma unroll
      for (int ii = 0; ii < HD / 32; ++ii) {
        const int i = half * (HD / 32) + ii;
        uint32_t r[16];
        tmem_ld_32x16(taddr + (uint32_t)(i * 16), r);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        const uint32_t base = sdO_s + row_s;
        st_shared_v4(base + (((uint32_t)(2 * i) ^ row_x) << 4), pk[0], pk[1], pk[2], pk[3]);
        st_shared_v4(base + (((uint32_t)(2 * i + 1) ^ row_x) << 4), pk[4], pk[5], pk[6], pk[7]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tma_store_4d(&tmdQ, sdO_s, h * HD, m0, b, 0);   // rows past N are clipped by the tensor map
        tma_store_commit();
      }
      __syncwarp();
      mbar_wait(d_bar, ph);          // MMA-D retired: Q and the dS tile of this slot are free, dK is up to date
      tc_fence_after();
      if (tid == 0) tma_store_wait_read();   // ... and so is the slot's dO buffer (dQ staging)
      __syncwarp();
    }

    // ---- strip epilogue: dK / dV (two 128-key blocks each: keys 0..127 and keys 64..191) -> bf16 -> four staging tiles
    // (the P atoms of slot 0 and the first of slot 1; every MMA has retired: all threads waited c_bar and d_bar of the
    // last tile) -> TMA stores into dKV[b, key, (K | V) head h]
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const uint32_t col = (uint32_t)((t < 2 ? COL_DK : COL_DV) + 64 * (t & 1));
      const uint32_t stage = sT_s + (uint32_t)(t < 3 ? (T_P + t * TILE) : (T_BYTES + T_P));
#pragma unroll
      for (int ii = 0; ii < HD / 32; ++ii) {
        const int i = half * (HD / 32) + ii;
        uint32_t r[16];
        tmem_ld_32x16(taddr + col + (uint32_t)(i * 16), r);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        const uint32_t base = stage + row_s;
        st_shared_v4(base + (((uint32_t)(2 * i) ^ row_x) << 4), pk[0], pk[1], pk[2], pk[3]);
        st_shared_v4(base + (((uint32_t)(2 * i + 1) ^ row_x) << 4), pk[4], pk[5], pk[6], pk[7]);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t stage = sT_s + (uint32_t)(t < 3 ? (T_P + t * TILE) : (T_BYTES + T_P));
        // block 0 = key rows 0.., block 1 = key rows 64..; rows past Nk are clipped by the tensor map
        tma_store_4d(&tmdKV, stage, (t < 2 ? 0 : p.C) + h * HD, 64 * (t & 1), b, 0);
      }
      tma_store_commit();
      tma_store_wait_read();   // the staging tiles are tile slots of the next strip
    }
    __syncwarp();
  }

  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// dq[B*N, C], dkv[B*Nk, 2C] (dK | dV column halves) from q [B*N, C], kv [B*Nk, 2C], do [B*N, C] and the saved
// probabilities p [B, heads, N, Nk] (all bf16, contiguous, 16-byte aligned). Nk % 32 == 0, Nk <= 192, head dim 64.
extern "C" int mvlt_sr_attention_bwd(const void* q_bf16, const void* kv_bf16, const void* do_bf16, const void* p_bf16,
                                     void* dq_bf16, void* dkv_bf16, int B, int N, int Nk, int heads, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MVLT_CHECK_ARG(q_bf16 && kv_bf16 && do_bf16 && p_bf16 && dq_bf16 && dkv_bf16, "sr_attention_bwd: null operand");
  MVLT_CHECK_ARG(B > 0 && N > 0 && heads > 0, "sr_attention_bwd: bad shape B=%d N=%d heads=%d", B, N, heads);
  MVLT_CHECK_ARG(Nk >= 32 && Nk <= NK_MAX && Nk % 32 == 0, "sr_attention_bwd: Nk=%d unsupported (multiple of 32, <= %d)", Nk, NK_MAX);
  MVLT_CHECK_ARG(((((uintptr_t)q_bf16) | ((uintptr_t)kv_bf16) | ((uintptr_t)do_bf16) | ((uintptr_t)p_bf16) | ((uintptr_t)dq_bf16) |
                   ((uintptr_t)dkv_bf16)) & 15) == 0, "sr_attention_bwd: operands must be 16-byte aligned");
  const int C = heads * HD;
  BwdParams p;
  p.B = B; p.heads = heads; p.N = N; p.Nk = Nk; p.C = C;
  p.num_m = (N + BM - 1) / BM;
  p.strips = B * heads;
  p.scale = scale;

  CUtensorMap tmQ, tmdO, tmKV, tmP, tmdQ, tmdKV;
  int rc;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)N, (uint64_t)B, 1};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)N * C * 2, (uint64_t)B * N * C * 2};
    const uint32_t box[4] = {HD, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmQ, q_bf16, dims, str, box, 0, 0)) != 0) return rc;
    if ((rc = mvlt_tensor_map_4d(&tmdO, do_bf16, dims, str, box, 0, 0)) != 0) return rc;
    if ((rc = mvlt_tensor_map_4d(&tmdQ, dq_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)2 * C, (uint64_t)Nk, (uint64_t)B, 1};
    const uint64_t str[3] = {(uint64_t)C * 4, (uint64_t)Nk * C * 4, (uint64_t)B * Nk * C * 4};
    const uint32_t box_ld[4] = {HD, (uint32_t)Nk, 1, 1};
    const uint32_t box_st[4] = {HD, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmKV, kv_bf16, dims, str, box_ld, 0, 0)) != 0) return rc;
    if ((rc = mvlt_tensor_map_4d(&tmdKV, dkv_bf16, dims, str, box_st, 0, 0)) != 0) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)Nk, (uint64_t)N, (uint64_t)heads, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)Nk * 2, (uint64_t)N * Nk * 2, (uint64_t)heads * N * Nk * 2};
    const uint32_t box[4] = {64, BM, 1, 1};
    if ((rc = mvlt_tensor_map_4d(&tmP, p_bf16, dims, str, box, 0, 0)) != 0) return rc;
  }
  static bool attr_set = false;   // idempotent
  if (!attr_set) {
    cudaFuncSetAttribute(sr_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_USED + 1024);
    attr_set = true;
  }
  int grid = mvlt_num_sms();
  if (p.strips < grid) grid = p.strips;
  mvlt_launch(sr_attention_bwd_kernel, grid, THREADS, (size_t)SMEM_USED + 1024, stream, tmQ, tmdO, tmKV, tmP, tmdQ, tmdKV, p);
  MVLT_CHECK_LAUNCH();
  return 0;
}
