// Random grid masking on the device, bit-exact against the reference's host implementation for a given
// per-sample MT19937 seed: /root/reference/mcloader/fashion_gen.py:225-254 (generate_grid_mask, with its
// sliding-window quirk at :246) and :176-177 (masked_fill_ with 1e-6). Integer kernel: numpy's legacy
// np.random.seed(int) + list shuffle (= Fisher-Yates from the top with masked-rejection random_interval)
// are restated here; one warp per sample keeps the 624-word generator state in shared memory.
//
// BERT token masking (15 % of the word pieces: 80 % -> [MASK], 10 % -> random vocabulary entry, 10 % kept; labels = the
// original ids, -1 elsewhere) is the same kind of integer kernel: /root/reference/mcloader/fashion_gen.py:383-409
// (random_masking_features) drives CPython's `random` module, i.e. MT19937 seeded through init_by_array, random() as a
// 53-bit double from two draws and choice() as getrandbits rejection sampling; restated bit for bit per sample seed.
#include "common.cuh"

namespace {

struct MT {
  uint32_t* mt;
  int idx;
};
__device__ void mt_seed(MT& s, uint32_t seed) {
  s.mt[0] = seed;
  for (int i = 1; i < 624; ++i) s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
  s.idx = 624;
}
__device__ uint32_t mt_next(MT& s) {
  if (s.idx >= 624) {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[(k + 1) % 624] & 0x7fffffffu);
      s.mt[k] = s.mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s.idx = 0;
  }
  uint32_t y = s.mt[s.idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__device__ uint32_t mt_interval(MT& s, uint32_t mx) {
  if (mx == 0) return 0;
  uint32_t mask = mx;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (mt_next(s) & mask)) > mx) {}
  return v;
}
__device__ void mt_shuffle(MT& s, uint8_t* v, int n) {
  for (int i = n - 1; i >= 1; --i) {
    const uint32_t j = mt_interval(s, (uint32_t)i);
    const uint8_t t = v[i]; v[i] = v[j]; v[j] = t;
  }
}

// CPython random.seed(int) for 0 <= seed < 2^32: init_by_array(key = [seed])  (Modules/_randommodule.c)
__device__ void mt_seed_py(MT& s, uint32_t seed) {
  mt_seed(s, 19650218u);
  uint32_t* mt = s.mt;
  int i = 1;
  for (int k = 624; k; --k) {           // key_length = 1: init_key[j] + j == seed + 0 every round
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + seed;
    if (++i >= 624) { mt[0] = mt[623]; i = 1; }
  }
  for (int k = 623; k; --k) {
    mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    if (++i >= 624) { mt[0] = mt[623]; i = 1; }
  }
  mt[0] = 0x80000000u;
  s.idx = 624;
}
// random.random(): 53-bit double from two 32-bit draws
__device__ double mt_random_py(MT& s) {
  const uint32_t a = mt_next(s) >> 5, b = mt_next(s) >> 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}
// random.choice(seq) index: _randbelow(n) = getrandbits(n.bit_length()) with rejection
__device__ uint32_t mt_randbelow_py(MT& s, uint32_t n) {
  const int k = 32 - __clz(n);
  uint32_t r = mt_next(s) >> (32 - k);
  while (r >= n) r = mt_next(s) >> (32 - k);
  return r;
}

// one warp per sample; lane 0 walks the generator over the word pieces between [CLS] and [SEP]
__global__ void __launch_bounds__(32) token_mask_kernel(const uint32_t* __restrict__ seeds, const long long* __restrict__ ori,
                                                        long long* __restrict__ ids, long long* __restrict__ labels, int B, int T,
                                                        int sep_id, int mask_id, int vocab, double rate) {
  pdl_prologue();
  __shared__ uint32_t state[624];
  const int b = blockIdx.x;
  const long long* o = ori + (long long)b * T;
  long long* x = ids + (long long)b * T;
  long long* l = labels + (long long)b * T;
  for (int i = threadIdx.x; i < T; i += 32) {
    x[i] = o[i];
    l[i] = -1;
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    MT s{state, 624};
    mt_seed_py(s, seeds[b]);
    for (int i = 1; i < T; ++i) {
      const long long tok = o[i];
      if (tok == sep_id) break;                       // word pieces are positions 1 .. (first [SEP]) - 1
      double prob = mt_random_py(s);
      if (prob < rate) {
        prob /= rate;
        if (prob < 0.8) x[i] = mask_id;
        else if (prob < 0.9) x[i] = (long long)mt_randbelow_py(s, (uint32_t)vocab);   // vocab.items() is in id order
        l[i] = tok;
      }
    }
  }
}

constexpr int MAX_PATCHES = 4096;
constexpr int MAX_NW = 64;

// one warp per sample; lane 0 walks the (inherently sequential) generator
__global__ void __launch_bounds__(32) grid_mask_kernel(const uint32_t* __restrict__ seeds, uint8_t* __restrict__ grid, int B,
                                                       int nw, int nh, int n_mask) {
  pdl_prologue();
  __shared__ uint32_t state[624];
  __shared__ uint8_t vals[MAX_PATCHES];
  __shared__ uint8_t row[MAX_NW];
  const int b = blockIdx.x;
  const int n = nw * nh;
  for (int i = threadIdx.x; i < n; i += 32) vals[i] = (i >= n - n_mask) ? 1 : 0;
  __syncwarp();
  if (threadIdx.x == 0) {
    MT s{state, 624};
    mt_seed(s, seeds[b]);
    mt_shuffle(s, vals, n);
    for (int r = 0; r < nh; ++r) {
      int len = nw;
      if (r + nw > n) len = n - r;
      for (int c = 0; c < len; ++c) row[c] = vals[r + c];  // sliding window vals[r : r+nw] (reference quirk)
      mt_shuffle(s, row, len);
      for (int c = 0; c < len; ++c) grid[(long long)b * n + r * nw + c] = row[c];
    }
  }
}

// masked[b,c,y,x] = grid[b][y/P][x/P] ? fill : img[b,c,y,x]; optionally mask_out[b,0,y,x] = grid as float
__global__ void __launch_bounds__(256) masked_fill_kernel(const float* __restrict__ img, const uint8_t* __restrict__ grid,
                                                          float* __restrict__ out, float* __restrict__ mask_out, int B,
                                                          int Cc, int H, int W, int P, float fill) {
  pdl_prologue();
  const int w4 = W / 4;
  const long long total = (long long)B * Cc * H * w4;
  const int nw = W / P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)((unsigned int)i % (unsigned int)w4) * 4;
    unsigned int t = (unsigned int)i / (unsigned int)w4;
    const int y = (int)(t % H); t /= H;
    const int c = (int)(t % Cc);
    const int b = (int)(t / Cc);
    const uint8_t g = grid[(long long)b * (H / P) * nw + (y / P) * nw + x / P];
    const long long off = (((long long)b * Cc + c) * H + y) * W + x;
    float4 v = *reinterpret_cast<const float4*>(img + off);
    if (g) v = make_float4(fill, fill, fill, fill);
    *reinterpret_cast<float4*>(out + off) = v;
    if (mask_out && c == 0) {
      const float m = g ? 1.f : 0.f;
      *reinterpret_cast<float4*>(mask_out + ((long long)b * H + y) * W + x) = make_float4(m, m, m, m);
    }
  }
}

}  // namespace

extern "C" int mvlt_grid_mask(const uint32_t* seeds_dev, uint8_t* grid_out, int B, int size_w, int size_h, int patch,
                              double mask_ratio, void* stream_) {
  MVLT_CHECK_ARG(patch > 0 && size_w % patch == 0 && size_h % patch == 0, "grid_mask: size must be divisible by patch");
  const int nw = size_w / patch, nh = size_h / patch;
  MVLT_CHECK_ARG(nw * nh <= MAX_PATCHES && nw <= MAX_NW, "grid_mask: too many patches");
  const int n_mask = (int)(mask_ratio * (double)(nw * nh));
  mvlt_launch(grid_mask_kernel, B, 32, 0, reinterpret_cast<cudaStream_t>(stream_), seeds_dev, grid_out, B, nw, nh, n_mask);
  MVLT_CHECK_LAUNCH();
  return 0;
}

// ori_ids / input_ids / labels: int64 [B, T]; row = [CLS] pieces... [SEP] [PAD]...; seeds: one CPython random.seed(int) per sample
extern "C" int mvlt_token_mask(const uint32_t* seeds_dev, const long long* ori_ids, long long* input_ids, long long* labels, int B,
                               int T, int sep_id, int mask_id, int vocab_size, double mask_rate, void* stream_) {
  MVLT_CHECK_ARG(B > 0 && T > 1 && vocab_size > 0 && mask_rate > 0.0 && mask_rate <= 1.0, "token_mask: bad arguments");
  mvlt_launch(token_mask_kernel, B, 32, 0, reinterpret_cast<cudaStream_t>(stream_), seeds_dev, ori_ids, input_ids, labels, B, T, sep_id,
                                                                           mask_id, vocab_size, mask_rate);
  MVLT_CHECK_LAUNCH();
  return 0;
}

extern "C" int mvlt_masked_fill(const float* img, const uint8_t* grid, float* out, float* mask_out, int B, int C, int H,
                                int W, int patch, float fill, void* stream_) {
  MVLT_CHECK_ARG(W % 4 == 0 && patch % 4 == 0 && H % patch == 0 && W % patch == 0, "masked_fill: bad geometry");
  const long long total = (long long)B * C * H * (W / 4);
  MVLT_CHECK_ARG(total < (1ll << 32), "masked_fill: tensor too large for the 32-bit index decomposition");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)mvlt_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  mvlt_launch(masked_fill_kernel, (int)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_), img, grid, out, mask_out, B, C, H, W,
                                                                                     patch, fill);
  MVLT_CHECK_LAUNCH();
  return 0;
}
