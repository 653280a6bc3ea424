/* ORACLE (test infrastructure): CPU restatement of the reference's random grid mask,
 * /root/reference/mcloader/fashion_gen.py:225-254, including numpy's legacy global RNG semantics that the
 * reference relies on (np.random.seed(int) -> MT19937 init_genrand; np.random.shuffle on a Python list ->
 * Fisher-Yates from the top with legacy random_interval = masked rejection sampling on 32-bit draws).
 * numpy (dependency, not vendored in the reference): numpy/random/_mt19937.pyx + src/legacy semantics.
 * Pinned by tests/golden/grid_mask_golden.npz, which was produced by calling the reference's own
 * generate_grid_mask under np.random.seed (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/liboracle_mask.so oracle/grid_mask.c
 */
#include <stdint.h>
#include <string.h>

typedef struct { uint32_t mt[624]; int idx; uint32_t draws; } mt_t;

static void mt_seed(mt_t* s, uint32_t seed) {
  s->mt[0] = seed;
  for (int i = 1; i < 624; ++i) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->idx = 624;
  s->draws = 0;
}

static uint32_t mt_next(mt_t* s) {
  if (s->idx >= 624) {
    for (int k = 0; k < 624; ++k) {
      uint32_t y = (s->mt[k] & 0x80000000u) | (s->mt[(k + 1) % 624] & 0x7fffffffu);
      s->mt[k] = s->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->idx = 0;
  }
  uint32_t y = s->mt[s->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  s->draws++;
  return y;
}

/* numpy legacy random_interval(max): uniform integer in [0, max] */
static uint32_t interval(mt_t* s, uint32_t max) {
  if (max == 0) return 0;
  uint32_t mask = max;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (mt_next(s) & mask)) > max) {}
  return v;
}

static void shuffle(mt_t* s, uint8_t* v, int n) {
  for (int i = n - 1; i >= 1; --i) {
    uint32_t j = interval(s, (uint32_t)i);
    uint8_t t = v[i]; v[i] = v[j]; v[j] = t;
  }
}

/* grid[r*nw + c] in {0,1}: 1 = masked patch. Returns the number of raw 32-bit draws consumed. */
uint32_t oracle_grid_mask(uint32_t seed, int size_w, int size_h, int patch, double ratio, uint8_t* grid) {
  const int nw = size_w / patch, nh = size_h / patch, n = nw * nh;
  const int nm = (int)(ratio * n);
  uint8_t vals[4096], row[64];
  mt_t s;
  if (n > 4096 || nw > 64) return 0;
  mt_seed(&s, seed);
  for (int i = 0; i < n; ++i) vals[i] = (i >= n - nm) ? 1 : 0;   /* zeros first, then ones (:236-239) */
  shuffle(&s, vals, n);                                           /* :242 */
  for (int r = 0; r < nh; ++r) {
    int len = nw;
    if (r + nw > n) len = n - r;
    memcpy(row, vals + r, (size_t)len);                           /* SLIDING window mask_split[i : i+nw] (:246) */
    shuffle(&s, row, len);                                        /* :247 */
    for (int c = 0; c < len; ++c) grid[r * nw + c] = row[c];
  }
  return s.draws;
}
