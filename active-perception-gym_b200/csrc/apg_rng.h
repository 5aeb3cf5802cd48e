// numpy-compatible random streams for the reset path, usable on host and device.
//
// Restates (integer-only, bit-exact) the published algorithms of numpy >= 1.17:
//   SeedSequence(entropy).generate_state(4, uint64)   numpy/random/bit_generator.pyx (hashmix / mix, pool of 4)
//   PCG64 (XSL-RR 128/64, setseq)                     numpy/random/src/pcg64/pcg64.h
//   Generator.integers: Lemire bounded draws          numpy/random/src/distributions/distributions.c
//       range <  2^32-1 : buffered 32-bit Lemire with rejection (low half of a raw 64-bit word first)
//       range == 2^32-1 : next buffered 32-bit word
//       range >  2^32-1 : 64-bit Lemire on a fresh raw word
// The reference draws these per environment at every (auto)reset (lidar_localization2d.py:303-311,547-557;
// dataset_iterator.py:30); keeping one Generator per environment on the host is untenable at 10^5..10^6
// environments, hence the device implementation.  tests/test_rng_host.py pins it against numpy itself.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define APG_HD __host__ __device__ __forceinline__
#else
#define APG_HD inline
#endif

namespace apg {

typedef unsigned __int128 u128;

struct Pcg64 {
  uint64_t state_hi, state_lo;
  uint64_t inc_hi, inc_lo;
  uint32_t has_uint32;
  uint32_t uinteger;
};

// ---- SeedSequence -------------------------------------------------------------------------------------
APG_HD uint32_t ss_hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= 0x931e8875u;
  value *= hash_const;
  value ^= value >> 16;
  return value;
}

APG_HD uint32_t ss_mix(uint32_t x, uint32_t y) {
  uint32_t r = 0xca01f9ddu * x - 0x4973f715u * y;
  r ^= r >> 16;
  return r;
}

// entropy: a non-negative integer < 2^64 (little-endian uint32 words, at least one word); out: 4 uint64 words
APG_HD void seedseq_state4(uint64_t entropy, uint64_t out[4]) {
  uint32_t ent[2];
  int n_ent = 1;
  ent[0] = (uint32_t)entropy;
  ent[1] = (uint32_t)(entropy >> 32);
  if (ent[1] != 0) n_ent = 2;
  uint32_t pool[4];
  uint32_t hc = 0x43b0d7e5u;
  for (int i = 0; i < 4; ++i) pool[i] = ss_hashmix(i < n_ent ? ent[i] : 0u, hc);
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], hc));
  // (entropy longer than the pool would be folded in here; n_ent <= 2 < 4)
  uint32_t hb = 0x8b51f9ddu;
  uint32_t w[8];
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i & 3];
    v ^= hb;
    hb *= 0x58f38dedu;
    v *= hb;
    v ^= v >> 16;
    w[i] = v;
  }
  for (int i = 0; i < 4; ++i) out[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
}

// ---- PCG64 ----------------------------------------------------------------------------------------------
APG_HD void pcg64_step(Pcg64& g) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | (u128)0x4385DF649FCCF645ull;
  u128 s = ((u128)g.state_hi << 64) | g.state_lo;
  const u128 inc = ((u128)g.inc_hi << 64) | g.inc_lo;
  s = s * mult + inc;
  g.state_hi = (uint64_t)(s >> 64);
  g.state_lo = (uint64_t)s;
}

APG_HD void pcg64_seed(Pcg64& g, uint64_t entropy) {
  uint64_t w[4];
  seedseq_state4(entropy, w);
  // PCG64.__init__: initstate = w[0] << 64 | w[1], initseq = w[2] << 64 | w[3]; pcg64_srandom_r
  const u128 initstate = ((u128)w[0] << 64) | w[1];
  const u128 initseq = ((u128)w[2] << 64) | w[3];
  const u128 inc = (initseq << 1) | 1;
  g.inc_hi = (uint64_t)(inc >> 64);
  g.inc_lo = (uint64_t)inc;
  g.state_hi = 0;
  g.state_lo = 0;
  pcg64_step(g);
  u128 s = (((u128)g.state_hi << 64) | g.state_lo) + initstate;
  g.state_hi = (uint64_t)(s >> 64);
  g.state_lo = (uint64_t)s;
  pcg64_step(g);
  g.has_uint32 = 0;
  g.uinteger = 0;
}

APG_HD uint64_t pcg64_next64(Pcg64& g) {
  pcg64_step(g);
  const uint64_t x = g.state_hi ^ g.state_lo;
  const unsigned rot = (unsigned)(g.state_hi >> 58);
  return (x >> rot) | (x << ((64u - rot) & 63u));
}

// jump the 128-bit LCG ahead by `delta` raw draws (the 32-bit half-word buffer is left untouched)
APG_HD void pcg64_advance(Pcg64& g, uint64_t delta) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | (u128)0x4385DF649FCCF645ull;
  u128 cur_mult = mult, cur_plus = ((u128)g.inc_hi << 64) | g.inc_lo;
  u128 acc_mult = 1, acc_plus = 0;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  u128 s = ((u128)g.state_hi << 64) | g.state_lo;
  s = acc_mult * s + acc_plus;
  g.state_hi = (uint64_t)(s >> 64);
  g.state_lo = (uint64_t)s;
}

// Generator.uniform(low, low + range): low + range * ((raw >> 11) * 2^-53)   (one raw draw per double)
APG_HD double rng_uniform(Pcg64& g, double low, double range) {
  const double u = (double)(pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0);
  return low + range * u;
}

APG_HD uint32_t pcg64_next32(Pcg64& g) {
  if (g.has_uint32) {
    g.has_uint32 = 0;
    return g.uinteger;
  }
  const uint64_t n = pcg64_next64(g);
  g.has_uint32 = 1;
  g.uinteger = (uint32_t)(n >> 32);
  return (uint32_t)n;
}

// Generator.integers(0, high_exclusive) for 1 <= high_exclusive <= 2^32  (int64 result)
APG_HD uint64_t rng_integers_below(Pcg64& g, uint64_t high_exclusive) {
  const uint64_t rng = high_exclusive - 1;
  if (rng == 0) return 0;
  if (rng == 0xFFFFFFFFull) return pcg64_next32(g);
  const uint32_t rng_excl = (uint32_t)rng + 1u;
  uint64_t m = (uint64_t)pcg64_next32(g) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xFFFFFFFFu - (uint32_t)rng) % rng_excl;
    while (leftover < threshold) {
      m = (uint64_t)pcg64_next32(g) * rng_excl;
      leftover = (uint32_t)m;
    }
  }
  return m >> 32;
}

// Generator.integers(0, 2**32, endpoint=True): range 2^32 > 0xFFFFFFFF -> 64-bit Lemire
APG_HD uint64_t rng_integers_0_to_2p32_inclusive(Pcg64& g) {
  const uint64_t rng = 0x100000000ull;
  const uint64_t rng_excl = rng + 1;
  u128 m = (u128)pcg64_next64(g) * rng_excl;
  uint64_t leftover = (uint64_t)m;
  if (leftover < rng_excl) {
    const uint64_t threshold = (0xFFFFFFFFFFFFFFFFull - rng) % rng_excl;
    while (leftover < threshold) {
      m = (u128)pcg64_next64(g) * rng_excl;
      leftover = (uint64_t)m;
    }
  }
  return (uint64_t)(m >> 64);
}

}  // namespace apg
