// Procedural rooms maps, bit-identical to the reference's generator
//   FloorMapDatasetRooms.get_data_point      ap_gym/envs/floor_map/floor_map_dataset_rooms.py:25-89
// as host/device code: the recursive `split_room` becomes an explicit stack of room views, and the numpy draws of
// `default_rng(idx)` are restated at the level of the bit generator (csrc/apg_rng.h), in the reference's order:
//   rng.binomial(n, 0.3)                     numpy random_binomial -> inversion (n * p <= 30): one next_double per attempt;
//                                            n == 0 draws nothing.  exp(n * log(q)) comes from a host-computed table
//                                            (libm's exp / log are not bit-reproducible on the device).
//   rng.choice(r, k - 1, replace=False)      Floyd's algorithm (pop_size <= 10000): bounded_lemire_uint32 draws with
//                                            range j (none for j == 0), then _shuffle_int: one bounded draw per i = k-2..1;
//                                            the result is sorted by the caller (np.sort), so only the draws matter
//   rng.integers(0, h, size)                 buffered 32-bit Lemire draws (range 0 draws nothing)
// The CPU test (tests/test_maps.py) compiles this header with g++ and compares with the reference-generated bank
// tests/golden/maps_rooms32.npz; the kernel in csrc/apg_maps.cu runs the same code per thread.
#pragma once

#include <math.h>
#include <stdint.h>

#include "apg_rng.h"

namespace apg_rooms {

constexpr int kMaxRooms = 16;   // max_rooms (sub-rooms of one split) supported
constexpr int kMaxStack = 64;   // pending room views (depth-first: <= (sub-rooms - 1) per level, a few levels deep)
constexpr double kP = 0.3;

// view(a, b) -> map(y0 + (tr ? b : a), x0 + (tr ? a : b)); shape (A, B); cap = max_rooms of the pending split
struct View {
  short y0, x0, A, B, cap;
  unsigned char tr;
};

APG_HD int floordiv(int a, int b) {  // python's a // b for b > 0
  int q = a / b;
  if ((a % b != 0) && (a < 0)) q -= 1;
  return q;
}

APG_HD uint64_t bounded(apg::Pcg64& g, uint64_t rng) {  // random_bounded_uint64(bitgen, 0, rng, 0, False), rng < 2^32 - 1
  return apg::rng_integers_below(g, rng + 1);
}

APG_HD double next_double(apg::Pcg64& g) { return (double)(apg::pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0); }

// numpy's random_binomial(p = 0.3, n) for n * p <= 30; qn = exp(n * log(1 - p)) from the host
APG_HD int binomial03(apg::Pcg64& g, int n, const double* qn_table) {
  if (n == 0) return 0;
  const double p = kP, q = 1.0 - p;
  const double qn = qn_table[n];
  const double np = (double)n * p;
  const double bd = np + 10.0 * sqrt(np * q + 1.0);
  const long long bound = (long long)((double)n < bd ? (double)n : bd);
  long long X = 0;
  double px = qn;
  double U = next_double(g);
  while (U > px) {
    X++;
    if (X > bound) {
      X = 0;
      px = qn;
      U = next_double(g);
    } else {
      U -= px;
      px = (((double)(n - X + 1) * p) * px) / ((double)X * q);
    }
  }
  return (int)X;
}

// distribute_integers(n, k, rng) of the reference: k non-negative parts of n from k - 1 sorted cuts
APG_HD void distribute_integers(apg::Pcg64& g, int n, int k, int* out) {
  const int zeros = k > n ? k - n : 0;
  const int pop = zeros + (n > 1 ? n - 1 : 0);
  const int size = k - 1;
  long long idx[kMaxRooms];
  for (int t = 0; t < size; ++t) {  // Floyd: j = pop - size + t
    const long long j = (long long)(pop - size + t);
    const long long val = (long long)bounded(g, (uint64_t)j);
    bool present = false;
    for (int u = 0; u < t; ++u) present = present || (idx[u] == val);
    idx[t] = present ? j : val;
  }
  for (int i = size - 1; i >= 1; --i) {  // _shuffle_int(bitgen, size, 1, idx)
    const long long j = (long long)bounded(g, (uint64_t)i);
    const long long tmp = idx[j]; idx[j] = idx[i]; idx[i] = tmp;
  }
  int cuts[kMaxRooms];
  for (int t = 0; t < size; ++t) cuts[t] = idx[t] < zeros ? 0 : (int)(idx[t] - zeros + 1);  // r[idx]
  for (int a = 1; a < size; ++a) {  // np.sort
    const int v = cuts[a];
    int b = a - 1;
    while (b >= 0 && cuts[b] > v) { cuts[b + 1] = cuts[b]; --b; }
    cuts[b + 1] = v;
  }
  int prev = 0;
  for (int t = 0; t < size; ++t) { out[t] = cuts[t] - prev; prev = cuts[t]; }
  out[size] = n - prev;
}

APG_HD void set_cell(unsigned* plane, int words, const View& v, int a, int b) {
  const int y = v.y0 + (v.tr ? b : a), x = v.x0 + (v.tr ? a : b);
  plane[y * words + (x >> 5)] |= 1u << (x & 31);
}

// One map.  wall / door: (H, words) bit planes (scratch; door cells are the reference's -1 marks), stack: kMaxStack views.
// Returns 0, or 1 when a limit of this restatement is exceeded (the caller reports it).  The result (wall & ~door,
// transposed when the final draw says so) is written to `out`.  Square maps only (the reference's transpose of a
// non-square map would change its shape).
APG_HD int generate(uint64_t index, int W, int H, int max_rooms, int door_width, const double* qn_table, unsigned* wall,
                    unsigned* door, View* stack, unsigned* out) {
  const int words = (W + 31) >> 5;
  const int mrs = door_width + 2;
  apg::Pcg64 g;
  apg::pcg64_seed(g, index);
  for (int i = 0; i < H * words; ++i) { wall[i] = 0u; door[i] = 0u; }
  View full{0, 0, (short)H, (short)W, 0, 0};
  for (int x = 0; x < W; ++x) { set_cell(wall, words, full, 0, x); set_cell(wall, words, full, H - 1, x); }
  for (int y = 0; y < H; ++y) { set_cell(wall, words, full, y, 0); set_cell(wall, words, full, y, W - 1); }
  int sp = 0;
  stack[sp++] = View{1, 1, (short)(H - 2), (short)(W - 2), (short)max_rooms, 0};
  while (sp > 0) {
    const View v = stack[--sp];
    const int A = v.A, B = v.B;
    int mrl = floordiv(A - mrs, mrs + 1) + 1;
    if (v.cap < mrl) mrl = v.cap;
    if (mrl <= 1) continue;
    if (mrl > kMaxRooms) return 1;
    const int sub = binomial03(g, mrl - 2, qn_table) + 2;
    int cap[kMaxRooms], sizes[kMaxRooms];
    distribute_integers(g, mrl, sub, cap);
    distribute_integers(g, A - sub * (1 + mrs) + 1, sub, sizes);
    int starts[kMaxRooms], ends[kMaxRooms];
    int acc = 0;
    for (int i = 0; i < sub; ++i) {
      acc += sizes[i] + mrs + 1;  // room_sizes_with_walls
      ends[i] = acc - 1;
      starts[i] = i == 0 ? 0 : ends[i - 1] + 2;
    }
    int doors[kMaxRooms];
    if (B - door_width < 1) return 1;  // numpy raises (low >= high)
    for (int i = 0; i + 1 < sub; ++i) doors[i] = (int)apg::rng_integers_below(g, (uint64_t)(B - door_width));
    for (int i = 0; i + 1 < sub; ++i) {  // wall rows (cells already marked -1 stay -1: wall & ~door at the end)
      const int w = starts[i + 1] - 1;
      for (int b = 0; b < B; ++b) set_cell(wall, words, v, w, b);
    }
    for (int i = 0; i + 1 < sub; ++i) {  // doors: rows w - (dw - 1) .. w + (dw - 1), columns d .. d + dw - 1
      const int w = starts[i + 1] - 1;
      for (int da = -(door_width - 1); da <= door_width - 1; ++da)
        for (int db = 0; db < door_width; ++db) {
          int a = w + da;
          if (a < 0) a += A;  // numpy's negative indices wrap
          if (a < 0 || a >= A || doors[i] + db >= B) return 1;  // numpy raises IndexError
          set_cell(door, words, v, a, doors[i] + db);
        }
    }
    if (sp + sub > kMaxStack) return 1;
    for (int i = sub - 1; i >= 0; --i) {  // depth-first in the reference's order: child 0 is popped first
      // room[s : e + 1].T : new(a', b') = old(s + b', a')
      View c;
      // (the sizes add up to A + 1 rows including the walls: the slice of the last room is clipped by numpy)
      const int e = ends[i] < A - 1 ? ends[i] : A - 1;
      c.A = (short)B; c.B = (short)(e - starts[i] + 1); c.cap = (short)cap[i]; c.tr = v.tr ? 0 : 1;
      c.y0 = (short)(v.y0 + (v.tr ? 0 : starts[i]));
      c.x0 = (short)(v.x0 + (v.tr ? starts[i] : 0));
      stack[sp++] = c;
    }
  }
  const bool transpose = apg::rng_integers_below(g, 2) == 0;
  for (int i = 0; i < H * words; ++i) out[i] = 0u;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const unsigned bit = ((wall[y * words + (x >> 5)] & ~door[y * words + (x >> 5)]) >> (x & 31)) & 1u;
      if (!bit) continue;
      const int oy = transpose ? x : y, ox = transpose ? y : x;
      out[oy * words + (ox >> 5)] |= 1u << (ox & 31);
    }
  return 0;
}

}  // namespace apg_rooms
