// Procedural maze maps, bit-identical to the reference's generator
//   FloorMapDatasetMaze.get_data_point      ap_gym/envs/floor_map/floor_map_dataset_maze.py:24-55
// as a device function (one thread generates one maze): see csrc/apg_maps.cu for the draw-by-draw description.  Used by
// the bank generator (apg_maps_generate_maze) and by the LIDAR engine's procedural-dataset mode (csrc/apg_lidar.cu).
#pragma once

#include "apg_rng.h"

namespace apg_maze {

__device__ __forceinline__ unsigned interval(apg::Pcg64& g, unsigned max, unsigned mask) {
  unsigned v;
  do { v = apg::pcg64_next32(g) & mask; } while (v > max);
  return v;
}

// direction order of one cell: numpy's Generator.permutation on a 4-row array = shuffle(arange(4)); 2 bits per entry
__device__ __forceinline__ unsigned draw_order(apg::Pcg64& g) {
  unsigned o[4] = {0u, 1u, 2u, 3u};
  {
    const unsigned j = interval(g, 3u, 3u);
    const unsigned t = o[3]; o[3] = o[j]; o[j] = t;
  }
  {
    const unsigned j = interval(g, 2u, 3u);
    const unsigned t = o[2]; o[2] = o[j]; o[j] = t;
  }
  {
    const unsigned j = interval(g, 1u, 1u);
    const unsigned t = o[1]; o[1] = o[j]; o[j] = t;
  }
  return o[0] | (o[1] << 2) | (o[2] << 4) | (o[3] << 6);
}

// frame: x (bits 0-7) | y (8-15) | order (16-23) | next direction slot k (24-26) | first (27)
// One maze: b = (H, words) output bitmap, st = explicit stack of ((W-1)/2) * ((H-1)/2) + 1 frames.
__device__ inline void generate(unsigned long long index, int W, int H, double branching_prob, unsigned* b, unsigned* st) {
  const int words = (W + 31) >> 5;
  for (int y = 0; y < H; ++y)
    for (int w = 0; w < words; ++w) {
      const int nb = min(32, W - w * 32);
      b[y * words + w] = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);  // all walls
    }
  apg::Pcg64 g;
  apg::pcg64_seed(g, index);
  b[1 * words] &= ~(1u << 1);  // starting cell (1, 1)
  int sp = 0;
  st[0] = 1u | (1u << 8) | (draw_order(g) << 16) | (1u << 27);
  const int dxs[4] = {2, -2, 0, 0}, dys[4] = {0, 0, 2, -2};
  while (sp >= 0) {
    unsigned fr = st[sp];
    const unsigned k = (fr >> 24) & 7u;
    if (k == 4u) {
      --sp;
      continue;
    }
    const int x = fr & 0xff, y = (fr >> 8) & 0xff;
    const unsigned d = (fr >> (16 + 2 * k)) & 3u;
    fr += 1u << 24;
    const int nx = x + dxs[d], ny = y + dys[d];
    bool push = false;
    if (nx > 0 && nx < W - 1 && ny > 0 && ny < H - 1 && ((b[ny * words + (nx >> 5)] >> (nx & 31)) & 1u)) {
      const bool first = (fr >> 27) & 1u;
      // `first or rng.random() < branching_prob`: the draw only happens when `first` is False
      if (first || ((double)(apg::pcg64_next64(g) >> 11) * (1.0 / 9007199254740992.0)) < branching_prob) {
        const int ix = x + dxs[d] / 2, iy = y + dys[d] / 2;
        b[iy * words + (ix >> 5)] &= ~(1u << (ix & 31));
        b[ny * words + (nx >> 5)] &= ~(1u << (nx & 31));
        fr &= ~(1u << 27);
        push = true;
      }
    }
    st[sp] = fr;
    if (push) {
      ++sp;
      st[sp] = (unsigned)nx | ((unsigned)ny << 8) | (draw_order(g) << 16) | (1u << 27);
    }
  }
}

}  // namespace apg_maze
