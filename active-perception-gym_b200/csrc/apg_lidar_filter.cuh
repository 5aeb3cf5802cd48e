// Conservative beam pre-filter of the LIDAR step: "may this ray touch a wall at all?" -- answered with a few integer
// instructions per cell row from a table of precomputed swath patterns.
//
// The beam directions are the same for every environment (lidar_localization2d.py:180-187), only the start point moves.
// For beam b and the sub-cell (qx, qy) of the start cell that holds the start point (Q x Q sub-cells per cell), the table
// holds the set of cells -- relative to the start cell -- that a segment start -> start + dir can touch for ANY start in
// that sub-cell: the cells meeting the swath  (sub-cell grown by 1/256) (+) [0, dir]  (Minkowski sum, a convex hexagon;
// closed cells).  The 1/256 margin is > 100x the float32 rounding of `target = pos + dir` and of every coordinate
// involved.  The exact walk (scan_one, apg_lidar.cu) only ever looks at cells whose closed square touches the segment
// (start cell, stepped-into cells, lattice-corner and grid-line neighbours), all of which lie inside the swath.  Hence
//     no wall bit under the patterns  =>  scan_one would find n_pos + n_zero == 0: distance = full length, no contact.
// Reference semantics: ap_gym/envs/lidar_localization2d.py:496-536 (an empty intersection falls into the `else` branch:
// distance = |target - pos|, no contact cell).  Checked against the C oracle on the CPU (tests/test_lidar_filter.py compiles
// this header as host code) and through the parity tests on the GPU.
//
// Table entry (16 bytes): kRows = 7 patterns of 16 bits, ascending rows; bit i of pattern k <-> cell
// (cx - 7 + i, cy + row0 + k); the 8th halfword holds row0 + 8.  Covers |dir| <= 5 (ceil(range) + 2 <= 7 rows).
#pragma once

#if defined(__CUDACC__)
#define APG_HD __host__ __device__ __forceinline__
#else
#define APG_HD inline
#endif

#include <math.h>
#include <stdint.h>

#if !defined(__CUDACC__)
struct uint4 { unsigned x, y, z, w; };  // host-only builds (tests/native): the CUDA vector type by its layout
#endif

namespace apg_lidar {

constexpr int kRows = 7;          // rows of a swath: |dy| + sub-cell + margins < 6 spans at most 7 cell rows
constexpr int kColBias = 7;       // pattern bit i <-> column cx - kColBias + i
constexpr float kMaxRange = 5.0f; // longer beams: no table, every ray is walked
constexpr double kMargin = 1.0 / 256.0;

struct BeamPattern {  // 16 bytes
  uint16_t rows[kRows];
  uint16_t row0_plus8;
};

// Host side: the pattern of beam direction (dx, dy) for sub-cell (qx, qy) of Q x Q.
inline BeamPattern make_beam_pattern(double dx, double dy, int qx, int qy, int Q) {
  const double e = kMargin;
  const double sx0 = (double)qx / Q - e, sx1 = (double)(qx + 1) / Q + e;
  const double sy0 = (double)qy / Q - e, sy1 = (double)(qy + 1) / Q + e;
  const double hx0 = sx0 + (dx < 0 ? dx : 0.0), hx1 = sx1 + (dx > 0 ? dx : 0.0);
  const double hy0 = sy0 + (dy < 0 ? dy : 0.0), hy1 = sy1 + (dy > 0 ? dy : 0.0);
  // projection of the grown sub-cell on the normal n = (-dy, dx) of the direction
  const double nx = -dy, ny = dx;
  double plo = 1e300, phi = -1e300;
  for (int c = 0; c < 4; ++c) {
    const double v = ((c & 1) ? sx1 : sx0) * nx + ((c & 2) ? sy1 : sy0) * ny;
    plo = v < plo ? v : plo; phi = v > phi ? v : phi;
  }
  const double slack = 1e-9 * (fabs(nx) + fabs(ny) + 1.0) * 16.0;  // rounding of the projections themselves
  BeamPattern out{};
  const int row0 = (int)floor(hy0);
  out.row0_plus8 = (uint16_t)(row0 + 8);
  if ((int)floor(hy1) - row0 >= kRows || floor(hx0) < -kColBias || floor(hx1) > 15 - kColBias || row0 < -8) {
    for (int k = 0; k < kRows; ++k) out.rows[k] = 0xffffu;  // does not fit the entry: never reject (callers check
    out.row0_plus8 = 8;                                     // beam_pattern_fits first, so this is belt and braces)
    return out;
  }
  for (int k = 0; k < kRows; ++k) {
    const int j = row0 + k;
    unsigned bits = 0;
    if ((double)j <= hy1 && (double)(j + 1) >= hy0) {
      for (int i = 0; i < 16; ++i) {
        const int ci = i - kColBias;
        if (!((double)ci <= hx1 && (double)(ci + 1) >= hx0)) continue;
        double clo = 1e300, chi = -1e300;
        for (int c = 0; c < 4; ++c) {
          const double v = (double)(ci + (c & 1)) * nx + (double)(j + ((c >> 1) & 1)) * ny;
          clo = v < clo ? v : clo; chi = v > chi ? v : chi;
        }
        if (clo <= phi + slack && chi >= plo - slack) bits |= 1u << i;
      }
    }
    out.rows[k] = (uint16_t)bits;
  }
  return out;
}

// true when the swath of the table covers every cell the direction can reach (7 rows, columns cx - 7 .. cx + 8)
inline bool beam_pattern_fits(double dx, double dy) {
  return fabs(dx) <= (double)kMaxRange + 1e-3 && fabs(dy) <= (double)kMaxRange + 1e-3;
}

APG_HD unsigned window32(unsigned lo, unsigned hi, int shift) {  // bits shift .. shift + 31 of hi:lo, 0 <= shift < 32
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, shift);
#else
  return shift ? ((lo >> shift) | (hi << (32 - shift))) : lo;
#endif
}

// `bits`: the padded bitmap of one map, (H + 2 pad) rows of pw words with one spare word behind the last row; bit x & 31
// of word x >> 5 of row y = wall at padded cell (x, y); pad >= 7.  `table`: (B, Q, Q) entries.  (posx, posy) inside the map.
APG_HD bool may_touch_wall(const unsigned* bits, int pw, int pad, const uint4* table, int Q, int beam, float posx,
                           float posy) {
  const float fx = floorf(posx), fy = floorf(posy);
  const int cx = (int)fx, cy = (int)fy;
  const float qf = (float)Q;
  int qx = (int)((posx - fx) * qf), qy = (int)((posy - fy) * qf);  // exact: Q is a power of two, the fraction is exact
  qx = qx < Q - 1 ? qx : Q - 1;
  qy = qy < Q - 1 ? qy : Q - 1;
#if defined(__CUDA_ARCH__)
  const uint4 t = __ldg(table + (beam * Q + qy) * Q + qx);
#else
  const uint4 t = table[(beam * Q + qy) * Q + qx];
#endif
  const int row0 = (int)(t.w >> 16) - 8;
  const int s = cx + pad - kColBias;  // first column of the 16-bit window (>= 0 because pad >= 7)
  const unsigned* row = bits + (cy + pad + row0) * pw + (s >> 5);
  const int sh = s & 31;
  const unsigned pat[4] = {t.x, t.y, t.z, t.w};
  unsigned acc = 0u;
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const unsigned p16 = (k & 1) ? (pat[k >> 1] >> 16) : (pat[k >> 1] & 0xffffu);
    acc |= window32(row[0], row[1], sh) & p16;
    row += pw;
  }
  return acc != 0u;
}

}  // namespace apg_lidar
