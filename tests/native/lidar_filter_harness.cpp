// TEST HARNESS (CPU): compiles the product's beam pre-filter (csrc/apg_lidar_filter.cuh) as host code and checks it
// against the C oracle (oracle/lidar_oracle.c): whenever the filter says "cannot touch a wall", the oracle's ray must
// be a miss (distance = |target - pos| in float32, contact cell (-1, -1)).  Built and driven by tests/test_lidar_filter.py.
#include <cmath>
#include <cstdint>
#include <vector>

#include "apg_lidar_filter.cuh"

extern "C" float apg_oracle_scan_one(const uint32_t* bits, int W, int H, int words, float posx, float posy, float tx,
                                     float ty, int32_t* cell);

// bits: (H, words) unpadded bitmap; pos: (n, 2); dirs: (B, 2); Q: sub-cells per cell and axis.  out[0] = rays the filter rejects, out[1] = rays the
// oracle reports as misses, out[2] = violations; bad: pos/target of the first violation.
extern "C" void apg_filter_check(const uint32_t* bits, int W, int H, int words, int pad, const float* pos, long long n,
                                 const float* dirs, int B, int Q, long long* out, float* bad) {
  std::vector<apg_lidar::BeamPattern> table((size_t)B * Q * Q);
  for (int b = 0; b < B; ++b)
    for (int qy = 0; qy < Q; ++qy)
      for (int qx = 0; qx < Q; ++qx)
        table[((size_t)b * Q + qy) * Q + qx] = apg_lidar::make_beam_pattern(dirs[2 * b], dirs[2 * b + 1], qx, qy, Q);
  static_assert(sizeof(apg_lidar::BeamPattern) == 16, "one uint4 per entry");
  // the padded bitmap exactly as lidar_step_kernel stages it (phase 2)
  const int pw = (W + 2 * pad + 31) / 32, prows = H + 2 * pad;
  std::vector<uint32_t> padded((size_t)prows * pw + 1, 0u);
  for (int yp = 0; yp < prows; ++yp) {
    const int y = yp - pad;
    if (y < 0 || y >= H) continue;
    uint32_t lo = 0u;
    for (int j = 0; j < pw; ++j) {
      const uint32_t hi = j < words ? bits[(size_t)y * words + j] : 0u;
      padded[(size_t)yp * pw + j] = (hi << pad) | (lo >> (32 - pad));
      lo = hi;
    }
  }
  out[0] = out[1] = out[2] = 0;
  for (long long i = 0; i < n; ++i) {
    const float px = pos[2 * i], py = pos[2 * i + 1];
    for (int b = 0; b < B; ++b) {
      const float tx = px + dirs[2 * b], ty = py + dirs[2 * b + 1];
      const bool may = apg_lidar::may_touch_wall(padded.data(), pw, pad, reinterpret_cast<const uint4*>(table.data()), Q, b,
                                                 px, py);
      int32_t cell[2];
      const float d = apg_oracle_scan_one(bits, W, H, words, px, py, tx, ty, cell);
      const float ddx = tx - px, ddy = ty - py;
      const float full = sqrtf(ddx * ddx + ddy * ddy);
      const bool miss = (d == full) && cell[0] == -1 && cell[1] == -1;
      if (!may) {
        out[0]++;
        if (!miss) {
          if (out[2] == 0) { bad[0] = px; bad[1] = py; bad[2] = tx; bad[3] = ty; bad[4] = d; }
          out[2]++;
        }
      }
      if (miss) out[1]++;
    }
  }
}
