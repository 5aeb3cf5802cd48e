// TEST HARNESS (CPU): the product's rooms generator (csrc/apg_rooms.cuh) compiled as host code; tests/test_maps.py
// compares its output with the reference-generated bank tests/golden/maps_rooms32.npz.
#include <cmath>
#include <cstdint>
#include <vector>

#include "apg_rooms.cuh"

extern "C" int apg_rooms_host(const uint64_t* idx, long long n, int W, int H, int max_rooms, int door_width, uint32_t* out) {
  const int words = (W + 31) / 32;
  double qn[apg_rooms::kMaxRooms + 1];
  for (int k = 0; k <= apg_rooms::kMaxRooms; ++k) qn[k] = std::exp((double)k * std::log(1.0 - apg_rooms::kP));
  std::vector<uint32_t> wall(H * words), door(H * words);
  std::vector<apg_rooms::View> stack(apg_rooms::kMaxStack);
  int bad = 0;
  for (long long i = 0; i < n; ++i)
    bad += apg_rooms::generate(idx[i], W, H, max_rooms, door_width, qn, wall.data(), door.data(), stack.data(),
                               out + (size_t)i * H * words);
  return bad;
}
