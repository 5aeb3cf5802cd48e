// Procedural maps on the device, bit-identical to the reference's generators.
//   FloorMapDatasetRooms.get_data_point     ap_gym/envs/floor_map/floor_map_dataset_rooms.py:25-89  -> csrc/apg_rooms.cuh
//   (host/device code, one thread per map: rooms_kernel below)
// Mazes:
//   FloorMapDatasetMaze.get_data_point      ap_gym/envs/floor_map/floor_map_dataset_maze.py:24-55
// (recursive depth-first `carve` with numpy draws).  One thread generates one maze: the recursion becomes an explicit
// stack of packed frames (cell, direction order, next direction, first flag), the draws are the reference's --
//   rng = default_rng(idx)                       SeedSequence -> PCG64                      (csrc/apg_rng.h)
//   rng.permutation(directions), once per cell   shuffle of arange(4): random_interval(i), i = 3, 2, 1 = masked 32-bit
//                                                draws with rejection (numpy's buffered next_uint32)
//   rng.random() < branching_prob                once per eligible neighbour after the first carved one
// -- in the same order, so that the bank index -> map function is the reference's and maps never have to be generated on
// the host (the reference's dataset has len 2**32; 63x63 mazes take 79 ms each there).  Restated on the CPU in
// oracle/floor_map_oracle.py; pinned against the reference-generated banks tests/golden/maps_maze{21,63}.npz.
#include <vector>

#include <cmath>

#include "apg_common.h"
#include "apg_rng.h"
#include "apg_maze.cuh"
#include "apg_rooms.cuh"

namespace {

__global__ void maze_kernel(const unsigned long long* __restrict__ idx, int n, int W, int H, double branching_prob,
                            unsigned* __restrict__ bits, unsigned* __restrict__ stack, int cells) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int words = (W + 31) >> 5;
  apg_maze::generate(idx[m], W, H, branching_prob, bits + (size_t)m * H * words, stack + (size_t)m * cells);
}

// one thread per rooms map (csrc/apg_rooms.cuh); scratch: two bit planes and the view stack per map
__global__ void rooms_kernel(const unsigned long long* __restrict__ idx, int n, int W, int H, int max_rooms, int door_width,
                             const double* __restrict__ qn_table, unsigned* __restrict__ bits, unsigned* __restrict__ planes,
                             apg_rooms::View* __restrict__ stacks, int* __restrict__ failed) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const int mw = H * ((W + 31) >> 5);
  unsigned* wall = planes + (size_t)m * 2 * mw;
  const int rc = apg_rooms::generate(idx[m], W, H, max_rooms, door_width, qn_table, wall, wall + mw,
                                     stacks + (size_t)m * apg_rooms::kMaxStack, bits + (size_t)m * mw);
  if (rc) atomicAdd(failed, 1);
}

}  // namespace

extern "C" int apg_maps_generate_rooms(int32_t device, const uint64_t* h_indices, int64_t n_maps, int32_t width,
                                       int32_t height, int32_t max_rooms, int32_t door_width, uint32_t* h_bits) {
  APG_CHECK(h_indices && h_bits && n_maps > 0, APG_ERR_INVALID, "apg_maps_generate_rooms: bad argument");
  APG_CHECK(width == height, APG_ERR_INVALID, "rooms maps must be square (the reference transposes half of them)");
  APG_CHECK(width >= 4 && width <= 255, APG_ERR_INVALID, "bad map size");
  APG_CHECK(max_rooms >= 1 && max_rooms <= apg_rooms::kMaxRooms, APG_ERR_INVALID, "max_rooms must be in [1, %d]",
            apg_rooms::kMaxRooms);
  APG_CHECK(door_width >= 1 && door_width <= 64, APG_ERR_INVALID, "bad door_width");
  int ndev = 0;
  cudaError_t de = cudaGetDeviceCount(&ndev);
  if (de != cudaSuccess || ndev <= 0)
    return apg::fail(APG_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(de));
  APG_CHECK(device >= 0 && device < ndev, APG_ERR_INVALID, "device %d out of range", device);
  APG_CUDA(cudaSetDevice(device));
  const int words = (width + 31) / 32;
  const size_t mw = (size_t)height * words, nbits = (size_t)n_maps * mw;
  // exp(n * log(1 - p)) of numpy's binomial inversion, from the host's libm (the same one numpy calls)
  double qn[apg_rooms::kMaxRooms + 1];
  for (int k = 0; k <= apg_rooms::kMaxRooms; ++k) qn[k] = std::exp((double)k * std::log(1.0 - apg_rooms::kP));
  unsigned long long* d_idx = nullptr;
  unsigned *d_bits = nullptr, *d_planes = nullptr;
  apg_rooms::View* d_stacks = nullptr;
  double* d_qn = nullptr;
  int* d_failed = nullptr;
  int failed = 0;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  A(cudaMalloc(&d_idx, n_maps * 8));
  A(cudaMalloc(&d_bits, nbits * 4));
  A(cudaMalloc(&d_planes, nbits * 8));
  A(cudaMalloc(&d_stacks, (size_t)n_maps * apg_rooms::kMaxStack * sizeof(apg_rooms::View)));
  A(cudaMalloc(&d_qn, sizeof(qn)));
  A(cudaMalloc(&d_failed, sizeof(int)));
  if (err == cudaSuccess) {
    A(cudaMemcpy(d_idx, h_indices, n_maps * 8, cudaMemcpyHostToDevice));
    A(cudaMemcpy(d_qn, qn, sizeof(qn), cudaMemcpyHostToDevice));
    A(cudaMemset(d_failed, 0, sizeof(int)));
    rooms_kernel<<<(unsigned)((n_maps + 63) / 64), 64>>>(d_idx, (int)n_maps, width, height, max_rooms, door_width, d_qn, d_bits,
                                                         d_planes, d_stacks, d_failed);
    apg::count_launch();
    A(cudaGetLastError());
    A(cudaMemcpy(h_bits, d_bits, nbits * 4, cudaMemcpyDeviceToHost));
    A(cudaMemcpy(&failed, d_failed, sizeof(int), cudaMemcpyDeviceToHost));
  }
  cudaFree(d_idx); cudaFree(d_bits); cudaFree(d_planes); cudaFree(d_stacks); cudaFree(d_qn); cudaFree(d_failed);
  if (err != cudaSuccess) return apg::fail(APG_ERR_CUDA, "apg_maps_generate_rooms: %s", cudaGetErrorString(err));
  APG_CHECK(failed == 0, APG_ERR_INVALID,
            "%d maps exceed the generator's limits (the reference raises for door_width >= room width)", failed);
  return APG_OK;
}

extern "C" int apg_maps_generate_maze(int32_t device, const uint64_t* h_indices, int64_t n_maps, int32_t width,
                                      int32_t height, double branching_prob, uint32_t* h_bits) {
  APG_CHECK(h_indices && h_bits && n_maps > 0, APG_ERR_INVALID, "apg_maps_generate_maze: bad argument");
  APG_CHECK(width >= 3 && height >= 3 && width <= 255 && height <= 255 && (width & 1) && (height & 1), APG_ERR_INVALID,
            "Width and height must be odd (and at most 255).");
  int ndev = 0;
  cudaError_t de = cudaGetDeviceCount(&ndev);
  if (de != cudaSuccess || ndev <= 0)
    return apg::fail(APG_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(de));
  APG_CHECK(device >= 0 && device < ndev, APG_ERR_INVALID, "device %d out of range", device);
  APG_CUDA(cudaSetDevice(device));
  const int words = (width + 31) / 32;
  const int cells = ((width - 1) / 2) * ((height - 1) / 2) + 1;
  const size_t nbits = (size_t)n_maps * height * words;
  unsigned long long* d_idx = nullptr;
  unsigned *d_bits = nullptr, *d_stack = nullptr;
  cudaError_t err = cudaSuccess;
  auto A = [&](cudaError_t r) { if (err == cudaSuccess) err = r; };
  A(cudaMalloc(&d_idx, n_maps * 8));
  A(cudaMalloc(&d_bits, nbits * 4));
  A(cudaMalloc(&d_stack, (size_t)n_maps * cells * 4));
  if (err == cudaSuccess) {
    A(cudaMemcpy(d_idx, h_indices, n_maps * 8, cudaMemcpyHostToDevice));
    maze_kernel<<<(unsigned)((n_maps + 63) / 64), 64>>>(d_idx, (int)n_maps, width, height, branching_prob, d_bits, d_stack, cells);
    apg::count_launch();
    A(cudaGetLastError());
    A(cudaMemcpy(h_bits, d_bits, nbits * 4, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_idx); cudaFree(d_bits); cudaFree(d_stack);
  if (err != cudaSuccess) return apg::fail(APG_ERR_CUDA, "apg_maps_generate_maze: %s", cudaGetErrorString(err));
  return APG_OK;
}
